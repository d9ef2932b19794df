// oracle/ref_stub/dovi_prelude.h -- what Source/DX11VideoProcessor.cpp:990-1122 (SetShaderDoviCurvesPoly / SetShaderDoviCurves: Dolby Vision
// reshaping metadata -> the PS_DOVI_CURVE constant buffers) needs to compile where it lies, outside its class: the DirectXMath
// float4, the ASSERT macro and an m_Dovi object.  The reference's own struct definitions follow on the compiler's stdin (sed -n of
// Include/IMediaSideData.h and Source/Shaders.h), then the loops themselves between the halves of dovi_wrap_*.inc.  Nothing of the
// reference is copied into the repository.  Test infrastructure (oracle/Makefile target "ref").
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <cmath>
namespace DirectX { struct XMFLOAT4 { float x, y, z, w; }; }
#define ASSERT(x) ((void)0)
typedef unsigned int UINT;
