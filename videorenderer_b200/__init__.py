"""videorenderer_b200 -- Python face of libmpcvr_b200.so, the B200-native frame path
(YUV->RGB convert, PQ/HLG->SDR tone-map, separable resize, ordered dither) behind the
reference's CVideoProcessor / IVideoRenderer verbs.

The compute lives in hand-written sm_100a CUDA inside the shared library; this package only
loads it (ctypes) and mirrors the C ABI of include/mpcvr_b200.h.  There is no CPU fallback:
if the library is missing, importing `lib()` raises.
"""
from ._abi import (lib, lib_path, Settings, MediaType, Sample, Rect, ProcAmp, PlanInfo, HRESULT,  # noqa: F401
                   CF, SURF, TEXFMT, CHROMA, UPSCALE, DOWNSCALE, PATH, exfmt, MpcvrError)
from .processor import VideoProcessor  # noqa: F401

__all__ = ["lib", "lib_path", "VideoProcessor", "Settings", "MediaType", "Sample", "Rect", "ProcAmp", "PlanInfo",
           "CF", "SURF", "TEXFMT", "CHROMA", "UPSCALE", "DOWNSCALE", "PATH", "exfmt", "MpcvrError", "HRESULT"]
