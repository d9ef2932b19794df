// tests/host_shim/cuda_host_shim.h -- TEST INFRASTRUCTURE ONLY.
// Lets the product's DEVICE headers (videorenderer_b200/csrc/mpcvr_math.cuh, pixel_ops.cuh, tonemap_ops.cuh) compile for the HOST with g++, so
// that the per-pixel functions the CUDA kernels call can be executed on a machine without a GPU and compared with the oracle's SPEC mode
// bit for bit (tests/test_device_code_on_host.py).  The headers are included unmodified; this file supplies what nvcc's device compilation
// environment provides:
//   * the round-to-nearest intrinsics __fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn / __fmaf_rn as the IEEE operations they are (the unit is
//     compiled with -ffp-contract=off, so nothing else fuses), the bit casts, __ldg as a plain load, __byte_perm, integer min / max;
//   * `asm("rcp.approx.ftz.f32 ...")` in div_rn_safe (the only inline PTX in those headers) becomes r = 1.0f / y: the residual-correction
//     sequence that follows produces the correctly rounded quotient from any reciprocal estimate of that quality, which is what the
//     function is for (and what the GPU tests hold it to).
// This is NOT a CPU path of the product: nothing under videorenderer_b200/ includes it, and it covers per-pixel arithmetic only -- no
// kernel, no launch, no memory system.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline int32_t __float_as_int(float f) { int32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int32_t u) { float f; memcpy(&f, &u, 4); return f; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t v = ((uint64_t)b << 32) | a; uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
#define asm(...) r = 1.0f / y
