"""Eleventh pin: the six post-scale correction shaders, executed.

On the D3D11-video-processor path the reference repairs or completes the driver's conversion with one pixel shader after scaling
(m_pPSCorrection, DX11VideoProcessor.cpp:1893-1931, drawn at :3354-3357): ps_convert_pq_to_sdr, ps_fixconvert_pq_to_sdr,
ps_fixconvert_hlg_to_sdr, ps_fix_bt2020, ps_convert_hlg_to_pq, ps_fix_ycgco (Shaders/d3d11/), built on Shaders/convert/conv_matrix.hlsl,
colorspace_gamut_conversion.hlsl and the transfer includes.  The files compile where they lie, as C++ over oracle/ref_stub/hlsl_vec.h
(float3x3 / float4x4 with row indexing, transpose, matrix products), into oracle/_ref/libref_hlsl_convert.so.

Two things are separated here, because only one of them can be decided on a CPU:

* WHICH OPERATIONS, IN WHICH ORDER -- decided: with the same matrices in both, the oracle's corr_pixel (oracle/orc_core.inc) equals every one
  of the six shaders BIT FOR BIT in LITERAL mode on frames of full-range pixels.
* WHICH CONSTANTS -- the shaders' matrices (`static const float4x4 fix_bt2020_matrix = mul(ycbcr2020nc_rgb, rgb_ycbcr709)`, the zimg-style
  primaries -> XYZ -> inverse chain of colorspace_gamut_conversion.hlsl) are folded by fxc at shader-compile time.  The oracle evaluates
  them in float64 and rounds once (fxc's warning X4122, which that file silences, speaks of "double precision"); evaluated in float32 --
  what this C++ build does -- entries that come out of cancellations differ by up to ~80 ULP.  Which one d3dcompiler_47 produces cannot be
  observed here; the test records the distance so that a reader with fxc can settle it: the pixel effect is below 5e-4 of full scale (measured 3.1e-4, HLG to SDR near black)."""
import ctypes as C
import os

import numpy as np
import pytest

import orclib as O

REF = os.path.join(O.ORC_DIR, "_ref", "libref_hlsl_convert.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_hlsl_convert.so not built (reference tree absent at build time)")

FP = C.POINTER(C.c_float)
f32 = np.float32
W, H = 64, 32
NAMES = {1: "convert_pq_to_sdr", 2: "fixconvert_pq_to_sdr", 3: "fixconvert_hlg_to_sdr", 4: "fix_bt2020", 5: "convert_hlg_to_pq", 6: "fix_ycgco"}


def _p(a):
    return a.ctypes.data_as(FP) if a is not None else None


def _libs():
    R = C.CDLL(REF); L = O.lib()
    R.ref_correction_run.argtypes = [C.c_int, FP, C.c_int, C.c_int, C.c_float, FP, FP, FP]
    L.orc_test_correction.argtypes = [C.c_int, C.c_int, C.c_float, FP, FP, C.c_int, FP]
    return R, L


def _frame(seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    tex = rng.uniform(-0.05, 1.05, (H, W, 4)).astype(f32)
    tex[:8] = rng.integers(0, 1024, (8, W, 4)).astype(f32) / f32(1023)
    tex[8, :4] = 0; tex[8, 4:8] = 1
    return np.ascontiguousarray(tex)


@pytest.mark.parametrize("corr", sorted(NAMES), ids=[NAMES[k] for k in sorted(NAMES)])
@pytest.mark.parametrize("nits", [125, 80, 400])
def test_correction_shader_operations_equal_the_oracle(corr, nits):
    R, L = _libs()
    tex = _frame(50 + corr)
    rgb = np.ascontiguousarray(tex[..., :3])
    lum = f32(10000.0) / f32(nits)
    lit = np.empty((H, W, 3), f32); oc = np.zeros(18, f32)
    assert L.orc_test_correction(1, corr, lum, _p(rgb), _p(lit), W * H, _p(oc)) == 0
    ref = np.empty((H, W, 4), f32)
    assert R.ref_correction_run(corr, _p(tex), W, H, lum, _p(ref), None, _p(oc)) == 0        # the text, with the oracle's matrices
    a = np.ascontiguousarray(ref[..., :3])
    assert np.array_equal(a.view(np.uint32), lit.view(np.uint32)), (NAMES[corr], int((a.view(np.uint32) != lit.view(np.uint32)).sum()))
    spec = np.empty((H, W, 3), f32)
    assert L.orc_test_correction(0, corr, lum, _p(rgb), _p(spec), W * H, None) == 0
    err = np.abs(a.astype(np.float64) - spec)
    assert np.percentile(err, 99.9) <= 3e-4 and err.max() <= 5e-3, (NAMES[corr], np.percentile(err, 99.9), err.max())


def test_correction_shader_constants_float32_against_float64_folding():
    R, L = _libs()
    tex = _frame(7); rgb = np.ascontiguousarray(tex[..., :3])
    worst_ulp = 0; worst_px = 0.0
    for corr in (1, 2, 3, 4, 6):
        lit = np.empty((H, W, 3), f32); oc = np.zeros(18, f32)
        assert L.orc_test_correction(1, corr, f32(80.0), _p(rgb), _p(lit), W * H, _p(oc)) == 0
        own = np.empty((H, W, 4), f32); rc = np.zeros(18, f32)
        assert R.ref_correction_run(corr, _p(tex), W, H, f32(80.0), _p(own), _p(rc), None) == 0   # the text with its own fp32-folded matrices
        used = rc != 0                                                                             # the matrices this shader has
        assert np.allclose(rc[used], oc[used], rtol=2e-5, atol=1e-7), (NAMES[corr], rc, oc)      # the same matrices up to folding precision
        worst_ulp = max(worst_ulp, int(np.abs(rc.view(np.int32)[used].astype(np.int64) - oc.view(np.int32)[used]).max()))
        worst_px = max(worst_px, float(np.nanmax(np.abs(own[..., :3].astype(np.float64) - lit))))
    assert worst_ulp <= 128 and worst_px <= 5e-4, (worst_ulp, worst_px)
