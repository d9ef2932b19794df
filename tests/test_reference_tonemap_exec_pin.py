"""Tenth pin: Shaders/d3d11/ps_hdr10_tonemap.hlsl, executed.

The HDR10 local tone-mapping step (TextureCopyRect with m_pPSHDR10ToneMapping, DX11VideoProcessor.cpp:3359-3367) is a 336-line shader file:
PQ -> linear, optional Dolby Vision trims, one of six curves (ACES, Reinhard, Habel, Moebius, the BT.2390 EETF with its Hermite knee, the
ST 2094-10 EETF with its adaptive knee and 3x3 system), -> PQ.  oracle/Makefile compiles the file where it lies, as C++, over
oracle/ref_stub/hlsl_vec.h (the edits are the syntax edits listed in gen_convert_shaders.py) into oracle/_ref/libref_hlsl_convert.so.
Here it runs over frames of PQ texels -- uniform noise with values outside [0, 1], every 10-bit code pattern, exact 0 and 1 -- for every
`selection` and for metadata sets that take each early-out and fallback of the text (display brighter than the content, maxCLL below 10,
maxFALL absent, a knee that clamps), and the oracle's tm_pixel (oracle/orc_core.inc) must reproduce it in LITERAL mode BIT FOR BIT.
SPEC mode (the contract of the CUDA kernel k_tonemap, tests/test_gpu_parity.py) is held to a measured distance.

With the Dolby Vision trims bound (cbuffer b1) the same holds, NaN pixels included.  The shader's DolbyVisionTrims has no guard (see
tests/test_reference_convert_exec_pin.py): a negative pow base or a zero luminance make every channel NaN, a zero channel under a
non-positive saturation gain or a value past ST2084ToLinear's pole make that channel NaN.  Curves 1-4 saturate the NaN to 0 and carry on;
BT.2390 spreads it over the pixel when its curve is active and passes it through per channel otherwise; ST 2094-10 always passes it through
per channel (the NaN luminance fails `x_nits > 0`).  tm_pixel reproduces each case (a NaN result is held as 0, which is what a UNORM
target stores); the test finds the poisoned pixels with the shader itself (selection 6 on a display brighter than the content is trims
only) and checks every pixel of every curve."""
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


def _libs():
    R = C.CDLL(REF); L = O.lib()
    R.ref_tonemap_run.argtypes = [FP, C.c_int, C.c_int, C.c_float * 6, FP, FP]
    L.orc_test_tonemap.argtypes = [C.c_int, C.c_float * 6, FP, FP, FP, C.c_int]
    return R, L


def _p(a):
    return a.ctypes.data_as(FP) if a is not None else None


def _frame(seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    tex = rng.uniform(-0.02, 1.02, (H, W, 4)).astype(f32)
    tex[:8] = rng.integers(0, 1024, (8, W, 4)).astype(f32) / f32(1023)          # what a 10-bit intermediate texture returns
    tex[8:12] = (rng.uniform(0, 1, (4, W, 1)) ** 3).astype(f32)                    # grays, dark-heavy
    tex[12, :4] = 0; tex[12, 4:8] = 1
    return np.ascontiguousarray(tex)


def _pq0(L):
    """LinearToST2084(0) in LITERAL arithmetic = c1^m2 = 7.31e-7"""
    L.orc_test_transfer.argtypes = [C.c_int, C.c_int, FP, FP, C.c_int, C.c_float]
    x = np.zeros(1, f32); y = np.zeros(1, f32)
    assert L.orc_test_transfer(1, 1, _p(x), _p(y), 1, 10000.0) == 0
    return y[0]


def _equal_or_nan_black(ref, lit, pq0):
    """bit-equal, or the reference's float is NaN where the oracle holds PQ(0).  The ST 2094-10 curve passes through (MasteringMinLuminance, 0):
    pixels darker than that get a NEGATIVE luminance ratio, and LinearToST2084 of a negative value is NaN in the shader; the oracle's pow
    clamps a non-positive base, which yields LinearToST2084(0) = 7.31e-7.  Every UNORM render target writes both as code 0 (and the final pass
    floors 7.31e-7 * Q + dither to 0 as well), so the pixels the reference PRESENTS are the same; only an RGBA16F back buffer -- an extension
    the reference does not have -- would tell them apart."""
    return (ref.view(np.uint32) == lit.view(np.uint32)) | (np.isnan(ref) & (lit == pq0))


def _run(R, L, tex, hp, sel, l2, mode):
    hpa = (C.c_float * 6)(*hp, sel)
    l2a = np.array(l2, f32) if l2 is not None else None
    ref = np.empty((H, W, 4), f32)
    assert R.ref_tonemap_run(_p(tex), W, H, hpa, _p(l2a), _p(ref)) == 0
    rgb = np.ascontiguousarray(tex[..., :3]); o = np.empty((H, W, 3), f32)
    assert L.orc_test_tonemap(mode, hpa, _p(l2a), _p(rgb), _p(o), W * H) == 0
    return np.ascontiguousarray(ref[..., :3]), o


# MasteringMinLuminanceNits, MasteringMaxLuminanceNits, maxCLL, maxFALL, displayMaxNits
METADATA = {
    "typical":               (0.005, 1000.0, 1000.0, 400.0, 300.0),
    "bright_master":         (0.0, 4000.0, 2500.0, 120.0, 600.0),
    "display_covers_content": (0.01, 1000.0, 800.0, 200.0, 1000.0),        # BT.2390 / ST 2094-10 return the colour unchanged
    "cll_below_10":          (0.005, 1000.0, 5.0, 400.0, 500.0),           # BT.2390's safeMaxCLL fallback
    "no_fall":               (0.005, 1200.0, 1100.0, 0.0, 400.0),          # ST 2094-10: knee from the default, division by maxFALL = 0 in the global curves
    "knee_clamps":           (0.05, 1000.0, 1000.0, 990.0, 150.0),
}


@pytest.mark.parametrize("sel", [1, 2, 3, 4, 5, 6, 0, 9])
@pytest.mark.parametrize("meta", sorted(METADATA))
def test_tonemap_shader_equals_the_oracle(sel, meta):
    R, L = _libs()
    tex = _frame(11 + sel)
    ref, lit = _run(R, L, tex, METADATA[meta], sel, None, 1)
    same = _equal_or_nan_black(ref, lit, _pq0(L))
    assert same.all(), (sel, meta, int((~same).sum()), ref[~same][:4], lit[~same][:4])
    assert int(np.isnan(ref).sum()) == 0 or sel == 6          # only the ST 2094-10 curve goes negative (below the mastering black level)
    _, spec = _run(R, L, tex, METADATA[meta], sel, None, 0)
    ok = ~(np.isnan(ref) | np.isnan(spec))
    err = np.abs(ref.astype(np.float64) - spec)[ok]
    # SPEC against the executed text, PQ output in [0, 1]: the two pow per direction move it by ~1e-6; near black PQ's slope amplifies
    assert np.percentile(err, 99.9) <= 1e-4 and err.max() <= 2e-3, (sel, meta, np.percentile(err, 99.9), err.max())


TRIMS = {       # ChromaWeight, SaturationGain, TrimSlope, TrimOffset, TrimPower
    "mild":            (0.05, 0.1, 1.02, 0.004, 0.97),
    "negative_offset": (0.05, 0.1, 1.02, -0.004, 0.97),
    "negative_satgain": (-0.1, -0.1, 0.95, 0.01, 1.1),
    "zero_satgain":    (0.1, 0.0, 1.0, -0.001, 1.0),
    "bright":          (0.0, 0.0, 1.5, 0.5, 1.5),           # values past the pole of ST2084ToLinear (~1.99)
    "bright_sat":      (0.5, 0.5, 1.4, 0.3, 1.5),
}


@pytest.mark.parametrize("sel", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("meta", sorted(METADATA))
@pytest.mark.parametrize("trims", sorted(TRIMS))
def test_tonemap_shader_with_dolby_vision_trims(sel, meta, trims):
    """every pixel, the NaN ones included: where the reference's float is NaN the oracle holds 0 (or LinearToST2084(0)), both code 0 in any
    UNORM target; where the reference saturates its NaN to 0 and carries on (curves 1-4) the oracle must produce the very same number"""
    R, L = _libs()
    tex = _frame(40 + sel)
    hp = METADATA[meta]
    # the pixels whose trims make NaN in the reference's text: selection 6 on a display brighter than maxCLL is PQ -> linear -> trims -> PQ
    probe, _ = _run(R, L, tex, (hp[0], hp[1], hp[2], hp[3], 10000.0), 6, TRIMS[trims], 1)
    poisoned = np.isnan(probe).any(axis=2)
    ref, lit = _run(R, L, tex, hp, sel, TRIMS[trims], 1)
    pq0 = _pq0(L)
    same = (ref.view(np.uint32) == lit.view(np.uint32)) | (np.isnan(ref) & ((lit == pq0) | (lit == 0)))
    assert same.all(), (sel, meta, trims, int((~same).sum()), tex[np.nonzero(~same)[:2]][:3], ref[~same][:4], lit[~same][:4])
    if sel <= 4:
        assert not np.isnan(ref).any()          # saturate() has swallowed the NaN: the comparison above was on numbers
    n = int(poisoned.sum())
    if trims == "mild":
        assert n == 0
    elif trims in ("negative_offset", "bright"):
        assert 0 < n < W * H, n
    # SPEC takes the same decisions
    _, spec = _run(R, L, tex, hp, sel, TRIMS[trims], 0)
    nanpix = np.isnan(ref)
    assert np.all((spec[nanpix] == 0) | (np.abs(spec[nanpix] - pq0) < 1e-8))
