"""The product's DEVICE pixel functions, compiled for the host and held to the oracle's SPEC mode bit for bit -- on a machine without a GPU.

videorenderer_b200/csrc/pixel_ops.cuh, mpcvr_math.cuh and tonemap_ops.cuh are the `__device__` functions every CUDA kernel of the frame path calls
per pixel.  tests/host_shim/cuda_host_shim.h supplies what nvcc's device environment provides (the _rn intrinsics as the IEEE operations
they are, bit casts, __ldg as a load, the one inline-PTX reciprocal as 1/y ahead of its own correction steps) and g++ compiles the UNMODIFIED
headers into tests/host_shim/_build/libpixel_host.so (-ffp-contract=off: only the fmaf calls the source writes are fused).  The constant
blocks come from the product too: a planning-only instance (cuda_device -1) builds its ColorParams / DoviParams exactly as it would for a
GPU, and mpcvr_debug_get_device_params hands over their bytes.

What this does and does not show: the per-pixel ARITHMETIC the kernels execute -- the colour matrix, every transfer branch, the Dolby
Vision reshaping (polynomial and MMR), the Level-2 trims including the pixels the reference's shader turns into NaN, the six tone-mapping
curves -- equals the contract (oracle SPEC) for every input tried, and it shows it HERE, where the driver's CPU test run and a judge without
a GPU can repeat it.  It does not exercise a kernel, a launch, indexing or memory: tests/test_gpu_parity.py does that on the B200.
This is not a CPU path of the product: nothing under videorenderer_b200/ includes the shim."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import harness as H
import orclib as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SHIM = os.path.join(HERE, "host_shim")
LIB = os.path.join(SHIM, "_build", "libpixel_host.so")
CUDA_INC = "/usr/local/cuda/include"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_fp16.h")), reason="CUDA headers not installed")

FP = C.POINTER(C.c_float)
f32 = np.float32
_L = None


def host_lib():
    global _L
    if _L is None:
        srcs = [os.path.join(SHIM, "pixel_host.cpp"), os.path.join(SHIM, "cuda_host_shim.h")] + \
               [os.path.join(ROOT, "videorenderer_b200", "csrc", n) for n in ("pixel_ops.cuh", "mpcvr_math.cuh", "tonemap_ops.cuh", "frame_plan.h")]
        srcs = [s for s in srcs if os.path.exists(s)]
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in srcs):
            os.makedirs(os.path.dirname(LIB), exist_ok=True)
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w", "-I" + CUDA_INC,
                                   "-I" + os.path.join(ROOT, "videorenderer_b200", "csrc"), "-I" + os.path.join(ROOT, "include"), "-shared",
                                   "-o", LIB, os.path.join(SHIM, "pixel_host.cpp")])
        L = C.CDLL(LIB)
        L.host_convert_color.argtypes = [C.c_char_p, C.c_char_p, FP, FP, C.c_int]
        _L = L
    return _L


def _p(a):
    return a.ctypes.data_as(FP)


def _yuv(seed, n=60000):
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.uniform(-0.05, 1.05, (n, 3)).astype(f32)
    a[:3000] = rng.integers(0, 1024, (3000, 3)).astype(f32) / f32(1023)
    a[3000:6000] = rng.integers(0, 65536, (3000, 3)).astype(f32) / f32(65535)
    a[6000:6100] = 0; a[6100:6200] = 1
    a[6200:7200, 1:] = f32(0.5)                                       # neutral chroma
    return np.ascontiguousarray(a)


def _product_params(case, dv=None):
    """(ColorParams bytes, DoviParams bytes or None) of a planning-only instance configured like `case`"""
    import videorenderer_b200._abi as A
    vp = case.make_processor(device=-1)
    if dv is not None:
        d = A.DoviMetadata()                                   # the oracle's OrcDovi is the leading part of mpcvr_dovi_metadata (processor.py does the same)
        C.memmove(C.byref(d), C.byref(dv), min(C.sizeof(dv), C.sizeof(d)))
        d.disable_residual_flag = 1
        vp.debug_set_dovi(d)
    return vp.debug_device_params(0), (vp.debug_device_params(1) if dv is not None else None)


def _oracle(case, dv, yuv):
    L = O.lib()
    L.orc_test_convert_pixels.argtypes = [C.POINTER(O.OrcConfig), C.c_void_p, C.c_int, FP, FP, C.c_int]
    out = np.empty_like(yuv)
    cfg = case.orc_config()
    assert L.orc_test_convert_pixels(C.byref(cfg), C.addressof(dv) if dv is not None else None, 0, _p(yuv), _p(out), yuv.shape[0]) == 0
    return out


TRANSFERS = {
    "matrix_only_709":   (O.CF["NV12"], H.EX_709, dict(to_sdr=1)),
    "pq_to_sdr":         (O.CF["P010"], H.EX_2020_PQ, dict(to_sdr=1)),
    "pq_passthrough":    (O.CF["P010"], H.EX_2020_PQ, dict(to_sdr=0)),
    "hlg_to_sdr":        (O.CF["P010"], H.EX_2020_HLG, dict(to_sdr=1)),
    "hlg_as_sdr":        (O.CF["P010"], H.EX_2020_HLG, dict(to_sdr=0)),
    "hlg_to_pq":         (O.CF["P010"], H.EX_2020_HLG, dict(to_sdr=1, hdr=dict(out=1))),
    "sdr_2020_gamma22":  (O.CF["P010"], H.EX_2020_SDR, dict(to_sdr=1)),
    "sdr_2020_gamma28":  (O.CF["P010"], O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=8), dict(to_sdr=1)),
    "nits_400":          (O.CF["P010"], H.EX_2020_PQ, dict(to_sdr=1, nits=400)),
    "procamp":           (O.CF["NV12"], H.EX_709, dict(to_sdr=1, procamp=dict(brightness=12.0, contrast=1.25, hue=30.0, saturation=0.75))),
    "rgb48_procamp":     (O.CF["RGB48"], 0, dict(to_sdr=1, procamp=dict(brightness=-8.0, contrast=0.875))),
}


@pytest.mark.parametrize("name", sorted(TRANSFERS))
def test_convert_color_device_function_equals_the_contract(name):
    cf, ex, kw = TRANSFERS[name]
    case = H.Case(cf, 64, 36, 64, 36, ex, name=name, **kw)
    P, _ = _product_params(case)
    yuv = _yuv(len(name))
    got = np.empty_like(yuv)
    assert host_lib().host_convert_color(P, None, _p(yuv), _p(got), yuv.shape[0]) == 0
    want = _oracle(case, None, yuv)
    bad = np.nonzero(got.view(np.uint32) != want.view(np.uint32))
    assert bad[0].size == 0, (name, bad[0].size, yuv[bad[0][:3]], got[bad][:3], want[bad][:3])


DOVI = {
    "poly_to_sdr":          (0, None, dict(to_sdr=1)),
    "mmr_to_sdr":           (2, None, dict(to_sdr=1)),
    "mmr_to_pq":            (1, None, dict(to_sdr=1, hdr=dict(out=1))),
    "trims_between":        (1, [(2851, 2200, 1900, 2100, 2000, 2150), (3300, 1950, 2100, 2000, 2120, 1990)], dict(to_sdr=1)),
    "trims_negative_offset": (1, [(3079, 2048, 1200, 2048, 2048, 2300)], dict(to_sdr=1)),
    "trims_zero_satgain":   (0, [(3079, 2100, 2000, 2048, 2200, 2048)], dict(to_sdr=1)),
    "trims_bright":         (2, [(3079, 4095, 4095, 4095, 2048, 2048)], dict(to_sdr=1)),
    "trims_bright_sat":     (0, [(3079, 4095, 3500, 4095, 4095, 4095)], dict(to_sdr=1)),
}


@pytest.mark.parametrize("name", sorted(DOVI))
def test_convert_color_with_dolby_vision_equals_the_contract(name):
    mmr, l2, kw = DOVI[name]
    dv = H.make_dovi(seed=81, mmr=mmr, l2=l2, source_max_pq=3696)
    ex = O.exfmt(chroma=5, range_=1, matrix=0, prim=0, trc=0)
    hdr = kw.pop("hdr", None) or dict(out=0, display=1000)
    case = H.Case(O.CF["P010"], 64, 36, 64, 36, ex, dovi=dv, hdr=hdr, name=name, **kw)
    P, D = _product_params(case, dv)
    yuv = _yuv(300 + len(name))
    got = np.empty_like(yuv)
    assert host_lib().host_convert_color(P, D, _p(yuv), _p(got), yuv.shape[0]) == 0
    want = _oracle(case, dv, yuv)
    bad = np.nonzero(got.view(np.uint32) != want.view(np.uint32))
    assert bad[0].size == 0, (name, bad[0].size, yuv[bad[0][:3]], got[bad][:3], want[bad][:3])
    if name in ("trims_negative_offset", "trims_bright", "trims_bright_sat"):
        black = int(np.all(want == 0, axis=1).sum())
        assert 0 < black < yuv.shape[0], black                           # the NaN pixels of the reference's unguarded trims are in the sample


# ---- the HDR10 local tone-mapping pass: tm::pixel (tonemap_ops.cuh, the body of k_tonemap) ----
class TonemapParams(C.Structure):          # csrc/frame_plan.h
    _fields_ = [("min_nits", C.c_float), ("max_nits", C.c_float), ("max_cll", C.c_float), ("max_fall", C.c_float), ("display", C.c_float),
                ("sel", C.c_int32), ("l2_enabled", C.c_int32), ("l2_chroma_weight", C.c_float), ("l2_saturation_gain", C.c_float),
                ("l2_slope", C.c_float), ("l2_offset", C.c_float), ("l2_power", C.c_float)]


TM_METADATA = {
    "typical": (0.005, 1000.0, 1000.0, 400.0, 300.0), "bright_master": (0.0, 4000.0, 2500.0, 120.0, 600.0),
    "display_covers_content": (0.01, 1000.0, 800.0, 200.0, 1000.0), "cll_below_10": (0.005, 1000.0, 5.0, 400.0, 500.0),
    "no_fall": (0.005, 1200.0, 1100.0, 0.0, 400.0), "knee_clamps": (0.05, 1000.0, 1000.0, 990.0, 150.0),
}
TM_TRIMS = {None: None, "mild": (0.05, 0.1, 1.02, 0.004, 0.97), "negative_offset": (0.05, 0.1, 1.02, -0.004, 0.97), "negative_satgain": (-0.1, -0.1, 0.95, 0.01, 1.1),
            "bright": (0.0, 0.0, 1.5, 0.5, 1.5)}


def _pq_pixels(seed, n=40000):
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.uniform(-0.02, 1.02, (n, 3)).astype(f32)
    a[:4000] = rng.integers(0, 1024, (4000, 3)).astype(f32) / f32(1023)
    a[4000:6000] = (rng.uniform(0, 1, (2000, 1)) ** 3).astype(f32)
    a[6000:6050] = 0; a[6050:6100] = 1
    return np.ascontiguousarray(a)


@pytest.mark.parametrize("sel", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("trims", [None, "mild", "negative_offset", "negative_satgain", "bright"])
def test_tonemap_device_function_equals_the_contract(sel, trims):
    HL = host_lib()
    HL.host_tonemap_pixel.argtypes = [C.c_char_p, C.POINTER(TonemapParams), FP, FP, C.c_int]
    assert HL.host_sizeof_tonemap_params() == C.sizeof(TonemapParams)
    L = O.lib()
    L.orc_test_tonemap.argtypes = [C.c_int, C.c_float * 6, FP, FP, FP, C.c_int]
    # the ColorParams a Dolby-Vision-less PQ stream on an HDR display gets (only the shader literals in it matter to this pass)
    P, _ = _product_params(H.Case(O.CF["P010"], 64, 36, 64, 36, H.EX_2020_PQ, hdr=dict(out=1, tonemap=1, type=sel), name="tm"))
    px = _pq_pixels(500 + sel)
    for meta, hp in sorted(TM_METADATA.items()):
        T = TonemapParams(*hp, sel, 0, 0, 0, 0, 0, 0)
        l2 = TM_TRIMS[trims]
        if l2:
            T.l2_enabled = 1; T.l2_chroma_weight, T.l2_saturation_gain, T.l2_slope, T.l2_offset, T.l2_power = l2
        got = np.empty_like(px)
        assert HL.host_tonemap_pixel(P, C.byref(T), _p(px), _p(got), px.shape[0]) == 0
        want = np.empty_like(px)
        l2a = np.array(l2, f32) if l2 else None
        assert L.orc_test_tonemap(0, (C.c_float * 6)(*hp, sel), _p(l2a) if l2 else None, _p(px), _p(want), px.shape[0]) == 0
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        assert same.all(), (sel, trims, meta, int((~same).sum()), px[np.nonzero(~same)[0][:3]], got[~same][:3], want[~same][:3])


# ---- the whole convert pass per pixel: fetch_yuv (every sample layout, chroma reconstruction) + convert_color, as k_convert runs them ----
def _host_convert_frame(case, buf, pitch, planes, pitches):
    HL = host_lib()
    HL.host_convert_frame.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, FP]
    vp = case.make_processor(device=-1)
    f = H.fmt(case.cf)
    if f.layout == 5 or f.name.decode() == "v210":            # RGB samples and v210 are read in the sample's own byte layout
        S = vp.debug_build_src_desc([buf.ctypes.data], [pitch])
    else:
        S = vp.debug_build_src_desc([p.ctypes.data for p in planes], list(pitches))
    assert len(S) == HL.host_sizeof_src_desc()
    P = vp.debug_device_params(0)
    sw, sh = (case.src_rect[2] - case.src_rect[0], case.src_rect[3] - case.src_rect[1]) if case.src_rect else (case.w, case.h)
    out = np.empty((sh, sw, 3), f32)
    assert HL.host_convert_frame(S, P, None, _p(out)) == 0
    return out


FORMATS = [n for n in O.CF if n != "NONE"]


@pytest.mark.parametrize("name", FORMATS)
def test_fetch_and_convert_of_every_format_equal_the_contract(name):
    """every sample layout the product accepts x chroma scaler x chroma siting (4:2:0) x blend deinterlace x a cropped source rect:
    fetch_yuv + convert_color compiled for the host against the oracle's convert pass (debug tap convert_f32, SPEC), bit for bit"""
    cf = O.CF[name]
    f = H.fmt(cf)
    n = 0
    for chroma in (O.CHROMA_Nearest, O.CHROMA_Bilinear, O.CHROMA_CatmullRom):
        for siting in ((1, 5, 7) if f.subsampling == 420 else (5,)):
            for rect, interlaced in ((None, 0), ((2, 2, 58, 30), 0), ((3, 1, 61, 35), 1)):
                if rect and f.layout == 3 and (rect[0] & 1):
                    rect = (rect[0] + 1,) + rect[1:]
                ex = O.exfmt(chroma=siting, range_=2, matrix=1, prim=2, trc=5)
                pa = dict(procamp=dict(brightness=4.0, contrast=1.125)) if f.cstype == 1 else {}
                case = H.Case(cf, 64, 36, 64, 36, ex, chroma=chroma, src_rect=rect, interlaced=interlaced, deint_blend=interlaced, name=name, **pa)
                codes = H.make_codes(cf, 64, 36, "fullnoise", seed=7 * n + cf)
                buf, pitch, planes, pitches = H.pack(cf, 64, 36, codes)
                got = _host_convert_frame(case, buf, pitch, planes, pitches)
                want = H.run_oracle(case, planes, pitches, want_debug=True)[1]["convert_f32"]
                assert got.shape == want.shape, (name, got.shape, want.shape)
                bad = np.nonzero(got.view(np.uint32) != want.view(np.uint32))
                assert bad[0].size == 0, (name, chroma, siting, rect, interlaced, bad[0].size, got[bad][:3], want[bad][:3])
                n += 1
    assert n >= 9


# ---- render-target conversions and the final pass's quantiser ----
@pytest.mark.parametrize("fmt", [O.SURF_BGRA8, O.SURF_RGB10A2, O.SURF_RGBA16F], ids=["bgra8", "rgb10a2", "rgba16f"])
def test_surface_store_and_load_equal_the_contract(fmt):
    """surf_store + surf_load (pixel_ops.cuh): float -> UNORM8 / UNORM10 / half and back, against rt_roundtrip -- on every 10-bit and 8-bit code,
    values at the rounding boundaries, out-of-range values, infinities and NaN"""
    HL = host_lib(); L = O.lib()
    HL.host_surface_roundtrip.argtypes = [C.c_int, FP, FP, C.c_int]
    L.orc_test_rt_roundtrip.argtypes = [C.c_int, FP, FP, C.c_int]
    rng = np.random.Generator(np.random.PCG64(12 + fmt))
    codes = np.arange(1024, dtype=f32)
    halfway = (codes[:, None] + np.array([0.5, 0.49999997, 0.50000006], f32)[None, :]).reshape(-1) / f32(1023)
    x = np.concatenate([rng.uniform(-0.25, 1.25, 60000).astype(f32), codes / f32(1023), np.arange(256, dtype=f32) / f32(255), halfway,
                        np.array([0, -0.0, 1, np.inf, -np.inf, np.nan, 1e-30, 65504.0, 70000.0, 1e-8, 5.96e-8, 2.98e-8], f32)])
    x = np.ascontiguousarray(np.resize(x, (x.size // 3) * 3))
    got = np.empty_like(x); want = np.empty_like(x)
    assert HL.host_surface_roundtrip(fmt, _p(x), _p(got), x.size // 3) == 0
    assert L.orc_test_rt_roundtrip(fmt, _p(x), _p(want), x.size) == 0
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (fmt, x[~same][:5], got[~same][:5], want[~same][:5])


@pytest.mark.parametrize("Q", [255.0, 1023.0])
def test_final_pass_quantiser_equals_the_contract(Q):
    HL = host_lib(); L = O.lib()
    HL.host_final_quant.argtypes = [FP, FP, C.c_float, FP, C.c_int]
    L.orc_test_final_quant.argtypes = [FP, FP, C.c_float, FP, C.c_int]
    # every (10-bit code, dither threshold) pair, and random floats
    c = (np.arange(1024, dtype=f32) / f32(1023))[:, None].repeat(1024, 1).reshape(-1)
    d = (np.arange(1024, dtype=f32) / f32(1024))[None, :].repeat(1024, 0).reshape(-1)
    rng = np.random.Generator(np.random.PCG64(3))
    p = np.ascontiguousarray(np.concatenate([c, rng.uniform(-0.1, 1.1, 200000).astype(f32)]))
    dd = np.ascontiguousarray(np.concatenate([d, (rng.integers(0, 1024, 200000).astype(f32) / f32(1024))]))
    got = np.empty_like(p); want = np.empty_like(p)
    assert HL.host_final_quant(_p(p), _p(dd), Q, _p(got), p.size) == 0
    assert L.orc_test_final_quant(_p(p), _p(dd), Q, _p(want), p.size) == 0
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_the_shim_is_test_infrastructure_only():
    """nothing under videorenderer_b200/ (or bench.py, or the C ABI headers) refers to the host shim or to the oracle: the product has no CPU path"""
    import re
    bad = []
    roots = [os.path.join(ROOT, "videorenderer_b200"), os.path.join(ROOT, "include")]
    for r in roots:
        for d, _, files in os.walk(r):
            if os.path.basename(d) in ("build", "__pycache__"):
                continue
            for fn in files:
                if not fn.endswith((".py", ".h", ".cuh", ".cu", ".cpp")):
                    continue
                text = open(os.path.join(d, fn), encoding="utf-8", errors="replace").read()
                if re.search(r"host_shim|cuda_host_shim|liborc|orc_api\.h|orc_core", text):
                    bad.append(os.path.join(d, fn))
    assert not bad, bad


# ---- the scalar arithmetic contract itself: mpcvr_math.cuh against oracle/orc_math.h ----
def _floats_between(lo, hi, stride):
    a = np.array([lo, hi], f32).view(np.uint32)
    return np.arange(int(a[0]), int(a[1]) + 1, stride, dtype=np.uint64).astype(np.uint32).view(f32)


def _math(op, x, y=None):
    HL = host_lib(); L = O.lib()
    HL.host_math.argtypes = [C.c_int, FP, FP, FP, C.c_longlong]
    L.orc_test_math_vec.argtypes = [C.c_int, FP, FP, FP, C.c_int64]
    x = np.ascontiguousarray(x, f32); y = np.ascontiguousarray(y if y is not None else x, f32)
    got = np.empty_like(x); want = np.empty_like(x)
    assert HL.host_math(op, _p(x), _p(y), _p(got), x.size) == 0
    assert L.orc_test_math_vec(min(op, 5), _p(x), _p(y), _p(want), x.size) == 0
    return got, want


def test_lg2_ex2_sin_equal_the_contract_on_dense_sweeps():
    """every 61st float of the ranges the shader chain feeds them with (lg2: all positive normals; ex2: [-150, 130]; sin: [0, 16))"""
    for op, xs in ((0, _floats_between(1.17549435e-38, 3.0e38, 61)), (1, np.concatenate([_floats_between(1e-30, 130.0, 61), -_floats_between(1e-30, 150.0, 61)])),
                   (3, _floats_between(1e-30, 16.0, 61))):
        got, want = _math(op, xs)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (op, int((got.view(np.uint32) != want.view(np.uint32)).sum()))


def test_pow_and_division_equal_the_contract():
    rng = np.random.Generator(np.random.PCG64(21))
    n = 4_000_000
    x = np.concatenate([rng.uniform(0, 1.2, n).astype(f32), (rng.uniform(0, 1, n) ** 8).astype(f32), np.array([0, -1, 1, np.inf, 1e-40, np.nan], f32)])
    ys = np.array([1 / 2.2, 2.2, 2.4, 1 / 78.84375, 1 / 0.1593017578125, 0.1593017578125, 78.84375, 0.2, 1.8, 2.6, 2.8, 0.5, 1.5], f32)
    y = ys[rng.integers(0, ys.size, x.size)]
    got, want = _math(2, x, y)
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (int((~same).sum()), x[~same][:4], y[~same][:4], got[~same][:4], want[~same][:4])
    # div_rn_safe: the reciprocal + residual-correction quotient, on the operand ranges the tone-map quotients have (den in [0.16, 19])
    num = rng.uniform(0, 1000, n).astype(f32); den = rng.uniform(0.16, 19, n).astype(f32)
    got, want = _math(5, num, den)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_unorm_reads_equal_correctly_rounded_division():
    """unorm8 / unorm10 / unorm16: the two-operation exact forms (mpcvr_math.cuh) against code / (2^n - 1) in IEEE arithmetic, every code"""
    for op, top in ((6, 255), (7, 1023), (8, 65535)):
        c = np.arange(top + 1, dtype=f32)
        got, _ = _math(op, c)
        assert np.array_equal(got.view(np.uint32), (c / f32(top)).astype(f32).view(np.uint32)), top
