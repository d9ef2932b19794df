"""The whole frame chain through the reference's EXECUTED shaders.

Pins 8 and 9 hold each shader to the oracle one at a time, in float.  Here they are chained the way CDX11VideoProcessor::Process
(DX11VideoProcessor.cpp:3285-3424) chains its draws -- generated convert shader -> m_TexConvertOutput -> X resize shader -> m_TexResize
(RGBA16F) -> Y resize shader -> post-scale texture -> ps_final_pass with the dither texture -> back buffer -- with the D3D11 render-target
conversions between the draws (float -> UNORM: floor(saturate(x) * (2^n - 1) + 0.5), NaN -> 0; float -> half: round to nearest even; the next
pass reads code / (2^n - 1) or the half), on small frames shaped like the four BASELINE configs.  The codes that reach the back buffer must
EQUAL the oracle's (orc_process, LITERAL mode) byte for byte; the pass planning (internal format, which axes resize, final pass or not) is
taken from the oracle's own decisions, which tests/test_reference_plan_pin.py and test_reference_resize_pin.py hold to the reference's code.
SPEC mode -- what the CUDA path reproduces bit for bit -- is then compared with the same reference-executed frame: the fraction of output
bytes that differ, and by how much, is the measured cost of SPEC's fused mads and fixed exp2 / log2 polynomials in presented codes."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as H
import orclib as O
import test_reference_convert_exec_pin as TC

REF_SH = os.path.join(O.ORC_DIR, "_ref", "libref_hlsl_shaders.so")
pytestmark = pytest.mark.skipif(not (os.path.exists(REF_SH) and os.path.exists(TC.REF)), reason="oracle/_ref shader libraries not built")

FP = C.POINTER(C.c_float)
f32 = np.float32
INTERP = {O.UPSCALE["Mitchell"]: "mitchell4", O.UPSCALE["CatmullRom"]: "catmull4", O.UPSCALE["Lanczos2"]: "lanczos2", O.UPSCALE["Lanczos3"]: "lanczos3"}
CONV = {O.DOWNSCALE["Box"]: "box", O.DOWNSCALE["Bilinear"]: "bilinear", O.DOWNSCALE["Hamming"]: "hamming", O.DOWNSCALE["Bicubic"]: "bicubic05",
        O.DOWNSCALE["BicubicSharp"]: "bicubic15", O.DOWNSCALE["Lanczos"]: "lanczos"}


def _p(a):
    return a.ctypes.data_as(FP) if a is not None else None


def rt_roundtrip(x, fmt):
    """render-target write of float RGBA `x`, then what the next draw samples: (stored codes or halves, float texels)"""
    if fmt == O.SURF_RGBA16F:
        h = x.astype(np.float16)
        return h, np.ascontiguousarray(h.astype(f32))
    N = f32(255 if fmt == O.SURF_BGRA8 else 1023)
    v = np.clip(np.where(np.isnan(x), f32(0), x), f32(0), f32(1)).astype(f32)
    code = np.floor((v * N).astype(f32) + f32(0.5)).astype(f32)
    return code.astype(np.int32), np.ascontiguousarray((code / N).astype(f32))


def _shader_lib():
    R = C.CDLL(REF_SH)
    R.ref_hlsl_run.argtypes = [C.c_char_p, FP, C.c_int, C.c_int, C.c_int, FP, C.c_int, C.c_int, FP, C.c_int, C.c_float, C.c_float, FP]
    return R


def _draw(R, name, tex, ow, oh, tex2=None, sx=None, sy=None):
    """one full-rect draw of `tex` (h, w, 4) onto ow x oh pixels with shader `name`: TEXCOORD = (i + .5) * (size / dst) / size"""
    h, w = tex.shape[:2]
    gx = f32(w) / f32(ow); gy = f32(h) / f32(oh)
    cx = (np.arange(ow, dtype=f32) + f32(.5)) * gx; cy = (np.arange(oh, dtype=f32) + f32(.5)) * gy
    uv = np.empty((oh, ow, 2), f32); uv[..., 0] = (cx / f32(w))[None, :]; uv[..., 1] = (cy / f32(h))[:, None]
    out = np.empty((oh, ow, 4), f32)
    t2 = (_p(tex2), tex2.shape[1], tex2.shape[0]) if tex2 is not None else (None, 0, 0)
    assert R.ref_hlsl_run(name.encode(), _p(tex), w, h, 0, *t2, _p(uv), ow * oh, gx if sx is None else sx, gy if sy is None else sy, _p(out)) == 0, name
    return out


# (case of the convert-shader table, destination size, settings): small frames shaped like BASELINE C2, C3, C4, C1 and two mixed ones
CHAINS = {
    "c2_p010_pq_lanczos3_2x_dither_bgra8": ("P010_pq_to_sdr", 128, 64, dict(up=O.UPSCALE["Lanczos3"], to_sdr=1, backbuffer=O.SURF_BGRA8)),
    "c3_yuy2_hlg_catmull_1p5x_rgb10_16f":  ("YUY2_hlg_to_sdr", 96, 48, dict(up=O.UPSCALE["CatmullRom"], to_sdr=1, backbuffer=O.SURF_RGB10A2, tex=O.TEXFMT_16FLOAT,
                                                                          chroma=O.CHROMA_CatmullRom)),
    "c4_p016_2020_bicubic_half_bgra8":     ("P016_2020_sdr", 32, 16, dict(down=O.DOWNSCALE["Bicubic"], interp50=0, to_sdr=0, backbuffer=O.SURF_BGRA8)),
    "half_size_takes_the_upscaler":        ("P016_2020_sdr", 32, 16, dict(up=O.UPSCALE["Lanczos3"], down=O.DOWNSCALE["Bicubic"], interp50=1, to_sdr=0, backbuffer=O.SURF_BGRA8)),
    "c1_nv12_709_copy_bgra8":              ("NV12_s5_c1", 64, 32, dict(to_sdr=0, backbuffer=O.SURF_BGRA8)),
    "x_up_y_down_mitchell_hamming_rgb10":  ("P010_pq_to_sdr", 128, 8, dict(up=O.UPSCALE["Mitchell"], down=O.DOWNSCALE["Hamming"], to_sdr=1, backbuffer=O.SURF_RGB10A2)),
    "x_down_lanczos_y_up_lanczos3_bgra8":  ("P010_pq_to_sdr", 24, 64, dict(up=O.UPSCALE["Lanczos3"], down=O.DOWNSCALE["Lanczos"], to_sdr=1, backbuffer=O.SURF_BGRA8)),
    "c2_shaped_1024x512_to_2048x1024_natural": ("P010_pq_to_sdr_1024x512", 2048, 1024, dict(up=O.UPSCALE["Lanczos3"], to_sdr=1, backbuffer=O.SURF_BGRA8, kind="natural")),
    "lanczos2_no_dither_8int":             ("NV12_s5_c1", 128, 64, dict(up=O.UPSCALE["Lanczos2"], to_sdr=0, backbuffer=O.SURF_BGRA8, dither=0, tex=O.TEXFMT_8INT)),
}


@pytest.mark.parametrize("chain", sorted(CHAINS))
def test_frame_through_the_executed_reference_shaders_equals_the_oracle(chain):
    case_name, ow, oh, kw = CHAINS[chain]
    kw = dict(kw); kind = kw.pop("kind", "fullnoise")
    i = [k for k, n in TC.CASES if n == case_name][0]
    RC, _, _ = TC._libs()
    info = (C.c_uint * 9)(); RC.ref_convert_case(i, info)
    cf, ex, siting, scaler, conv, blend, dovi, W, Hh = [int(x) for x in info]
    assert kw.get("chroma", O.CHROMA_Bilinear) == scaler and int(kw["to_sdr"]) == int(conv == 1)
    # 1. the convert draw: the reference's generated shader on the frame run_case builds (full-range noise, seed = case index)
    conv_ref = TC.run_case(i, kind=kind, seed=i)[4]
    conv_rgba = np.concatenate([conv_ref, np.ones(conv_ref.shape[:2] + (1,), f32)], axis=2)
    # the oracle on the same planes, with this chain's settings
    codes = H.make_codes(cf, W, Hh, kind, seed=i)
    _, _, planes, pitches = H.pack(cf, W, Hh, codes)
    cfg = O.make_config(cf, W, Hh, ow, oh, ex, **kw)
    out_lit, dbg = O.process(cfg, planes, pitches, mode=O.MODE_LITERAL, want_debug=True)
    out_spec = O.process(cfg, planes, pitches, mode=O.MODE_SPEC)
    fmt_int, final, rx, ry = dbg["internal_fmt"], dbg["final_pass"], dbg["resize_x"], dbg["resize_y"]
    bb = kw["backbuffer"]
    R = _shader_lib()
    # 2. m_TexConvertOutput
    _, tex = rt_roundtrip(conv_rgba, fmt_int)
    target = fmt_int if final else bb                                     # what the last resize (or the copy) draws into
    def scaler_name(kind, axis):
        return ("%s_%s" % (INTERP[cfg.iUpscaling], "xy"[axis])) if kind == 1 else ("conv_%s_%s" % (CONV[cfg.iDownscaling], "xy"[axis]))
    # 3. ResizeShaderPass (:3103-3187): X into m_TexResize (RGBA16F, window width x source height), then Y; one draw if only one axis resizes
    if rx and ry:
        x = _draw(R, scaler_name(rx, 0), tex, ow, Hh)
        _, tex = rt_roundtrip(x, O.SURF_RGBA16F)
        y = _draw(R, scaler_name(ry, 1), tex, ow, oh)
    elif rx or ry:
        y = _draw(R, scaler_name(rx or ry, 0 if rx else 1), tex, ow, oh)
    else:
        y = tex                                                              # ps_simple copy of equal-sized rects
    stored, tex = rt_roundtrip(y, target)
    # 4. FinalPass (:3189-3233): ps_final_pass for the back buffer's depth, dither texture addressed by Tex * (size / 32), WRAP
    if final:
        O.lib().orc_dither_thresholds.restype = C.POINTER(C.c_uint16)
        K = np.ctypeslib.as_array(O.lib().orc_dither_thresholds(), (1024,)).astype(f32) / f32(1024)
        dither = np.ascontiguousarray(np.repeat(K.reshape(32, 32, 1), 4, axis=2))
        fin = _draw(R, "final_8" if bb == O.SURF_BGRA8 else "final_10", tex, ow, oh, tex2=dither, sx=f32(ow) / f32(32), sy=f32(oh) / f32(32))
        stored, _ = rt_roundtrip(fin, bb)
    want = np.asarray(stored)[..., :3]
    got = H.rgb_view(out_lit, bb)
    assert np.array_equal(got, want), (chain, int((got != want).sum()), fmt_int, final, rx, ry)
    # SPEC against the reference-executed frame, in presented codes
    spec = H.rgb_view(out_spec, bb)
    d = np.abs(spec.astype(np.int64) - want.astype(np.int64))
    print("CHAIN %-40s internal fmt %d final %d resize (%d, %d)  SPEC vs executed reference: %d of %d bytes differ (%.3f %%), max |delta| %d"
          % (chain, fmt_int, final, rx, ry, int((d != 0).sum()), d.size, 100.0 * float((d != 0).mean()), int(d.max())))
    assert d.max() <= 1 and (d != 0).mean() <= 0.01, (chain, int(d.max()), float((d != 0).mean()))


def test_the_chain_bench_times_renders_the_oracles_frame():
    """bench.py's `--impl reference` / cpu_reference_shaders leg runs the same executed shaders through whole-draw C entry points (OpenMP over
    pixels, the target conversions in C): the frame it renders must be the oracle's LITERAL frame, byte for byte -- what is timed is the path"""
    import bench
    r = bench.reference_shader_chain(threads=4)
    assert r is not None
    step, info = r
    step()
    cf, W, Hh, ow, oh, ex, planes, pitches = info["oracle_case"]
    cfg = O.make_config(cf, W, Hh, ow, oh, ex, up=O.UPSCALE["Lanczos3"], to_sdr=1, backbuffer=O.SURF_BGRA8)
    want = H.rgb_view(O.process(cfg, planes, pitches, mode=O.MODE_LITERAL), O.SURF_BGRA8)
    got = np.rint(info["final_rgba"][..., :3] * f32(255)).astype(np.int32)
    assert np.array_equal(got, want), int((got != want).sum())
