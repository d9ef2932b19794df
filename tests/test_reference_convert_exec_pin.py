"""Ninth pin: the convert shader the reference GENERATES, executed -- the whole ConvertColorPass pixel function, end to end.

CDX11VideoProcessor does not ship its convert shader: GetShaderConvertColor (Source/Shaders.cpp:593-930) writes HLSL text per media type --
ShaderGetPixels' fetch branch for the format / chroma siting / chroma scaler / blend-deinterlace, the Dolby Vision reshaping, the colour
matrix, and the transfer steps spliced from Shaders/convert/{st2084,hlg,hdr_tone_mapping}.hlsl -- and hands it to D3DCompile.  That
generator is compiled where it lies (oracle/_ref/libref_shaders.so, CompileShader stubbed to keep the text).  oracle/Makefile asks it for
the text of 117 cases (oracle/ref_stub/gen_convert_shaders.py: every planar / bi-planar / packed layout x chroma scaler x siting,
blend deinterlace, PQ and HLG to SDR, HLG to PQ, pass-through, the BT.2020 SDR gammas, RGB and gray sources, Dolby Vision with polynomial
and MMR curves) and compiles each text AS C++ over an HLSL vector / swizzle / Texture2D vocabulary (oracle/ref_stub/hlsl_vec.h) into
oracle/_ref/libref_hlsl_convert.so; the edits to the text are syntax only and listed in the generator.  Nothing is copied into the repo.

Here every case's shader runs over a whole frame of full-range random codes, fed with
  * the textures the D3D11 UNORM views of the frame's planes return (built below from the same plane memory the oracle reads: DXGI formats
    per Source/Helper.cpp:311-359; this, the point / bilinear sampler in hlsl_vec.h and the texcoord = texel centre / size of a
    full-rect draw are the test's model of D3D11, stated once and shared by nothing else),
  * the PS_COLOR_TRANSFORM constants of the reference's own SetShaderConvertColorParams (ref_convert_color_params, compiled in place),
  * for Dolby Vision the curve / trim constant buffers of the reference's own SetShaderDoviCurves and CopySample level block
    (oracle/_ref/libref_dovi.so),
and the result is compared with the convert-pass output of the oracle (orc_process, debug tap convert_f32 = the pixel before the render
target write) in LITERAL mode -- the mode that claims to be the shader text, one IEEE operation per written operation: BIT-EQUAL, all
117 cases.  SPEC mode (what the CUDA kernels are bit-exact with) is held to a measured distance from the executed text.

So the fetch arithmetic (chroma positions, Catmull-Rom weights and tap order, 4:2:2 interpolation, the packed-format swizzles, the blend
deinterlace), the order of matrix / saturate / transfer / gamut steps, the Dolby Vision reshaping (pivot search, polynomial, MMR orders)
and trims are no longer only restated from reading the reference: they are checked against the reference's own text run on the same bytes."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as H
import orclib as O

REF = os.path.join(O.ORC_DIR, "_ref", "libref_hlsl_convert.so")
REF_S = os.path.join(O.ORC_DIR, "_ref", "libref_shaders.so")
REF_D = os.path.join(O.ORC_DIR, "_ref", "libref_dovi.so")
pytestmark = pytest.mark.skipif(not all(os.path.exists(p) for p in (REF, REF_S, REF_D)),
                                reason="oracle/_ref/libref_hlsl_convert.so not built (reference tree absent at build time)")

FP = C.POINTER(C.c_float)
f32 = np.float32
PIV = (C.c_float * 7) * 3
COEF = ((C.c_float * 4) * 8) * 3
MMR = ((C.c_float * 4) * 48) * 3
PAR = (C.c_int32 * 4) * 3
_L = None


def _libs():
    global _L
    if _L is None:
        RC, RS, RD = C.CDLL(REF), C.CDLL(REF_S), C.CDLL(REF_D)
        RC.ref_convert_case_count.restype = C.c_int
        RC.ref_convert_case.restype = C.c_char_p; RC.ref_convert_case.argtypes = [C.c_int, C.c_uint * 9]
        RC.ref_convert_dovi_lms.argtypes = [C.c_double * 9]
        RC.ref_convert_run.argtypes = [C.c_int, FP * 3, C.c_int * 3, C.c_int * 3, C.c_float * 12, C.c_float, FP, FP, FP, C.POINTER(C.c_int32), FP, C.c_int, FP]
        RC.ref_convert_run_rect.argtypes = [C.c_int, FP * 3, C.c_int * 3, C.c_int * 3, C.c_float * 12, C.c_float, FP, FP, FP, C.POINTER(C.c_int32), FP, C.c_int,
                                            C.c_int * 4, FP]
        RS.ref_convert_color_params.argtypes = [C.c_int, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                                (C.c_float * 3) * 3, C.c_float * 3, C.POINTER(C.c_int32)]
        RD.ref_dovi_curves.argtypes = [C.POINTER(O.OrcDovi), PIV, COEF, MMR, PAR]
        RD.ref_dovi_dynamic.argtypes = [C.POINTER(O.OrcDovi), C.c_int, C.c_int, C.c_int, C.c_float * 3, C.POINTER(C.c_int32), C.c_float * 5, C.POINTER(C.c_int32)]
        _L = (RC, RS, RD)
    return _L


def _cases():
    if not os.path.exists(REF):
        return []
    RC = C.CDLL(REF)
    RC.ref_convert_case.restype = C.c_char_p; RC.ref_convert_case.argtypes = [C.c_int, C.c_uint * 9]
    out = []
    for i in range(RC.ref_convert_case_count()):
        info = (C.c_uint * 9)()
        out.append((i, RC.ref_convert_case(i, info).decode()))
    return out


CASES = _cases()


def _p(a):
    return a.ctypes.data_as(FP)


def _rgba(r, g=None, b=None, a=None):
    t = np.zeros(r.shape + (4,), f32); t[..., 0] = r
    if g is not None: t[..., 1] = g
    if b is not None: t[..., 2] = b
    t[..., 3] = 1 if a is None else a                       # components a format lacks read 0, 0, 1
    return np.ascontiguousarray(t)


def textures(cf, w, h, planes, pitches):
    """The float4 texels registers t0.. return for the frame: UNORM views (code / (2^n - 1), correctly rounded) of the planes in memory
    order -- R8 / R16 (+ R8G8 / R16G16 chroma) for planar and bi-planar layouts, R8G8B8A8 / R16G16B16A16 of half the width for packed
    4:2:2, R8G8B8A8 / R10G10B10A2 / R16G16B16A16 for packed 4:4:4 and RGB (B8G8R8X8 for 8-bit RGB), R8 / R16 for gray."""
    f = H.fmt(cf)
    dt = np.uint8 if f.bytes == 1 else np.uint16
    maxc = f32(255 if f.bytes == 1 else 65535)

    def words(k, rows, comps):
        a = np.frombuffer(planes[k], np.uint8)[:rows * pitches[k]].reshape(rows, pitches[k])[:, :comps * f.bytes]
        a = np.ascontiguousarray(a).view(dt).astype(np.uint32)
        if f.bytes == 2 and f.lsb_shift:
            a = (a << f.lsb_shift) & 0xFFFF              # CopyPlane10to16 (Source/Helper.cpp:541-565) stores the code MSB-aligned
        return a

    def U(a):
        return a.astype(f32) / maxc
    cw, ch = w // f.div_w, h // f.div_h
    if f.layout == 1:
        uv = words(1, ch, cw * 2).reshape(ch, cw, 2)
        return [_rgba(U(words(0, h, w))), _rgba(U(uv[..., 0]), U(uv[..., 1]))]
    if f.layout == 2:
        return [_rgba(U(words(0, h, w))), _rgba(U(words(1, ch, cw))), _rgba(U(words(2, ch, cw)))]
    if f.layout == 3:
        a = words(0, h, (w // 2) * 4).reshape(h, w // 2, 4)
        return [_rgba(U(a[..., 0]), U(a[..., 1]), U(a[..., 2]), U(a[..., 3]))]
    if f.layout in (4, 5):
        raw = np.frombuffer(planes[0], np.uint8)[:h * pitches[0]].reshape(h, pitches[0])
        if f.packed_kind == 1:                                           # R10G10B10A2
            v = np.ascontiguousarray(raw[:, :w * 4]).view(np.uint32)
            q = lambda s: ((v >> s) & 1023).astype(f32) / f32(1023)
            return [_rgba(q(0), q(10), q(20), ((v >> 30) & 3).astype(f32) / f32(3))]
        if f.packed_kind == 0:
            b = raw[:, :w * 4].reshape(h, w, 4).astype(f32) / f32(255)
            if f.layout == 4:
                return [_rgba(b[..., 0], b[..., 1], b[..., 2], b[..., 3])]    # AYUV bytes as R8G8B8A8
            return [_rgba(b[..., 2], b[..., 1], b[..., 0], b[..., 3])]        # B8G8R8X8: .r is byte 2
        a = np.ascontiguousarray(raw[:, :w * 8]).view(np.uint16).reshape(h, w, 4).astype(f32) / f32(65535)
        return [_rgba(a[..., 0], a[..., 1], a[..., 2], a[..., 3])]
    return [_rgba(U(words(0, h, w)))]


DEFAULT_L2 = [(2851, 2200, 1900, 2100, 2000, 2150), (3300, 1950, 2100, 2000, 2120, 1990)]


def run_case(i, l2=DEFAULT_L2, src_rect=None, kind="fullnoise", seed=None):
    RC, RS, RD = _libs()
    info = (C.c_uint * 9)()
    name = RC.ref_convert_case(i, info).decode()
    cf, ex, siting, scaler, conv, blend, dovi, W, Hh = [int(x) for x in info]
    dovi = dovi if dovi < 2 else -1                                       # (unsigned)-1
    f = H.fmt(cf)
    codes = H.make_codes(cf, W, Hh, kind, seed=i if seed is None else seed)
    _, _, planes, pitches = H.pack(cf, W, Hh, codes)
    # RGB sources only get a convert pass with a non-neutral ProcAmp (m_PSConvColorData.bEnable, DX11VideoProcessor.cpp:846-850)
    pa = dict(brightness=10.0, contrast=1.25) if f.cstype == 1 else {}
    x0, y0, x1, y1 = src_rect or (0, 0, W, Hh)
    cfg = O.make_config(cf, W, Hh, x1 - x0, y1 - y0, ex, chroma=scaler, to_sdr=1 if conv == 1 else 0, deint_blend=blend,
                        hdr={"out": 1} if conv == 2 else None, src_rect=src_rect, **pa)
    dv = None; dvargs = [None] * 5 + [0]; dm = doff = None; keep = None
    if dovi >= 0:
        # two Level-2 trim sets bracketing the 1000-nit display (PQ code 3079): the reference interpolates between them
        dv = H.make_dovi(seed=3 + i, mmr=dovi * 2, l2=l2 if conv == 1 else None)
        lms = (C.c_double * 9)(); RC.ref_convert_dovi_lms(lms)             # the matrix the case's text embeds
        for k in range(9):
            dv.rgb_to_lms_matrix[k] = lms[k]
        piv, coef, mmr, par = PIV(), COEF(), MMR(), PAR()
        assert RD.ref_dovi_curves(C.byref(dv), piv, coef, mmr, par) == 0
        l1, pp, tr, enb = (C.c_float * 3)(), C.c_int32(), (C.c_float * 5)(), C.c_int32()
        assert RD.ref_dovi_dynamic(C.byref(dv), 1000, 1, 1, l1, C.byref(pp), tr, C.byref(enb)) == 0
        assert enb.value == int(conv == 1)
        keep = [np.frombuffer(bytes(x), f32).copy() for x in (piv, coef, mmr)] + [np.frombuffer(bytes(par), np.int32).copy(), np.frombuffer(bytes(tr), f32).copy()]
        dvargs = [_p(keep[0]), _p(keep[1]), _p(keep[2]), keep[3].ctypes.data_as(C.POINTER(C.c_int32)), _p(keep[4]), enb.value]
        dm = (C.c_double * 9)(*dv.ycc_to_rgb_matrix); doff = (C.c_double * 3)(*dv.ycc_to_rgb_offset)
    lit = O.process(cfg, planes, pitches, mode=O.MODE_LITERAL, want_debug=True, dovi=dv)[1]["convert_f32"]
    spec = O.process(cfg, planes, pitches, mode=O.MODE_SPEC, want_debug=True, dovi=dv)[1]["convert_f32"]
    tex = textures(cf, W, Hh, planes, pitches)
    while len(tex) < 3:
        tex.append(tex[0])
    cm, cc, en = ((C.c_float * 3) * 3)(), (C.c_float * 3)(), C.c_int32()
    assert RS.ref_convert_color_params(cf, ex, pa.get("brightness", 0.0), pa.get("contrast", 1.0), 0.0, 1.0, dm, doff, cm, cc, C.byref(en)) == 0
    assert en.value == 1, name
    cm12 = (C.c_float * 12)(*([cm[r][c] for r in range(3) for c in range(3)] + list(cc)))
    ref = np.empty((y1 - y0, x1 - x0, 4), f32)
    assert RC.ref_convert_run_rect(i, (FP * 3)(*[_p(t) for t in tex]), (C.c_int * 3)(*[t.shape[1] for t in tex]), (C.c_int * 3)(*[t.shape[0] for t in tex]),
                                   cm12, f32(10000.0) / f32(125.0), *dvargs, (C.c_int * 4)(x0, y0, x1, y1), _p(ref)) == 0
    return name, conv, dovi, int(info[1]), np.ascontiguousarray(ref[..., :3]), lit, spec


@pytest.mark.parametrize("i,name", CASES, ids=[n for _, n in CASES])
def test_generated_convert_shader_equals_the_oracle(i, name):
    name, conv, dovi, ex, ref, lit, spec = run_case(i)
    bad = np.nonzero(ref.view(np.uint32) != lit.view(np.uint32))
    assert bad[0].size == 0, (name, bad[0].size, ref[bad][:4], lit[bad][:4])
    # SPEC (the CUDA kernels' contract) against the executed reference text.  Without transfer steps it is the colour matrix's fused mads:
    # a few ULP at 1.0.  With pow / exp2 / log2 steps the fixed polynomials and the fused mads move the result by up to ~1e-4 of full
    # scale, and more only where a channel sits within rounding of the clamp at 0 before the final 1/2.2 power (x = 1e-6 -> 2e-3).
    err = np.abs(ref.astype(np.float64) - spec)
    trc = (ex >> 27) & 31; prim = (ex >> 22) & 31
    plain = dovi < 0 and conv != 1 and not (prim == 9 and conv == 0 and trc != 15) and not (trc == 16 and conv == 2)
    if plain:
        assert err.max() <= 6e-7, (name, err.max())
    else:
        assert np.percentile(err, 99.9) <= 3e-4 and err.max() <= 5e-3, (name, np.percentile(err, 99.9), err.max())


@pytest.mark.parametrize("kind", ["natural", "noise", "ramp"])
def test_generated_convert_shader_on_other_content(kind):
    """legal-range noise, smooth "natural" frames and ramps (the per-case test uses full-range noise): all 117 cases again"""
    for i, name in CASES:
        name, conv, dovi, ex, ref, lit, spec = run_case(i, kind=kind, seed=1000 + i)
        bad = np.nonzero(ref.view(np.uint32) != lit.view(np.uint32))
        assert bad[0].size == 0, (name, kind, bad[0].size)


@pytest.mark.parametrize("rect", [(3, 5, 61, 30), (1, 0, 64, 31), (2, 2, 34, 18)], ids=["odd_offsets", "odd_x", "even"])
def test_generated_convert_shader_on_a_cropped_source(rect):
    """m_srcRect with odd offsets moves every chroma phase (frac(Tex * wh / 2), the fmod(Tex.x * w, 2) test of the 4:2:2 branches, the blend
    deinterlace rows): the draw's texcoords start inside the texture and the neighbours it fetches are real texels, not clamped ones"""
    n = 0
    for i, name in CASES:
        if "dovi" in name or "trc" in name or name.split("_")[0] in ("XRGB32", "RGB48", "Y8", "Y16"):
            continue
        x0, y0, x1, y1 = rect
        name, conv, dovi, ex, ref, lit, spec = run_case(i, src_rect=rect)
        assert ref.shape == lit.shape
        bad = np.nonzero(ref.view(np.uint32) != lit.view(np.uint32))
        assert bad[0].size == 0, (name, rect, bad[0].size)
        n += 1
    assert n >= 70


# (target_max_pq, slope, offset, power, chroma_weight, saturation_gain) raw 12-bit fields; the display is 1000 nits = PQ code 3079
TRIM_SETS = {
    "negative_offset":   [(3079, 2048, 1200, 2048, 2048, 2300)],       # offset -0.2: dark channels go below 0 -> pow of a negative base
    "negative_satgain":  [(3079, 2048, 2040, 2048, 2048, 1000)],       # saturation gain < 0 (a channel AT 0 would make 0 * inf; PQ(0) > 0 keeps them off it)
    "zero_satgain":      [(3079, 2100, 2000, 2048, 2200, 2048)],       # saturation gain = 0: pow(0, 0) = exp2(0 * -inf)
    "bright":            [(3079, 4095, 4095, 4095, 2048, 2048)],       # slope 1.5, offset +0.5, power 1.5: values past ST2084ToLinear's pole (~1.99)
    "bright_sat":        [(3079, 4095, 3500, 4095, 4095, 4095)],
    "between_two":       DEFAULT_L2,
    "dimmer_than_all":   [(3500, 1800, 1700, 2300, 2500, 1500), (3800, 2500, 2400, 1800, 1500, 2600)],
    "brighter_than_all": [(2081, 1800, 2500, 2300, 2500, 1500)],
}


@pytest.mark.parametrize("trims", sorted(TRIM_SETS))
def test_dolby_vision_trims_equal_the_generated_shader_including_its_nan_pixels(trims):
    """DolbyVisionTrims (Shaders.cpp:787-795) has no guard: a negative pow base, a zero luminance, a zero channel under a non-positive
    saturation gain, or a value past the pole of ST2084ToLinear make NaN, which the gamut matrix spreads over the pixel and the final
    saturate turns into (0, 0, 0).  The executed text shows it; the oracle (dovi_trims' poisoned flag) and with it the CUDA kernel
    (pixel_ops.cuh convert_color) must produce the same black pixels -- bit-equal on whole frames of full-range noise, for trim sets that
    provoke each cause and for the three ways the reference picks / interpolates Level-2 blocks."""
    seen_black = seen_live = 0
    for i, name in CASES:
        if not name.endswith("_to_sdr") or "dovi" not in name:
            continue
        name, conv, dovi, ex, ref, lit, spec = run_case(i, l2=TRIM_SETS[trims])
        bad = np.nonzero(ref.view(np.uint32) != lit.view(np.uint32))
        assert bad[0].size == 0, (name, trims, bad[0].size, ref[bad][:6], lit[bad][:6])
        black = np.all(ref == 0, axis=2)
        assert np.array_equal(black, np.all(spec == 0, axis=2)), (name, trims)          # SPEC poisons exactly the same pixels
        seen_black += int(black.sum()); seen_live += int((~black).sum())
    assert seen_live > 0
    if trims in ("negative_offset", "bright", "bright_sat"):
        assert seen_black > 0, trims


def test_case_table_covers_the_fetch_and_transfer_branches():
    names = [n for _, n in CASES]
    assert len(names) >= 110
    for key in ("NV12_s1_c2", "YV12_s7_c1", "YUY2_c2", "UYVY_c1", "Y210_c0", "P210_c2", "YUV422P16_c2", "AYUV", "Y410", "Y416", "XRGB32", "RGB48", "Y8",
                "NV12_blend_catmull", "P010_pq_to_sdr", "P010_hlg_to_sdr", "P010_hlg_to_pq", "P010_2020_trc8", "P010_dovi0_to_sdr", "P010_dovi1_to_pq"):
        assert key in names, key
