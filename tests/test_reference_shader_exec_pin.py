"""Eighth pin: the reference's resize and final-pass PIXEL SHADERS, executed.

Shaders/d3d11/ps_interpolation_spline4.hlsl (Mitchell, Catmull-Rom), ps_interpolation_lanczos2.hlsl, ps_interpolation_lanczos3.hlsl,
ps_convolution.hlsl with each filter of Shaders/resize/convolution_filters.hlsl (box, bilinear, Hamming, bicubic A=-0.5 / -1.5, Lanczos3),
both AXIS builds of each, and ps_final_pass.hlsl for QUANTIZATION 255 / 1023 -- the 22 shader builds the reference's resize and final stages
choose from -- are assembled where they lie into one C++ unit (oracle/ref_stub/gen_hlsl_shaders.py applies syntax edits only: register /
semantic annotations off, cbuffer -> namespace, [loop] off, file-scope `static const` -> macro, the #include spliced in; the arithmetic
lines reach g++ as written) over a Texture2D / point SamplerState model (oracle/ref_stub/hlsl_compat_shaders.h), compiled with
-fsingle-precision-constant -ffp-contract=off into oracle/_ref/libref_hlsl_shaders.so.  Nothing is copied into the repo.

Each shader then draws whole rectangles of random texels (offset source rects, rects that touch the texture border so CLAMP addressing
takes part, non-integer ratios) and the oracle's LITERAL mode draws the same rectangles through resize_pass / final_pass
(oracle/orc_core.inc) on RAW surfaces.  LITERAL claims to be the shader text in written order, one IEEE operation per operation, libm sin /
cos: so the float4 results must be BIT-EQUAL, the D3D11 lanczos3 shader's Q1 := Q0 quirk (ps_interpolation_lanczos3.hlsl:33-34) included.
SPEC mode -- the contract the CUDA kernels are bit-exact with (tests/test_gpu_parity.py) -- differs from LITERAL by fused mads, folded
literal quotients and the fixed sin / cos polynomials; its distance from the executed reference text is asserted too.

Texture sizes are powers of two so that texcoord = texel coordinate / size and texcoord * wh are exact: the test then feeds the shaders the
interpolated TEXCOORD an ideal rasteriser produces (origin + (i + .5) * srcLen / dstLen), which is what the oracle states."""
import ctypes as C
import os

import numpy as np
import pytest

import orclib as O

REF = os.path.join(O.ORC_DIR, "_ref", "libref_hlsl_shaders.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_hlsl_shaders.so not built (reference tree absent at build time)")

FP = C.POINTER(C.c_float)
f32 = np.float32
U, D = O.UPSCALE, O.DOWNSCALE


def _libs():
    R = C.CDLL(REF); L = O.lib()
    R.ref_hlsl_run.argtypes = [C.c_char_p, FP, C.c_int, C.c_int, C.c_int, FP, C.c_int, C.c_int, FP, C.c_int, C.c_float, C.c_float, FP]
    R.ref_hlsl_shader_count.restype = C.c_int
    R.ref_hlsl_shader_name.restype = C.c_char_p; R.ref_hlsl_shader_name.argtypes = [C.c_int]
    L.orc_test_resize_draw.argtypes = [C.c_int] * 5 + [FP, C.c_int, C.c_int, C.c_int * 4, C.c_int, C.c_int, FP]
    L.orc_test_final_pass.argtypes = [C.c_int, FP, C.c_int, C.c_int, C.c_float, FP]
    L.orc_dither_thresholds.restype = C.POINTER(C.c_uint16)
    return R, L


def _p(a):
    return a.ctypes.data_as(FP)


def _draw(R, L, name, axis, kind, filt, W, H, rect, ow, oh, seed):
    """-> (reference shader's RGB, oracle LITERAL RGB, oracle SPEC RGB) for one draw of in-rect `rect` of a W x H texture to ow x oh pixels"""
    rng = np.random.Generator(np.random.PCG64(seed))
    tex = rng.uniform(0, 1, (H, W, 4)).astype(f32)
    l, t, r, b = rect
    gx = f32(r - l) / f32(ow); gy = f32(b - t) / f32(oh)                       # TextureResizeShader :351-354: scale = src size / dst size
    cx = (np.arange(ow, dtype=f32) + f32(.5)) * gx + f32(l)
    cy = (np.arange(oh, dtype=f32) + f32(.5)) * gy + f32(t)
    uv = np.empty((oh, ow, 2), f32); uv[..., 0] = (cx / f32(W))[None, :]; uv[..., 1] = (cy / f32(H))[:, None]
    ref = np.empty((oh, ow, 4), f32)
    assert R.ref_hlsl_run(name.encode(), _p(tex), W, H, 0, None, 0, 0, _p(uv), ow * oh, gx, gy, _p(ref)) == 0
    rgb = np.ascontiguousarray(tex[..., :3])
    outs = []
    for mode in (1, 0):
        o = np.empty((oh, ow, 3), f32)
        assert L.orc_test_resize_draw(mode, axis, kind, filt, O.VARIANT_D3D11, _p(rgb), W, H, (C.c_int * 4)(*rect), ow, oh, _p(o)) == 0
        outs.append(o)
    return np.ascontiguousarray(ref[..., :3]), outs[0], outs[1]


INTERP = [("mitchell4", U["Mitchell"]), ("catmull4", U["CatmullRom"]), ("lanczos2", U["Lanczos2"]), ("lanczos3", U["Lanczos3"])]
CONV = [("box", D["Box"]), ("bilinear", D["Bilinear"]), ("hamming", D["Hamming"]), ("bicubic05", D["Bicubic"]), ("bicubic15", D["BicubicSharp"]),
        ("lanczos", D["Lanczos"])]
# (W, H, in rect, out w, out h): the filtered axis is stretched / shrunk, the other one is point sampled (sizes equal or not)
UP_X = [(64, 32, (0, 0, 64, 32), 128, 32), (64, 32, (3, 2, 61, 30), 97, 28), (64, 32, (0, 0, 64, 32), 173, 32), (32, 16, (5, 1, 30, 16), 25, 15),
        (128, 8, (0, 0, 128, 8), 640, 8)]
DN_X = [(128, 32, (0, 0, 128, 32), 64, 32), (128, 32, (3, 2, 125, 30), 37, 28), (256, 8, (0, 0, 256, 8), 61, 8), (128, 16, (1, 0, 128, 16), 100, 16),
        (512, 4, (0, 0, 512, 4), 43, 4)]
SPEC_ABS_TOL = 4 * 2.0 ** -23        # SPEC vs the executed text, texels in [0, 1]: measured <= 1.2e-7 (1 ULP at 1.0) on these draws; asserted at 4 ULP


def _transpose(g):
    W, H, (l, t, r, b), ow, oh = g
    return (H, W, (t, l, b, r), oh, ow)


@pytest.mark.parametrize("name,filt", INTERP, ids=[n for n, _ in INTERP])
@pytest.mark.parametrize("axis", [0, 1], ids=["x", "y"])
def test_interpolation_shaders_equal_the_reference_hlsl(name, filt, axis):
    R, L = _libs()
    worst = 0.0
    for k, g in enumerate(UP_X):
        W, H, rect, ow, oh = g if axis == 0 else _transpose(g)
        ref, lit, spec = _draw(R, L, "%s_%s" % (name, "xy"[axis]), axis, 1, filt, W, H, rect, ow, oh, seed=10 + k)
        assert np.array_equal(ref.view(np.uint32), lit.view(np.uint32)), (name, axis, g, int((ref.view(np.uint32) != lit.view(np.uint32)).sum()))
        worst = max(worst, float(np.abs(ref.astype(np.float64) - spec).max()))
    assert worst <= SPEC_ABS_TOL, worst


@pytest.mark.parametrize("name,filt", CONV, ids=[n for n, _ in CONV])
@pytest.mark.parametrize("axis", [0, 1], ids=["x", "y"])
def test_convolution_shaders_equal_the_reference_hlsl(name, filt, axis):
    R, L = _libs()
    worst = 0.0
    for k, g in enumerate(DN_X):
        W, H, rect, ow, oh = g if axis == 0 else _transpose(g)
        ref, lit, spec = _draw(R, L, "conv_%s_%s" % (name, "xy"[axis]), axis, 2, filt, W, H, rect, ow, oh, seed=40 + k)
        assert np.array_equal(ref.view(np.uint32), lit.view(np.uint32)), (name, axis, g, int((ref.view(np.uint32) != lit.view(np.uint32)).sum()))
        worst = max(worst, float(np.abs(ref.astype(np.float64) - spec).max()))
    assert worst <= SPEC_ABS_TOL, worst


def test_every_assembled_shader_is_exercised():
    R, _ = _libs()
    names = {R.ref_hlsl_shader_name(i).decode() for i in range(R.ref_hlsl_shader_count())}
    want = {"%s_%s" % (n, a) for n, _ in INTERP for a in "xy"} | {"conv_%s_%s" % (n, a) for n, _ in CONV for a in "xy"} | {"final_8", "final_10"}
    assert names == want


@pytest.mark.parametrize("name,Q", [("final_8", 255.0), ("final_10", 1023.0)])
def test_final_pass_equals_the_reference_hlsl(name, Q):
    """ps_final_pass.hlsl: floor(pixel * QUANTIZATION + dither) / QUANTIZATION with the 32x32 dither texture addressed by
    Tex * (size / 32) through a WRAP sampler (FinalPass, DX11VideoProcessor.cpp:3204) -- the shader's float4 against final_pass in LITERAL
    and SPEC mode (bit-equal both: the mad's fusion cannot move a floor here unless the sum sits within half an ULP of an integer), and
    the 8 / 10 bit codes after the render-target conversion."""
    R, L = _libs()
    K = np.ctypeslib.as_array(L.orc_dither_thresholds(), (1024,)).astype(f32) / f32(1024)
    dither = np.ascontiguousarray(np.repeat(K.reshape(32, 32, 1), 4, axis=2))
    W, H = 128, 64
    rng = np.random.Generator(np.random.PCG64(5))
    tex = rng.uniform(-0.05, 1.05, (H, W, 4)).astype(f32)
    tex[:8] = rng.integers(0, 1024, (8, W, 4)).astype(f32) / f32(1023)           # the values a 10-bit intermediate texture returns
    tex[8:16] = rng.integers(0, 256, (8, W, 4)).astype(f32) / f32(255)
    uv = np.empty((H, W, 2), f32)
    uv[..., 0] = ((np.arange(W, dtype=f32) + f32(.5)) / f32(W))[None, :]; uv[..., 1] = ((np.arange(H, dtype=f32) + f32(.5)) / f32(H))[:, None]
    ref = np.empty((H, W, 4), f32)
    assert R.ref_hlsl_run(name.encode(), _p(tex), W, H, 0, _p(dither), 32, 32, _p(uv), W * H, f32(W) / f32(32), f32(H) / f32(32), _p(ref)) == 0
    rgb = np.ascontiguousarray(tex[..., :3])
    for mode in (1, 0):
        o = np.empty((H, W, 3), f32)
        assert L.orc_test_final_pass(mode, _p(rgb), W, H, Q, _p(o)) == 0
        assert np.array_equal(np.ascontiguousarray(ref[..., :3]).view(np.uint32), o.view(np.uint32)), (name, mode)
    codes = np.floor(np.clip(ref[..., :3].astype(np.float64), 0, 1) * Q + 0.5)     # D3D11 float -> UNORM
    assert codes.min() == 0 and codes.max() == Q                                    # the sample covers the whole code range


# ---- the one-pass Jinc2 upscaler (Shaders/examples/ps_resize_onepass_jinc2.hlsl, resource IDF_PS_11_INTERP_JINC2) ----
REF_C = os.path.join(O.ORC_DIR, "_ref", "libref_hlsl_convert.so")       # the unit built over hlsl_vec.h also holds this shader


@pytest.mark.skipif(not os.path.exists(REF_C), reason="oracle/_ref/libref_hlsl_convert.so not built")
@pytest.mark.parametrize("geom", [(64, 32, (0, 0, 64, 32), 128, 64), (64, 32, (3, 2, 61, 30), 97, 45), (32, 32, (0, 0, 32, 32), 83, 71),
                                  (16, 64, (1, 0, 16, 63), 40, 150), (64, 32, (0, 0, 64, 32), 64, 32)],
                         ids=["2x", "offset_rect", "odd_ratio", "tall", "1to1"])
def test_jinc2_shader_equals_the_reference_hlsl(geom):
    """The two-dimensional 4x4 Jinc2 resampler with anti-ringing: distances to the 16 texels around floor(pc - .5), the windowed-jinc weight
    of each (sin * sin / d^2, wa * wb at d = 0), mul(weights[r], float4x3(...)) row by row, the weight total, clamp to the four central
    texels, lerp by 0.8 -- the shader text compiled as C++ (float4x4 / float4x3 / mul in hlsl_vec.h; resampler()'s vector conditional
    rewritten to hlsl_select) against the oracle's jinc_pass in LITERAL mode, bit for bit, border clamping included."""
    R = C.CDLL(REF_C); L = O.lib()
    R.ref_jinc2_run.argtypes = [FP, C.c_int, C.c_int, FP, C.c_int, FP]
    L.orc_test_jinc_draw.argtypes = [C.c_int, FP, C.c_int, C.c_int, C.c_int * 4, C.c_int, C.c_int, FP]
    W, H, rect, ow, oh = geom
    rng = np.random.Generator(np.random.PCG64(90 + ow))
    tex = rng.uniform(0, 1, (H, W, 4)).astype(f32)
    l, t, r, b = rect
    gx = f32(r - l) / f32(ow); gy = f32(b - t) / f32(oh)
    cx = (np.arange(ow, dtype=f32) + f32(.5)) * gx + f32(l); cy = (np.arange(oh, dtype=f32) + f32(.5)) * gy + f32(t)
    uv = np.empty((oh, ow, 2), f32); uv[..., 0] = (cx / f32(W))[None, :]; uv[..., 1] = (cy / f32(H))[:, None]
    ref = np.empty((oh, ow, 4), f32)
    assert R.ref_jinc2_run(_p(tex), W, H, _p(uv), ow * oh, _p(ref)) == 0
    rgb = np.ascontiguousarray(tex[..., :3])
    lit = np.empty((oh, ow, 3), f32); spec = np.empty((oh, ow, 3), f32)
    assert L.orc_test_jinc_draw(1, _p(rgb), W, H, (C.c_int * 4)(*rect), ow, oh, _p(lit)) == 0
    assert L.orc_test_jinc_draw(0, _p(rgb), W, H, (C.c_int * 4)(*rect), ow, oh, _p(spec)) == 0
    a = np.ascontiguousarray(ref[..., :3])
    assert np.array_equal(a.view(np.uint32), lit.view(np.uint32)), (geom, int((a.view(np.uint32) != lit.view(np.uint32)).sum()))
    assert np.all(ref[..., 3] == 1)
    assert float(np.abs(a.astype(np.float64) - spec).max()) <= 1e-6            # SPEC: fused mads and the sin polynomial


# ---- the Direct3D 9 twins (Shaders/d3d9/*.hlsl) against the oracle's variant = D3D9 ----
D9_INTERP = [("mitchell4", U["Mitchell"]), ("catmull4", U["CatmullRom"]), ("lanczos2", U["Lanczos2"]), ("lanczos3", U["Lanczos3"])]


def _draw9(R, L, name, axis, kind, filt, W, H, rect, ow, oh, seed):
    """as _draw, for a D3D9 shader: TEXCOORD0 is in texel units and carries the -0.5 CDX9VideoProcessor::TextureResizeShader gives the quad
    (DX9VideoProcessor.cpp:2880-2883: tx0 = srcRect.left - 0.5, ...), i.e. tex = (origin + (i + .5) * scale) - 0.5"""
    rng = np.random.Generator(np.random.PCG64(seed))
    tex = rng.uniform(0, 1, (H, W, 4)).astype(f32)
    l, t, r, b = rect
    gx = f32(r - l) / f32(ow); gy = f32(b - t) / f32(oh)
    cx = ((np.arange(ow, dtype=f32) + f32(.5)) * gx + f32(l)) - f32(.5)
    cy = ((np.arange(oh, dtype=f32) + f32(.5)) * gy + f32(t)) - f32(.5)
    txy = np.empty((oh, ow, 2), f32); txy[..., 0] = cx[None, :]; txy[..., 1] = cy[:, None]
    ref = np.empty((oh, ow, 4), f32)
    assert R.ref_d3d9_run(name.encode(), _p(tex), W, H, _p(txy), ow * oh, gx, gy, _p(ref)) == 0, name
    rgb = np.ascontiguousarray(tex[..., :3])
    lit = np.empty((oh, ow, 3), f32)
    assert L.orc_test_resize_draw(1, axis, kind, filt, O.VARIANT_D3D9, _p(rgb), W, H, (C.c_int * 4)(*rect), ow, oh, _p(lit)) == 0
    return np.ascontiguousarray(ref[..., :3]), lit


@pytest.mark.skipif(not os.path.exists(REF_C), reason="oracle/_ref/libref_hlsl_convert.so not built")
@pytest.mark.parametrize("axis", [0, 1], ids=["x", "y"])
@pytest.mark.parametrize("family", ["interpolation", "convolution"])
def test_d3d9_resize_shaders_equal_the_reference_hlsl(family, axis):
    """The reference keeps a Direct3D 9 renderer whose resize shaders differ from the D3D11 files in the conventions SURVEY Appendix B
    lists: texel-space texcoords carrying -0.5, Lanczos3 reads six different texels (no Q1 := Q0), the convolution is centred (no half-texel
    shift).  The oracle and the kernels implement both (`variant`); here the 20 D3D9 shader builds run as C++ and the oracle's LITERAL mode
    with variant = D3D9 equals them bit for bit."""
    R = C.CDLL(REF_C); L = O.lib()
    R.ref_d3d9_run.argtypes = [C.c_char_p, FP, C.c_int, C.c_int, FP, C.c_int, C.c_float, C.c_float, FP]
    L.orc_test_resize_draw.argtypes = [C.c_int] * 5 + [FP, C.c_int, C.c_int, C.c_int * 4, C.c_int, C.c_int, FP]
    table, geoms, kind = (D9_INTERP, UP_X, 1) if family == "interpolation" else (CONV, DN_X, 2)
    for name, filt in table:
        for k, g in enumerate(geoms):
            W, H, rect, ow, oh = g if axis == 0 else _transpose(g)
            full = "d9_%s%s_%s" % ("" if kind == 1 else "conv_", name, "xy"[axis])
            ref, lit = _draw9(R, L, full, axis, kind, filt, W, H, rect, ow, oh, seed=200 + k)
            assert np.array_equal(ref.view(np.uint32), lit.view(np.uint32)), (full, g, int((ref.view(np.uint32) != lit.view(np.uint32)).sum()))


# ---- turned draws: SetRotation / SetFlip (FillVertices, DX11VideoProcessor.cpp:130-179) ----
def _turned_uv(rect, W, H, ow, oh, rot, flip):
    """TEXCOORD of every output pixel of a turned quad: u comes from x (rot 0 / 180) or y (90 / 270) and runs backwards for 180 / 270 (xor flip);
    v comes from the other index and runs backwards for 90 / 180 (the reference's vertex table; tests/test_reference_resize_pin.py holds the
    oracle's reading of it to FillVertices compiled in place)"""
    l, t, r, b = rect
    rotated = rot in (90, 270)
    u_len, v_len = (oh, ow) if rotated else (ow, oh)
    u_rev = (rot in (180, 270)) ^ bool(flip); v_rev = rot in (90, 180)
    su = f32(r - l) / f32(u_len); sv = f32(b - t) / f32(v_len)
    iu = np.arange(u_len, dtype=f32) + f32(.5); iv = np.arange(v_len, dtype=f32) + f32(.5)
    cu = (f32(r) - iu * su) if u_rev else (iu * su + f32(l))
    cv = (f32(b) - iv * sv) if v_rev else (iv * sv + f32(t))
    uv = np.empty((oh, ow, 2), f32)
    if rotated:
        uv[..., 0] = (cu / f32(W))[:, None]; uv[..., 1] = (cv / f32(H))[None, :]
    else:
        uv[..., 0] = (cu / f32(W))[None, :]; uv[..., 1] = (cv / f32(H))[:, None]
    return uv


@pytest.mark.skipif(not os.path.exists(REF_C), reason="oracle/_ref/libref_hlsl_convert.so not built")
@pytest.mark.parametrize("rot,flip", [(90, 0), (180, 0), (270, 0), (0, 1), (90, 1), (180, 1)])
def test_turned_draws_equal_the_reference_hlsl(rot, flip):
    """the X / Y shaders and Jinc2 on rotated and mirrored quads: interpolation (Lanczos3, incl. the Q1 := Q0 tap on a reversed axis), the
    convolution with the reference's same-named-axis `scale` constant on a rotated quad, and the two-dimensional Jinc2"""
    R, L = _libs()
    RJ = C.CDLL(REF_C); RJ.ref_jinc2_run.argtypes = [FP, C.c_int, C.c_int, FP, C.c_int, FP]
    L.orc_test_oriented_draw.argtypes = [C.c_int] * 7 + [FP, C.c_int, C.c_int, C.c_int * 4, C.c_int, C.c_int, FP]
    W, H, rect = 64, 32, (3, 2, 61, 30)
    rng = np.random.Generator(np.random.PCG64(500 + rot + flip))
    tex = rng.uniform(0, 1, (H, W, 4)).astype(f32)
    rgb = np.ascontiguousarray(tex[..., :3])
    rotated = rot in (90, 270)
    for label, in_axis, kind, filt, ow, oh in (("lanczos3_x", 0, 1, U["Lanczos3"], 45, 97) if rotated else ("lanczos3_x", 0, 1, U["Lanczos3"], 97, 45),
                                               ("catmull4_y", 1, 1, U["CatmullRom"], 70, 29) if rotated else ("catmull4_y", 1, 1, U["CatmullRom"], 29, 70),
                                               ("conv_hamming_x", 0, 2, D["Hamming"], 28, 20) if rotated else ("conv_hamming_x", 0, 2, D["Hamming"], 20, 28),
                                               ("conv_bicubic05_y", 1, 2, D["Bicubic"], 9, 58) if rotated else ("conv_bicubic05_y", 1, 2, D["Bicubic"], 58, 9),
                                               ("jinc2", 2, 1, U["Jinc2"], 101, 77)):
        uv = _turned_uv(rect, W, H, ow, oh, rot, flip)
        ref = np.empty((oh, ow, 4), f32)
        if in_axis == 2:
            assert RJ.ref_jinc2_run(_p(tex), W, H, _p(uv), ow * oh, _p(ref)) == 0
        else:
            # TextureResizeShader's constant: scale = srcRect size / dstRect size of the same-named axes, rotated quad or not (:351-354)
            sx = f32(rect[2] - rect[0]) / f32(ow); sy = f32(rect[3] - rect[1]) / f32(oh)
            assert R.ref_hlsl_run(label.encode(), _p(tex), W, H, 0, None, 0, 0, _p(uv), ow * oh, sx, sy, _p(ref)) == 0
        lit = np.empty((oh, ow, 3), f32)
        assert L.orc_test_oriented_draw(1, in_axis, kind, filt, O.VARIANT_D3D11, rot, flip, _p(rgb), W, H, (C.c_int * 4)(*rect), ow, oh, _p(lit)) == 0, label
        a = np.ascontiguousarray(ref[..., :3])
        assert np.array_equal(a.view(np.uint32), lit.view(np.uint32)), (label, rot, flip, int((a.view(np.uint32) != lit.view(np.uint32)).sum()))
