"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle in SPEC mode, bit for bit.

Integer/byte outputs (BGRA8, RGB10A2 codes) must be identical; RGBA16F outputs must be identical fp16
bit patterns.  The float stage is additionally checked before quantisation (<= 1 ulp; in practice 0)."""
import numpy as np
import pytest

import harness as H
import orclib as O

pytestmark = pytest.mark.gpu

PATH_AUTO, PATH_MULTIPASS, PATH_FUSED = 0, 1, 2


def _run_gpu(case, buf, pitch, path=PATH_MULTIPASS, capture=False):
    vp = case.make_processor(device=0, path=path)
    if capture:
        vp.debug_capture_f32(True)
    side = getattr(case, "side_data", None) or {}
    vp.ProcessSample(data=buf, pitch=pitch, hdr=side.get("hdr"), hdr_cll=side.get("cll"), dovi=getattr(case, "dovi", None))
    out = vp.ReadBackbuffer()
    return vp, out


def _check(case, kind="noise", seed=0, path=None):
    """path None: both kernel families (multipass, and whatever the planner picks = fused when the plan has a fused form)"""
    codes = H.make_codes(case.cf, case.w, case.h, kind, seed)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes, bottom_up=bool(case.bottom_up))
    ref = H.run_oracle(case, planes, pitches)
    vp = None
    for p in ((PATH_MULTIPASS, PATH_AUTO, PATH_FUSED) if path is None else (path,)):
        try:
            vp, out = _run_gpu(case, buf, pitch, p)
        except Exception as e:        # PATH_FUSED on a plan without a fused form must fail with E_NOTIMPL, nothing else
            assert p == PATH_FUSED and path is None and "80004001" in str(e).lower(), e
            continue
        n, d = H.compare(out, ref, case.backbuffer)
        assert n == 0, "%s/%s path %d(%d): %d channel values differ (max delta %d)" % (case.name, kind, p, vp.debug_plan().path, n, d)
    return vp


@pytest.mark.parametrize("idx", range(4))
@pytest.mark.parametrize("kind", ["ramp", "noise", "natural"])
def test_baseline_configs_scaled(idx, kind):
    case = H.baseline_cases(scale=0.125)[idx]
    _check(case, kind, seed=idx + 1)


def test_float_stage_before_quantisation():
    """convert + tone-map output before the RT write: <= 1 ulp vs oracle (bit-equal expected)"""
    case = H.baseline_cases(scale=0.1)[1]
    codes = H.make_codes(case.cf, case.w, case.h, "fullnoise", 3)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    _, dbg = H.run_oracle(case, planes, pitches, want_debug=True)
    vp, _ = _run_gpu(case, buf, pitch, capture=True)
    f32, _ = vp.debug_read_surface(3)
    got = f32.view(np.float32).reshape(case.h, case.w, 4)[..., :3]
    exp = dbg["convert_f32"]
    ulp = np.abs(got.view(np.int32).astype(np.int64) - exp.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1, "max ulp distance %d" % ulp.max()
    assert (ulp != 0).sum() == 0, "%d values off by one ulp" % (ulp != 0).sum()
    # intermediate surfaces too
    conv, fmt_ = vp.debug_read_surface(0)
    assert H.compare(conv, dbg["convert_surf"][:, :conv.shape[1]], fmt_) == (0, 0)


def test_integer_stages_from_oracle_intermediate():
    """resize + dither given the oracle's own m_TexConvertOutput: bit exact"""
    case = H.baseline_cases(scale=0.1)[1]
    codes = H.make_codes(case.cf, case.w, case.h, "natural", 5)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    ref, dbg = H.run_oracle(case, planes, pitches, want_debug=True)
    vp = case.make_processor(device=0, path=PATH_MULTIPASS)
    surf = np.ascontiguousarray(dbg["convert_surf"][:, :case.w * 4])
    vp.debug_render_from_convert(surf)
    out = vp.ReadBackbuffer()
    assert H.compare(out, ref, case.backbuffer) == (0, 0)


FORMATS = ["NV12", "P010", "P016", "YUY2", "UYVY", "P210", "P216", "Y210", "Y216", "V210", "AYUV", "Y410", "Y416", "YV12", "YV16",
           "YV24", "YUV420P8", "YUV422P8", "YUV444P8", "YUV420P10", "YUV420P16", "YUV422P10", "YUV422P16", "YUV444P10",
           "YUV444P16"]


@pytest.mark.parametrize("name", FORMATS)
@pytest.mark.parametrize("chroma", [O.CHROMA_Nearest, O.CHROMA_Bilinear, O.CHROMA_CatmullRom])
def test_every_format_and_chroma_mode(name, chroma):
    case = H.Case(O.CF[name], 64, 36, 96, 54, H.EX_709, chroma=chroma, up=O.UPSCALE["Lanczos2"], name=name)
    _check(case, "fullnoise", seed=7)


@pytest.mark.parametrize("siting", [1, 5, 7])
@pytest.mark.parametrize("chroma", [O.CHROMA_Bilinear, O.CHROMA_CatmullRom])
def test_chroma_siting(siting, chroma):
    ex = O.exfmt(chroma=siting, range_=2, matrix=1, prim=2, trc=5)
    _check(H.Case(O.CF["NV12"], 48, 32, 48, 32, ex, chroma=chroma, up=O.UPSCALE["Nearest"], name="siting%d" % siting), "fullnoise", 2)


@pytest.mark.parametrize("up", ["Nearest", "Mitchell", "CatmullRom", "Lanczos2", "Lanczos3"])
@pytest.mark.parametrize("variant", [O.VARIANT_D3D11, O.VARIANT_D3D9])
def test_upscalers(up, variant):
    _check(H.Case(O.CF["P010"], 60, 34, 133, 77, H.EX_2020_PQ, up=O.UPSCALE[up], variant=variant, name="up_" + up), "noise", 3)


@pytest.mark.parametrize("down", ["Box", "Bilinear", "Hamming", "Bicubic", "BicubicSharp", "Lanczos"])
@pytest.mark.parametrize("variant", [O.VARIANT_D3D11, O.VARIANT_D3D9])
def test_downscalers(down, variant):
    _check(H.Case(O.CF["NV12"], 160, 90, 50, 31, H.EX_709, down=O.DOWNSCALE[down], variant=variant, name="down_" + down), "natural", 4)


def test_mixed_axes_and_interpolate_at_50pct():
    # X upscales (interpolation), Y downscales > 2x (convolution)
    _check(H.Case(O.CF["NV12"], 64, 128, 100, 40, H.EX_709, up=O.UPSCALE["CatmullRom"], down=O.DOWNSCALE["Lanczos"], name="mixed"), "noise", 5)
    # exact 0.5x: interpolation by default, convolution when bInterpolateAt50pct is off
    _check(H.Case(O.CF["NV12"], 128, 72, 64, 36, H.EX_709, up=O.UPSCALE["Lanczos3"], interp50=1, name="half_interp"), "noise", 6)
    _check(H.Case(O.CF["NV12"], 128, 72, 64, 36, H.EX_709, down=O.DOWNSCALE["Bicubic"], interp50=0, name="half_conv"), "noise", 6)
    # one axis only
    _check(H.Case(O.CF["NV12"], 64, 36, 128, 36, H.EX_709, up=O.UPSCALE["Lanczos3"], name="x_only"), "noise", 7)
    _check(H.Case(O.CF["NV12"], 64, 36, 64, 90, H.EX_709, up=O.UPSCALE["Mitchell"], name="y_only"), "noise", 7)


@pytest.mark.parametrize("trc", ["PQ", "HLG", "BT709", "G22", "G18", "G20", "G26", "G28", "sRGB", "L10"])
@pytest.mark.parametrize("to_sdr", [0, 1])
def test_transfer_functions_bt2020(trc, to_sdr):
    ex = O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=O.TRC[trc])
    _check(H.Case(O.CF["P010"], 64, 36, 64, 36, ex, to_sdr=to_sdr, up=O.UPSCALE["Nearest"], name="trc_" + trc), "fullnoise", 8)


@pytest.mark.parametrize("tex,bb", [(O.TEXFMT_8INT, O.SURF_BGRA8), (O.TEXFMT_10INT, O.SURF_BGRA8), (O.TEXFMT_16FLOAT, O.SURF_BGRA8),
                                    (O.TEXFMT_10INT, O.SURF_RGB10A2), (O.TEXFMT_16FLOAT, O.SURF_RGB10A2),
                                    (O.TEXFMT_16FLOAT, O.SURF_RGBA16F), (O.TEXFMT_AUTOINT, O.SURF_RGB10A2)])
@pytest.mark.parametrize("dither", [0, 1])
def test_internal_and_backbuffer_formats(tex, bb, dither):
    _check(H.Case(O.CF["P010"], 48, 28, 96, 56, H.EX_2020_PQ, backbuffer=bb, tex=tex, dither=dither, up=O.UPSCALE["Lanczos3"], name="fmt"), "natural", 9)
    _check(H.Case(O.CF["P010"], 48, 28, 48, 28, H.EX_2020_PQ, backbuffer=bb, tex=tex, dither=dither, name="fmt_noresize"), "natural", 9)


def test_rects_crop_letterbox_offscreen():
    # source crop
    _check(H.Case(O.CF["NV12"], 96, 64, 120, 80, H.EX_709, src_rect=(8, 6, 88, 58), up=O.UPSCALE["Lanczos3"], name="crop"), "noise", 10)
    # letterboxed: video rect inside a larger window (rest stays black), dither anchored to the window (B11)
    _check(H.Case(O.CF["P010"], 64, 36, 128, 72, H.EX_2020_PQ, dst_rect=(16, 10, 144, 82), win=(160, 100), up=O.UPSCALE["CatmullRom"], name="letterbox"), "noise", 11)
    # video rect partly outside the window
    _check(H.Case(O.CF["P010"], 64, 36, 128, 72, H.EX_2020_PQ, dst_rect=(-20, -10, 108, 62), win=(100, 60), up=O.UPSCALE["Lanczos3"], name="offscreen"), "noise", 12)
    # same size but offset -> copy path with final pass
    _check(H.Case(O.CF["P010"], 64, 36, 64, 36, H.EX_2020_PQ, dst_rect=(5, 3, 69, 39), win=(80, 50), name="offset_copy"), "noise", 13)


def test_procamp_and_ranges():
    pa = dict(brightness=12.0, contrast=1.15, hue=25.0, saturation=0.8)
    _check(H.Case(O.CF["NV12"], 64, 36, 64, 36, H.EX_709, procamp=pa, up=O.UPSCALE["Nearest"], name="procamp"), "fullnoise", 14)
    ex_pc = O.exfmt(chroma=5, range_=1, matrix=2, prim=2, trc=5)
    _check(H.Case(O.CF["YUY2"], 64, 36, 64, 36, ex_pc, name="pc601"), "fullnoise", 15)
    ex_ycgco = O.exfmt(chroma=0, range_=1, matrix=7, prim=2, trc=5)
    _check(H.Case(O.CF["YUV444P8"], 64, 36, 64, 36, ex_ycgco, name="ycgco"), "fullnoise", 16)


def test_deint_blend():
    _check(H.Case(O.CF["NV12"], 64, 36, 64, 36, H.EX_709, interlaced=1, deint_blend=1, up=O.UPSCALE["Nearest"], name="deint"), "fullnoise", 17)


def test_nits_sweep():
    for nits in (25, 80, 125, 200, 400):
        _check(H.Case(O.CF["P010"], 48, 28, 48, 28, H.EX_2020_PQ, nits=nits, name="nits%d" % nits), "fullnoise", nits)


def test_math_contract_on_device():
    """LG2 / EX2 / POW and the UNORM reads evaluated on the GPU equal the oracle's, bit for bit"""
    import ctypes as C
    import videorenderer_b200 as V
    L = V.lib(); OL = O.lib()
    rng = np.random.Generator(np.random.PCG64(0))
    x = np.concatenate([np.exp(rng.uniform(-80, 10, 200000)), rng.uniform(0, 1, 200000), [1.0, 0.5, 2.0, 1e-38, 1.17549435e-38, 0.0]]).astype(np.float32)
    out = np.empty_like(x)
    assert L.mpcvr_debug_math(0, 0, x.ctypes.data, None, out.ctypes.data, x.size) == 0
    exp = np.array([OL.orc_test_lg2(float(v)) for v in x[x > 1e-37][:50000]], np.float32)
    np.testing.assert_array_equal(out[x > 1e-37][:50000].view(np.uint32), exp.view(np.uint32))
    t = np.concatenate([rng.uniform(-130, 5, 200000), [0.0, -126.0, -125.5, -125.49, -0.5, 0.5, 127.4, 128.0]]).astype(np.float32)
    out = np.empty_like(t)
    assert L.mpcvr_debug_math(0, 1, t.ctypes.data, None, out.ctypes.data, t.size) == 0
    exp = np.array([OL.orc_test_ex2(float(v)) for v in t[:50000]], np.float32)
    np.testing.assert_array_equal(out[:50000].view(np.uint32), exp.view(np.uint32))
    # pow over the exponents the shaders use
    for y in (1 / 78.84375, 1 / 0.1593017578125, 1 / 2.2, 2.2, 0.2, 0.1593017578125, 78.84375):
        xs = rng.uniform(0, 1.2, 20000).astype(np.float32); ys = np.full_like(xs, np.float32(y)); out = np.empty_like(xs)
        assert L.mpcvr_debug_math(0, 2, xs.ctypes.data, ys.ctypes.data, out.ctypes.data, xs.size) == 0
        exp = np.array([OL.orc_test_pow(float(a), float(np.float32(y)), 0) for a in xs], np.float32)
        np.testing.assert_array_equal(out.view(np.uint32), exp.view(np.uint32))
    # UNORM reads are the correctly rounded quotients
    for op, n in ((3, 65536), (4, 1024), (5, 256)):
        c = np.arange(n, dtype=np.float32); out = np.empty_like(c)
        assert L.mpcvr_debug_math(0, op, c.ctypes.data, None, out.ctypes.data, n) == 0
        np.testing.assert_array_equal(out, (c / np.float32(n - 1)).astype(np.float32))


def test_full_size_properties_c2():
    """BASELINE config 2 at full size (3840x2160 -> 7680x4320): size-independent properties.
    (a) a constant frame stays constant and equals the small-frame oracle value (filters have unit DC gain
        after the wc compensation, dither pattern repeats every 32 px);
    (b) the full frame equals the oracle on a band of rows (oracle run on a cropped source with matching rects)."""
    case = H.baseline_cases(1.0)[1]
    codes = H.make_codes(case.cf, case.w, case.h, "natural", 21)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    vp, out = _run_gpu(case, buf, pitch, PATH_MULTIPASS)
    # band check: source rows 1000..1063 with halo -> output rows 2000..2127; compare the interior rows
    y0, y1, halo = 1000, 1064, 8
    sub = H.Case(case.cf, case.w, case.h, case.dst_w, 2 * (y1 - y0 + 2 * halo), case.ex, up=case.up,
                 src_rect=(0, y0 - halo, case.w, y1 + halo), name="band")
    ref = H.run_oracle(sub, planes, pitches)
    # dither is anchored to window coordinates: the band's window row r corresponds to full-frame row r + 2*(y0-halo),
    # which is a multiple of 32 when 2*(y0-halo) % 32 == 0
    assert (2 * (y0 - halo)) % 32 == 0
    a = out[2 * y0:2 * y1]; b = ref[2 * halo:2 * halo + 2 * (y1 - y0)]
    assert H.compare(a, b, case.backbuffer) == (0, 0)


# ------------------------------------------------------------------------------------------------
# fused single-kernel family (exact 2x interpolation upscales)
# ------------------------------------------------------------------------------------------------
def _check_fused(case, kind="noise", seed=0):
    vp = _check(case, kind, seed, path=PATH_FUSED)
    assert vp.debug_plan().path == PATH_FUSED
    assert vp.LastFrameKernels() <= 2          # the frame kernel (+ a clear when the video rect does not cover the window)
    return vp


@pytest.mark.parametrize("up", ["Mitchell", "CatmullRom", "Lanczos2", "Lanczos3"])
@pytest.mark.parametrize("variant", [O.VARIANT_D3D11, O.VARIANT_D3D9])
def test_fused_2x_filters(up, variant):
    _check_fused(H.Case(O.CF["P010"], 150, 70, 300, 140, H.EX_2020_PQ, up=O.UPSCALE[up], variant=variant, name="fused_" + up), "noise", 31)


@pytest.mark.parametrize("name", ["NV12", "P010", "P016", "YUY2", "UYVY", "Y210", "AYUV", "Y410", "Y416", "YV12", "YUV422P10", "YUV444P16", "P210"])
def test_fused_2x_formats(name):
    _check_fused(H.Case(O.CF[name], 72, 40, 144, 80, H.EX_709, up=O.UPSCALE["Lanczos3"], name="fused_" + name), "fullnoise", 32)


@pytest.mark.parametrize("tex,bb", [(O.TEXFMT_8INT, O.SURF_BGRA8), (O.TEXFMT_10INT, O.SURF_BGRA8), (O.TEXFMT_16FLOAT, O.SURF_BGRA8),
                                    (O.TEXFMT_10INT, O.SURF_RGB10A2), (O.TEXFMT_16FLOAT, O.SURF_RGB10A2),
                                    (O.TEXFMT_16FLOAT, O.SURF_RGBA16F), (O.TEXFMT_8INT, O.SURF_RGB10A2)])
@pytest.mark.parametrize("dither", [0, 1])
def test_fused_2x_surface_formats(tex, bb, dither):
    _check_fused(H.Case(O.CF["P010"], 66, 34, 132, 68, H.EX_2020_PQ, backbuffer=bb, tex=tex, dither=dither, up=O.UPSCALE["Lanczos3"], name="fused_fmt"), "natural", 33)
    _check_fused(H.Case(O.CF["YUY2"], 64, 36, 128, 72, H.EX_2020_HLG, backbuffer=bb, tex=tex, dither=dither, up=O.UPSCALE["CatmullRom"], name="fused_fmt_hlg"), "natural", 34)


def test_fused_2x_rects():
    _check_fused(H.Case(O.CF["NV12"], 96, 64, 160, 104, H.EX_709, src_rect=(8, 6, 88, 58), up=O.UPSCALE["Lanczos3"], name="fused_crop"), "noise", 35)
    _check_fused(H.Case(O.CF["P010"], 64, 36, 128, 72, H.EX_2020_PQ, dst_rect=(16, 11, 144, 83), win=(160, 100), up=O.UPSCALE["Lanczos3"], name="fused_letterbox"), "noise", 36)
    # tile edges: sizes around the 64-pixel tile
    for w, h in ((62, 66), (64, 64), (66, 62), (130, 10), (10, 130), (2, 2)):
        c = H.Case(O.CF["NV12"], w, h, 2 * w, 2 * h, H.EX_709, up=O.UPSCALE["Lanczos3"], name="fused_edge_%dx%d" % (w, h))
        _check_fused(c, "noise", 37)
    # odd crop sizes (source rect may be odd even if the frame is even)
    _check_fused(H.Case(O.CF["NV12"], 96, 64, 2 * 77, 2 * 51, H.EX_709, src_rect=(3, 5, 80, 56), up=O.UPSCALE["CatmullRom"], name="fused_oddcrop"), "noise", 38)


def test_fused_falls_back_when_unsupported():
    """plans whose tile window cannot fit in shared memory must still render (multipass) under PATH_AUTO, and fail loudly under PATH_FUSED"""
    import videorenderer_b200 as V
    case = H.Case(O.CF["NV12"], 1024, 512, 64, 32, H.EX_709, down=O.DOWNSCALE["Lanczos"], name="huge_downscale")
    codes = H.make_codes(case.cf, case.w, case.h, "noise", 1)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    ref = H.run_oracle(case, planes, pitches)
    vp = case.make_processor(device=0, path=0)
    vp.ProcessSample(data=buf, pitch=pitch)
    assert vp.debug_plan().path == PATH_MULTIPASS
    assert H.compare(vp.ReadBackbuffer(), ref, case.backbuffer) == (0, 0)
    vp = case.make_processor(device=0, path=PATH_FUSED)
    hr = vp.ProcessSample(data=buf, pitch=pitch, check=False)
    assert hr == V.HRESULT.E_NOTIMPL


def test_generic_fused_covers_the_baseline_configs():
    """every single-GPU BASELINE config runs as ONE kernel per frame (C1: 1:1 kernel, C2: 2x kernel, C3/C4: generic kernel)"""
    for case in H.baseline_cases(scale=0.125):
        vp = _check_fused(case, "natural", 51)
        assert vp.LastFrameKernels() == 1


@pytest.mark.parametrize("kind", ["ramp", "noise", "natural"])
def test_fused_c2_scaled(kind):
    _check_fused(H.baseline_cases(scale=0.25)[1], kind, 39)


def test_full_size_c2_fused_equals_multipass_and_oracle_band():
    case = H.baseline_cases(1.0)[1]
    codes = H.make_codes(case.cf, case.w, case.h, "natural", 21)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    _, out_f = _run_gpu(case, buf, pitch, PATH_FUSED)
    _, out_m = _run_gpu(case, buf, pitch, PATH_MULTIPASS)
    assert H.compare(out_f, out_m, case.backbuffer) == (0, 0)
    y0, y1, halo = 1000, 1064, 8
    sub = H.Case(case.cf, case.w, case.h, case.dst_w, 2 * (y1 - y0 + 2 * halo), case.ex, up=case.up,
                 src_rect=(0, y0 - halo, case.w, y1 + halo), name="band")
    ref = H.run_oracle(sub, planes, pitches)
    a = out_f[2 * y0:2 * y1]; b = ref[2 * halo:2 * halo + 2 * (y1 - y0)]
    assert H.compare(a, b, case.backbuffer) == (0, 0)


def test_two_lane_math_equals_scalar_math():
    """mpcvr_math2.cuh (FFMA2 forms) against mpcvr_math.cuh on the device, bit for bit, both lanes"""
    import videorenderer_b200 as V
    L = V.lib()
    rng = np.random.Generator(np.random.PCG64(5))

    def run(op, x, y=None):
        out = np.empty_like(x)
        assert L.mpcvr_debug_math(0, op, x.ctypes.data, None if y is None else y.ctypes.data, out.ctypes.data, x.size) == 0
        return out.view(np.uint32)

    x = np.concatenate([rng.uniform(0, 1, 300000), np.exp(rng.uniform(-80, 0, 100000)), [0.0, 1.0, 1e-38, 1.17549435e-38, 1e-40, 0.5]]).astype(np.float32)
    np.testing.assert_array_equal(run(10, x[x > 1e-37]), run(0, x[x > 1e-37]))
    np.testing.assert_array_equal(run(20, x[x > 1e-37]), run(0, x[x > 1e-37]))
    t = np.concatenate([rng.uniform(-140, 10, 300000), [0.0, -126.0, -127.0, 128.0]]).astype(np.float32)
    np.testing.assert_array_equal(run(11, t), run(1, t))
    for yv in (1 / 78.84375, 1 / 2.2):
        y = np.full_like(x, np.float32(yv))
        np.testing.assert_array_equal(run(12, x, y), run(2, x, y))
        np.testing.assert_array_equal(run(13, x, y), run(2, x, y))
    y = np.full_like(x, np.float32(1.0) / np.float32(0.1593017578125))
    np.testing.assert_array_equal(run(12, x, y), run(2, x, y))
    np.testing.assert_array_equal(run(14, x, y), run(2, x, y))
    np.testing.assert_array_equal(run(21, x, y), run(2, x, y))
    # divisions on the tone-map operand ranges: fast exact sequence == IEEE division
    num = np.concatenate([rng.uniform(0, 0.165, 200000), rng.uniform(0, 1000, 200000), [0.0]]).astype(np.float32)
    den = np.concatenate([rng.uniform(0.16, 19, 200000), rng.uniform(0.05, 2000, 200000), [0.1640625]]).astype(np.float32)
    ref = run(17, num, den)
    np.testing.assert_array_equal(run(16, num, den), ref)
    np.testing.assert_array_equal(run(15, num, den), ref)
    v = rng.uniform(-0.2, 1.2, 300000).astype(np.float32)
    np.testing.assert_array_equal(run(18, v), run(19, v))


def test_two_lane_pq_convert_equals_scalar_convert():
    """the FFMA2 form of the PQ->SDR shader body against the scalar form, bit for bit, on random yuv"""
    case = H.baseline_cases(0.1)[1]
    vp = case.make_processor(device=0)
    rng = np.random.Generator(np.random.PCG64(9))
    n = 200000
    yuv = np.stack([rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n)], -1).astype(np.float32)
    yuv[:1000, 0] = rng.uniform(0.06, 0.08, 1000)          # near black: the PQ curve is steepest there
    a = np.empty_like(yuv); b = np.empty_like(yuv)
    assert vp._L.mpcvr_debug_probe_pq(vp._h, yuv.ctypes.data, a.ctypes.data, n, 0) == 0
    assert vp._L.mpcvr_debug_probe_pq(vp._h, yuv.ctypes.data, b.ctypes.data, n, 1) == 0
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("name", ["NV12", "P010", "YUY2", "AYUV", "YV12", "YUV444P10", "Y210"])
@pytest.mark.parametrize("tex,bb,dither", [(O.TEXFMT_AUTOINT, O.SURF_BGRA8, 1), (O.TEXFMT_16FLOAT, O.SURF_RGB10A2, 1), (O.TEXFMT_16FLOAT, O.SURF_RGBA16F, 0),
                                           (O.TEXFMT_10INT, O.SURF_BGRA8, 0), (O.TEXFMT_8INT, O.SURF_RGB10A2, 1)])
def test_fused_1x_plans(name, tex, bb, dither):
    """no axis resampled: convert (+ final pass) runs as ONE kernel straight into the back buffer (BASELINE config 1 shape)"""
    ex = H.EX_2020_PQ if name in ("P010", "YUV444P10", "Y210") else H.EX_709
    _check_fused(H.Case(O.CF[name], 66, 38, 66, 38, ex, backbuffer=bb, tex=tex, dither=dither, name="fused1x_" + name), "fullnoise", 41)
    # same size, offset inside a larger window (copy path of ResizeShaderPass)
    _check_fused(H.Case(O.CF[name], 66, 38, 66, 38, ex, backbuffer=bb, tex=tex, dither=dither, dst_rect=(7, 5, 73, 43), win=(90, 50), name="fused1x_off_" + name), "fullnoise", 42)


def test_fused_c1_full_size():
    case = H.baseline_cases(1.0)[0]
    vp = _check_fused(case, "natural", 43)
    assert vp.LastFrameKernels() == 1


# ---- RGB / planar RGB / gray samples (read in their own byte layout; the oracle goes through the reference's
#      CPU repack into a DXGI texture first) ----
RGB_SAMPLES = ["RGB24", "XRGB32", "ARGB32", "r210", "RGB48", "BGR48", "BGRA64", "B64A"]
PLANAR_GRAY = ["GBRP8", "GBRP10", "GBRP16", "Y8", "Y10", "Y16"]
PROCAMP = dict(brightness=12.0, contrast=1.25)


@pytest.mark.parametrize("name", RGB_SAMPLES + PLANAR_GRAY)
@pytest.mark.parametrize("procamp", [None, PROCAMP], ids=["default", "procamp"])
def test_rgb_and_gray_inputs_upscale(name, procamp):
    """default ProcAmp + RGB sample = no convert pass (the chain samples the source itself, DX11VideoProcessor.cpp:3320-3322)"""
    _check(H.Case(O.CF[name], 64, 36, 96, 54, 0, up=O.UPSCALE["Lanczos3"], procamp=procamp, name=name), "fullnoise", seed=11)


@pytest.mark.parametrize("name", ["RGB24", "r210", "RGB48", "GBRP10", "Y16"])
@pytest.mark.parametrize("backbuffer", [O.SURF_BGRA8, O.SURF_RGB10A2])
def test_rgb_inputs_one_to_one_and_final_pass(name, backbuffer):
    """no scaling: plain copy, or the final pass dithering straight from the source ("pTex = pInputTexture", :3350-3352)"""
    _check(H.Case(O.CF[name], 80, 45, 80, 45, 0, backbuffer=backbuffer, up=O.UPSCALE["Nearest"], name=name), "fullnoise", seed=12)


@pytest.mark.parametrize("name", ["XRGB32", "BGRA64", "GBRP8"])
def test_rgb_inputs_downscale_and_crop(name):
    """convolution downscale of a cropped source rect: without a convert pass the taps outside the rect are real texels"""
    _check(H.Case(O.CF[name], 160, 90, 50, 31, 0, down=O.DOWNSCALE["Lanczos"], src_rect=(16, 8, 144, 80), name=name), "natural", seed=13)
    _check(H.Case(O.CF[name], 160, 90, 200, 120, 0, up=O.UPSCALE["CatmullRom"], src_rect=(16, 8, 144, 80), dst_rect=(10, 6, 190, 110),
                  win=(200, 120), name=name), "fullnoise", seed=14)


@pytest.mark.parametrize("name", ["RGB24", "XRGB32"])
@pytest.mark.parametrize("procamp", [None, PROCAMP], ids=["default", "procamp"])
def test_bottom_up_dib(name, procamp):
    """BI_RGB with positive biHeight: rows stored last to first, negative pitch (:1801-1803, :1246)"""
    _check(H.Case(O.CF[name], 62, 33, 93, 50, 0, up=O.UPSCALE["Mitchell"], procamp=procamp, bottom_up=1, name=name + "_bu"), "fullnoise", seed=15)


def test_v210_equals_y210():
    """the same codes delivered as v210 words and as Y210 give the same picture"""
    outs = []
    for name in ("V210", "Y210"):
        case = H.Case(O.CF[name], 66, 30, 99, 45, H.EX_709, chroma=O.CHROMA_CatmullRom, up=O.UPSCALE["Lanczos2"], name=name)
        codes = H.make_codes(O.CF["Y210"], 66, 30, "fullnoise", 16)
        buf, pitch, _, _ = H.pack(case.cf, 66, 30, codes)
        outs.append(_run_gpu(case, buf, pitch)[1])
    assert np.array_equal(outs[0], outs[1])


# ---- HDR output -- passthrough convert types and the HDR10 local tone-map post-scale step ----
def _hdr_case(cf, ex, tm_type, tonemap=1, side=None, tex=O.TEXFMT_AUTOINT, backbuffer=O.SURF_RGB10A2, display=600, **kw):
    """side: dict(hdr=dict(max_lum, min_lum), cll=(MaxCLL, MaxFALL)) as IMediaSideData delivers it; the oracle gets the raw
    SetHDR10ShaderParams arguments the reference derives from it (DX11VideoProcessor.cpp:2255-2267, :2673-2684, :2721-2724)"""
    if side and "hdr" in side:
        mx, mn = round(side["hdr"]["max_lum"]), round(side["hdr"]["min_lum"] * 10000.0)
        cll, fall = side.get("cll", (0, 0))
    else:                                                      # no (valid) mastering-display blob: Display-P3 defaults, and the
        mx, mn, cll, fall = 1000, 50, 0, 0                     # light-level blob alone is never applied (:2673-2684)
    hdr = dict(out=1, tonemap=tonemap, min=float(mn), max=float(mx), cll=float(cll), fall=float(fall), display=float(display), type=tm_type)
    case = H.Case(cf, 64, 36, 96, 54, ex, backbuffer=backbuffer, tex=tex, up=O.UPSCALE["Lanczos3"], hdr=hdr, name="hdr_tm%d" % tm_type, **kw)
    case.side_data = side
    return case


@pytest.mark.parametrize("tm_type", [1, 2, 3, 4, 5, 6])
def test_hdr10_local_tonemap_types(tm_type):
    side = dict(hdr=dict(max_lum=4000.0, min_lum=0.005), cll=(1200, 300))
    _check(_hdr_case(O.CF["P010"], H.EX_2020_PQ, tm_type, side=side), "natural", seed=21)


@pytest.mark.parametrize("tm_type", [1, 5, 6])
def test_hdr10_tonemap_without_metadata_and_with_final_pass(tm_type):
    """no side data -> Display-P3 defaults; 16F intermediates -> tone-map step writes the ring's second texture, final pass dithers"""
    _check(_hdr_case(O.CF["P010"], H.EX_2020_PQ, tm_type, tex=O.TEXFMT_16FLOAT), "natural", seed=22)
    _check(_hdr_case(O.CF["P016"], H.EX_2020_PQ, tm_type, side=dict(cll=(800, 120)), display=400), "noise", seed=23)


def test_hdr_passthrough_pq_and_hlg():
    """HDR output without tone mapping: PQ is passed through (matrix only), HLG is converted to PQ (Source/Shaders.cpp:886-891)"""
    _check(_hdr_case(O.CF["P010"], H.EX_2020_PQ, 1, tonemap=0), "natural", seed=24)
    _check(_hdr_case(O.CF["P010"], H.EX_2020_HLG, 1, tonemap=0), "natural", seed=25)
    _check(_hdr_case(O.CF["YUY2"], H.EX_2020_HLG, 3, tonemap=1, side=dict(cll=(1000, 200))), "natural", seed=26)


def test_hdr_settings_without_hdr_display_change_nothing():
    """bHdrPassthrough defaults to true, but without an HDR display the SDR conversion stays (m_bHdrPassthroughSupport false)"""
    case = H.baseline_cases(scale=0.05)[1]
    _check(case, "natural", seed=27)


# ---- gamma + UNORM10 quantiser through the threshold table (fused 2x PQ kernel) ----
def test_gamma_threshold_table_is_exhaustively_exact():
    """every float of [0,1]: table path == direct Q10(pow(x, 1/2.2)) read back, bit for bit (1.07e9 values, on the device)"""
    case = H.baseline_cases(scale=0.05)[1]
    vp = case.make_processor(device=0)
    assert vp.debug_verify_gamma_table() == 0


def test_gamma_threshold_table_on_off_same_frame():
    case = H.baseline_cases(scale=0.25)[1]
    codes = H.make_codes(case.cf, case.w, case.h, "fullnoise", 31)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    outs = []
    try:
        for on in (1, 0):
            vp = case.make_processor(device=0, path=PATH_FUSED)
            vp.debug_set_gamma_table(on)
            vp.ProcessSample(data=buf, pitch=pitch)
            outs.append(vp.ReadBackbuffer())
    finally:
        vp.debug_set_gamma_table(1)
    assert np.array_equal(outs[0], outs[1])
    ref = H.run_oracle(case, planes, pitches)
    assert H.compare(outs[0], ref, case.backbuffer) == (0, 0)


# ---- SetRotation / SetFlip (FillVertices, Source/DX11VideoProcessor.cpp:130-179; ResizeShaderPass :3106-3186) ----
def _rot_case(rot, flip, sw, sh, dw, dh, **kw):
    """dw x dh is the upright target size; a quarter turn presents it as dh x dw"""
    if rot in (90, 270):
        dw, dh = dh, dw
    return H.Case(O.CF["NV12"], sw, sh, dw, dh, H.EX_709, rotation=rot, flip=flip, name="rot%d_f%d" % (rot, flip), **kw)


@pytest.mark.parametrize("rot", [0, 90, 180, 270])
@pytest.mark.parametrize("flip", [0, 1])
def test_rotation_flip_copy_and_upscale(rot, flip):
    if rot == 0 and not flip:
        pytest.skip("covered everywhere else")
    _check(_rot_case(rot, flip, 64, 36, 64, 36, up=O.UPSCALE["Nearest"]), "fullnoise", 41)                     # ps_simple copy
    _check(_rot_case(rot, flip, 64, 36, 96, 54, up=O.UPSCALE["Lanczos3"]), "natural", 42)                       # both up: quarter turns filter one direction only
    _check(_rot_case(rot, flip, 64, 36, 96, 36, up=O.UPSCALE["CatmullRom"]), "natural", 43)                     # one direction scaled
    _check(_rot_case(rot, flip, 64, 36, 64, 54, up=O.UPSCALE["Mitchell"]), "natural", 44)


@pytest.mark.parametrize("rot", [90, 180, 270])
def test_rotation_downscale_mixed_and_final_pass(rot):
    _check(_rot_case(rot, 0, 160, 90, 50, 31, down=O.DOWNSCALE["Hamming"]), "natural", 45)                       # both down (conv scale-constant quirk on quarter turns)
    _check(_rot_case(rot, 1, 160, 90, 200, 31, down=O.DOWNSCALE["Bicubic"], up=O.UPSCALE["Lanczos2"]), "natural", 46)   # up one way, down the other: two passes
    c = _rot_case(rot, 0, 64, 36, 96, 54, up=O.UPSCALE["Lanczos3"], tex=O.TEXFMT_16FLOAT)                     # final pass after the turned draw
    _check(c, "natural", 47)
    _check(_rot_case(rot, 1, 64, 36, 80, 60, up=O.UPSCALE["CatmullRom"], dst_rect=None, src_rect=(8, 4, 56, 32)), "fullnoise", 48)


def test_flip_alone_is_lost_when_final_pass_reads_the_convert_output():
    """rSrc == dstRect, a final pass and no rotation: 'pTex = pInputTexture' (:3350-3352) skips the only draw that would flip"""
    a = H.Case(O.CF["P010"], 64, 36, 64, 36, H.EX_709, up=O.UPSCALE["Nearest"], flip=1, name="flip_lost")
    _check(a, "fullnoise", 49)
    assert a.orc_config().flip == 1


def test_rotation_reports_turned_video_size():
    vp = _rot_case(90, 0, 64, 36, 36, 64).make_processor(device=0)
    assert vp.GetVideoSize() == (36, 64)


# ---- UPSCALE_Jinc2, the one-pass 2-D resizer (Shaders/examples/ps_resize_onepass_jinc2.hlsl) ----
def test_sin_contract_on_device():
    import videorenderer_b200 as V
    L = V.lib(); OL = O.lib()
    rng = np.random.Generator(np.random.PCG64(5))
    x = np.concatenate([rng.uniform(0, 16, 60000), np.arange(0, 6) * np.pi, [0.0, 1e-6, np.pi / 2, 15.99]]).astype(np.float32)
    out = np.empty_like(x)
    assert L.mpcvr_debug_math(0, 6, x.ctypes.data, None, out.ctypes.data, x.size) == 0
    exp = np.array([OL.orc_test_sin(float(v)) for v in x], np.float32)
    np.testing.assert_array_equal(out.view(np.uint32), exp.view(np.uint32))
    assert np.abs(exp.astype(np.float64) - np.sin(x.astype(np.float64))).max() < 2e-7


@pytest.mark.parametrize("dst", [(128, 72), (96, 54), (128, 36), (64, 72), (128, 20), (40, 72), (91, 67)])
@pytest.mark.parametrize("variant", [O.VARIANT_D3D11, O.VARIANT_D3D9])
def test_jinc2(dst, variant):
    """both directions up (one draw), one direction only, and up one way / down the other (Jinc2 + convolution through m_TexResize)"""
    _check(H.Case(O.CF["P010"], 64, 36, dst[0], dst[1], H.EX_709, up=O.UPSCALE["Jinc2"], down=O.DOWNSCALE["Hamming"], variant=variant,
                  name="jinc2_%dx%d" % dst), "natural", 51)


def test_jinc2_rects_and_final_pass():
    _check(H.Case(O.CF["NV12"], 96, 64, 160, 104, H.EX_709, src_rect=(8, 6, 88, 58), up=O.UPSCALE["Jinc2"], name="jinc2_crop"), "fullnoise", 52)
    _check(H.Case(O.CF["P010"], 64, 36, 128, 72, H.EX_2020_PQ, dst_rect=(16, 11, 144, 83), win=(160, 100), up=O.UPSCALE["Jinc2"], name="jinc2_letterbox"), "noise", 53)
    _check(H.Case(O.CF["XRGB32"], 64, 36, 100, 60, 0, src_rect=(4, 2, 60, 34), up=O.UPSCALE["Jinc2"], name="jinc2_rgb_direct"), "natural", 54)
    _check(H.Case(O.CF["P010"], 64, 36, 96, 54, H.EX_709, tex=O.TEXFMT_16FLOAT, backbuffer=O.SURF_RGB10A2, up=O.UPSCALE["Jinc2"], name="jinc2_16f"), "natural", 55)


@pytest.mark.parametrize("rot", [0, 90, 180, 270])
@pytest.mark.parametrize("flip", [0, 1])
def test_jinc2_on_rotated_and_flipped_draws(rot, flip):
    """the 2-D Jinc2 shader on a turned quad (Source/DX11VideoProcessor.cpp:3113-3141, FillVertices :130-179): both directions up (one
    draw), one direction only, up one way / down the other (the first of the two passes carries the orientation), crop + final pass"""
    if rot == 0 and not flip:
        pytest.skip("covered by test_jinc2")
    J = O.UPSCALE["Jinc2"]
    _check(_rot_case(rot, flip, 64, 36, 96, 54, up=J), "natural", 61)
    _check(_rot_case(rot, flip, 64, 36, 96, 36, up=J), "natural", 62)
    _check(_rot_case(rot, flip, 64, 36, 64, 54, up=J), "natural", 63)
    _check(_rot_case(rot, flip, 160, 40, 200, 31, up=J, down=O.DOWNSCALE["Bicubic"]), "natural", 64)        # X up (Jinc2, turned), Y down
    _check(_rot_case(rot, flip, 40, 160, 31, 200, up=J, down=O.DOWNSCALE["Hamming"]), "natural", 65)        # X down (turned convolution), Y up (Jinc2)
    _check(_rot_case(rot, flip, 64, 36, 80, 60, up=J, tex=O.TEXFMT_16FLOAT, src_rect=(8, 4, 56, 32)), "fullnoise", 66)


# ---- Dolby Vision reshaping inside the convert stage (Source/Shaders.cpp:531-589, :703-859) ----
@pytest.mark.parametrize("mmr", [0, 1, 2])
@pytest.mark.parametrize("to_sdr", [1, 0])
def test_dolby_vision_reshape(mmr, to_sdr):
    """polynomial-only and MMR (single / several pieces of orders 1..3, mixed with a polynomial piece) curves; reshaped signal
    through the custom YCC->RGB matrix, PQ->linear->LMS->PQ, then tone-mapped to SDR or left as PQ"""
    ex = O.exfmt(chroma=5, range_=1, matrix=0, prim=0, trc=0)            # profile-5 style stream: no PQ/HLG tag
    for seed in (61, 62):
        case = H.Case(O.CF["P010"], 64, 36, 96, 54, ex, up=O.UPSCALE["Lanczos2"], to_sdr=to_sdr, dovi=H.make_dovi(seed, mmr=mmr),
                      backbuffer=O.SURF_BGRA8 if to_sdr else O.SURF_RGB10A2, name="dovi_mmr%d" % mmr)
        _check(case, "fullnoise", seed)


def test_dolby_vision_procamp_and_12bit_base():
    case = H.Case(O.CF["P016"], 48, 28, 48, 28, O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=0), up=O.UPSCALE["Nearest"],
                  dovi=H.make_dovi(63, mmr=1, bits=12, denom=20), procamp=dict(brightness=6.0, contrast=1.1), name="dovi_procamp")
    _check(case, "natural", 63)


def test_dolby_vision_is_ignored_for_tagged_hdr10_unless_preferred():
    """m_bHdrPreferDoVi false and the stream tagged PQ: the side data is not taken (Source/DX11VideoProcessor.cpp:2271)"""
    plain = H.Case(O.CF["P010"], 48, 28, 72, 42, H.EX_2020_PQ, up=O.UPSCALE["CatmullRom"], name="dovi_ignored")
    codes = H.make_codes(plain.cf, plain.w, plain.h, "natural", 64)
    buf, pitch, planes, pitches = H.pack(plain.cf, plain.w, plain.h, codes)
    ref = H.run_oracle(plain, planes, pitches)                            # oracle without Dolby Vision data
    plain.dovi = H.make_dovi(64, mmr=1)                                   # the product is handed the side data anyway
    _, out = _run_gpu(plain, buf, pitch)
    assert H.compare(out, ref, plain.backbuffer) == (0, 0)


@pytest.mark.parametrize("scenario", ["between", "brighter", "dimmer", "equal_targets"])
def test_dolby_vision_l2_trims(scenario):
    """Level 2 trims picked / interpolated for the display (Source/DX11VideoProcessor.cpp:2383-2466) and applied in the to-SDR chain
    (DolbyVisionTrims, Source/Shaders.cpp:787-795).  display = iHdrDisplayMaxNits = 1000 nits -> PQ code ~3079"""
    lo = (2081, 2100, 2030, 2120, 2010, 2075)       # 100-nit target
    mid = (2851, 1990, 2060, 2000, 2080, 2020)      # 600-nit target
    hi = (3388, 2070, 2040, 2090, 2060, 2000)       # 2000-nit target
    l2 = dict(between=[lo, hi, mid], brighter=[lo, mid], dimmer=[hi], equal_targets=[mid, mid])[scenario]
    ex = O.exfmt(chroma=5, range_=1, matrix=0, prim=0, trc=0)
    case = H.Case(O.CF["P010"], 64, 36, 96, 54, ex, up=O.UPSCALE["CatmullRom"], dovi=H.make_dovi(71, mmr=1, l2=l2, source_max_pq=3696),
                  hdr=dict(out=0, display=1000), name="dovi_l2_" + scenario)
    _check(case, "natural", 71)


@pytest.mark.parametrize("trims", ["negative_offset", "bright", "bright_sat", "zero_satgain"])
def test_dolby_vision_l2_trims_nan_pixels(trims):
    """DolbyVisionTrims has no guard in the reference's shader: a negative pow base, a zero luminance or a value past ST2084ToLinear's pole
    make NaN, the gamut matrix spreads it over the pixel and the final saturate writes (0, 0, 0).  tests/test_reference_convert_exec_pin.py
    shows it on the executed shader text and holds the oracle to it; here the kernel against the oracle, with trim sets that provoke each
    cause on full-range noise (1:1 so the black pixels reach the output unfiltered)."""
    l2 = dict(negative_offset=[(3079, 2048, 1200, 2048, 2048, 2300)], bright=[(3079, 4095, 4095, 4095, 2048, 2048)],
              bright_sat=[(3079, 4095, 3500, 4095, 4095, 4095)], zero_satgain=[(3079, 2100, 2000, 2048, 2200, 2048)])[trims]
    ex = O.exfmt(chroma=5, range_=1, matrix=0, prim=0, trc=0)
    case = H.Case(O.CF["P010"], 64, 36, 64, 36, ex, up=O.UPSCALE["Nearest"], dovi=H.make_dovi(72, mmr=1, l2=l2, source_max_pq=3696),
                  hdr=dict(out=0, display=1000), name="dovi_l2_nan_" + trims)
    _check(case, "fullnoise", 72)
    if trims != "zero_satgain":
        codes = H.make_codes(case.cf, case.w, case.h, "fullnoise", 72)
        _, _, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
        rgb = H.rgb_view(H.run_oracle(case, planes, pitches), case.backbuffer)
        black = int(np.all(rgb == 0, axis=2).sum())
        assert 0 < black < rgb.shape[0] * rgb.shape[1], black


def _rpu_mastering(dv):
    """m_DoviMaxMasteringLuminance / m_DoviMinMasteringLuminance as CopySample derives them from the RPU (:2472-2490)"""
    import ctypes as C
    f = O.lib().orc_pq_to_linear_nits
    f.restype = C.c_float; f.argtypes = [C.c_float]
    mx = int(f(np.float32(dv.source_max_pq) / np.float32(4095.0)))
    mn = int(np.float32(f(np.float32(dv.source_min_pq) / np.float32(4095.0))) * np.float32(10000.0))
    for i in range(32):
        if dv.ext[i].level == 6:
            mx, mn = int(dv.ext[i].v[0]), int(dv.ext[i].v[1])
            break
    return mx, mn


@pytest.mark.parametrize("tm_type,l1", [(1, True), (5, True), (3, False), (6, True)])
def test_dolby_vision_on_an_hdr_display_with_local_tone_mapping(tm_type, l1):
    """Dolby Vision -> PQ (no SDR conversion), then ps_hdr10_tonemap fed from Level 1 (+ Level 3) brightness in nits, BT.2390
    replaced by ST 2094-10 (Source/DX11VideoProcessor.cpp:2716-2720), with the L2 trims bound as cbuffer b1 (:3362-3364)"""
    import ctypes as C
    lo = (2081, 2100, 2030, 2120, 2010, 2075); hi = (3388, 2070, 2040, 2090, 2060, 2000)
    dv = H.make_dovi(81, mmr=2, l2=[lo, hi], source_max_pq=3696, l1=(40, 3300, 1500) if l1 else None, l3=(2040, 2100, 2000) if l1 else None)
    if l1:
        nits = (C.c_float * 3)()
        assert O.lib().orc_dovi_l1_nits(C.addressof(dv), nits) == 1
        hdr = dict(out=1, tonemap=1, min=nits[0], max=nits[1], cll=nits[1], fall=nits[2], display=600.0, type=6 if tm_type == 5 else tm_type)
    else:
        # no Level 1: the static HDR10 metadata drives the curve; the stream carries none, so the Display-P3 defaults apply with the
        # mastering luminance the RPU states (source_max_pq / source_min_pq, :2472-2475, :2685-2686)
        mx, mn = _rpu_mastering(dv)
        hdr = dict(out=1, tonemap=1, min=float(mn or 50), max=float(mx or 1000), cll=0.0, fall=0.0, display=600.0, type=tm_type)
    ex = O.exfmt(chroma=5, range_=1, matrix=0, prim=0, trc=0)
    hdr["settings_type"] = tm_type                            # the user's curve choice; the oracle gets what SetHDR10ShaderParams receives
    case = H.Case(O.CF["P010"], 64, 36, 96, 54, ex, backbuffer=O.SURF_RGB10A2, up=O.UPSCALE["Lanczos2"], dovi=dv, hdr=hdr, name="dovi_hdr_tm%d" % tm_type)
    _check(case, "natural", 81)


def test_dolby_vision_level6_and_sticky_extension_blocks():
    """(a) a Level 6 block overrides the mastering luminance derived from source_max_pq (:2477-2490); (b) the L2 trims of an earlier
    frame stay in force for frames that carry no Level 2 block until Flush() (m_DoviExtensionMetadata is only cleared there, :4082)"""
    import copy
    ex = O.exfmt(chroma=5, range_=1, matrix=0, prim=0, trc=0)
    # (a)
    dv = H.make_dovi(83, mmr=0, source_max_pq=3696)
    dv.ext[9].level = 6
    for k, v in enumerate((1200, 40, 900, 300)):
        dv.ext[9].v[k] = v
    hdr = dict(out=1, tonemap=1, min=40.0, max=1200.0, cll=0.0, fall=0.0, display=600.0, type=2, settings_type=2)
    _check(H.Case(O.CF["P010"], 64, 36, 96, 54, ex, backbuffer=O.SURF_RGB10A2, up=O.UPSCALE["Lanczos2"], dovi=dv, hdr=hdr, name="dovi_l6"), "natural", 83)
    # (b) to-SDR chain: frame A carries Level 2, frame B (other curves) does not
    l2 = [(2081, 2100, 2030, 2120, 2010, 2075), (3388, 2070, 2040, 2090, 2060, 2000)]
    dvA = H.make_dovi(84, mmr=1, l2=l2)
    dvB = H.make_dovi(85, mmr=1)
    dvB_with_A = H.make_dovi(85, mmr=1, l2=l2)                     # what the reference renders B with: B's curves, A's trims
    caseB = H.Case(O.CF["P010"], 64, 36, 96, 54, ex, up=O.UPSCALE["Lanczos2"], dovi=dvB, name="dovi_sticky")
    codes = H.make_codes(caseB.cf, caseB.w, caseB.h, "natural", 84)
    buf, pitch, planes, pitches = H.pack(caseB.cf, caseB.w, caseB.h, codes)
    ref_sticky = H.run_oracle(H.Case(O.CF["P010"], 64, 36, 96, 54, ex, up=O.UPSCALE["Lanczos2"], dovi=dvB_with_A, name="x"), planes, pitches)
    ref_plain = H.run_oracle(caseB, planes, pitches)
    assert H.compare(ref_sticky, ref_plain, caseB.backbuffer)[0] > 0             # the trims matter on this frame
    vp = caseB.make_processor(device=0, path=PATH_MULTIPASS)
    vp.ProcessSample(data=buf, pitch=pitch, dovi=dvA)
    vp.ProcessSample(data=buf, pitch=pitch, dovi=dvB)
    assert H.compare(vp.ReadBackbuffer(), ref_sticky, caseB.backbuffer) == (0, 0)
    vp.Flush()
    vp.ProcessSample(data=buf, pitch=pitch, dovi=H.make_dovi(86, mmr=1))         # a different RPU, so that the curves are rebuilt
    vp.ProcessSample(data=buf, pitch=pitch, dovi=dvB)
    assert H.compare(vp.ReadBackbuffer(), ref_plain, caseB.backbuffer) == (0, 0)


def test_dolby_vision_profile4_and_fel_streams_are_ignored():
    """CheckDoviMetadata (Source/VideoProcessor.cpp:273-284): profile 4 (EL spatial resampling, residual enabled, vdr_bit_depth != 12) and
    profile 7 FEL (non-trivial NLQ with the residual enabled) metadata is not taken: the frame renders as if there were none"""
    ex = O.exfmt(chroma=5, range_=1, matrix=0, prim=0, trc=0)
    case = H.Case(O.CF["P010"], 64, 36, 96, 54, ex, up=O.UPSCALE["Lanczos2"], name="dovi_rejected")
    codes = H.make_codes(case.cf, case.w, case.h, "natural", 87)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    ref_none = H.run_oracle(case, planes, pitches)
    for header in (dict(el_spatial_resampling_filter_flag=1, disable_residual_flag=0, vdr_bit_depth=14, nlq=[dict(vdr_in_max=1 << 23)] * 3),
                   dict(disable_residual_flag=0, nlq_method_idc=0, nlq=[dict(nlq_offset=5, vdr_in_max=1 << 23)] * 3)):
        dv = H.make_dovi(87, mmr=0)
        dv.header = header
        vp = case.make_processor(device=0, path=PATH_MULTIPASS)
        vp.ProcessSample(data=buf, pitch=pitch, dovi=dv)
        assert H.compare(vp.ReadBackbuffer(), ref_none, case.backbuffer) == (0, 0), header
    # the same metadata with the residual disabled (profile 5 / 8) IS taken
    dv = H.make_dovi(87, mmr=0)
    vp = case.make_processor(device=0, path=PATH_MULTIPASS)
    vp.ProcessSample(data=buf, pitch=pitch, dovi=dv)
    ref_dv = H.run_oracle(H.Case(O.CF["P010"], 64, 36, 96, 54, ex, up=O.UPSCALE["Lanczos2"], dovi=dv, name="x"), planes, pitches)
    assert H.compare(vp.ReadBackbuffer(), ref_dv, case.backbuffer) == (0, 0)
    assert H.compare(ref_dv, ref_none, case.backbuffer)[0] > 0


# ---- path-B correction shaders as a post-scale step on RGB from an upstream fixed-function converter ----
EX_PQ_TAGS = O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=15)
EX_HLG_TAGS = O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=16)
EX_2020_SDR_TAGS = O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=5)


@pytest.mark.parametrize("name,ex,pq,to_sdr,hdr,mode", [
    ("r210",   EX_PQ_TAGS,       True,  1, None,          1),     # ps_convert_pq_to_sdr
    ("RGB48",  EX_PQ_TAGS,       False, 1, None,          2),     # ps_fixconvert_pq_to_sdr
    ("XRGB32", EX_HLG_TAGS,      False, 1, None,          3),     # ps_fixconvert_hlg_to_sdr
    ("r210",   EX_2020_SDR_TAGS, False, 1, None,          4),     # ps_fix_bt2020
    ("RGB48",  EX_HLG_TAGS,      False, 0, None,          4),     # HLG shown as SDR: ps_fix_bt2020
    ("r210",   EX_HLG_TAGS,      False, 1, dict(out=1),   5),     # HDR display: ps_convert_hlg_to_pq
    ("r210",   EX_PQ_TAGS,       False, 1, dict(out=1),   0),     # HDR display, PQ: passed through
    ("XRGB32", H.EX_709,         False, 1, None,          0),     # nothing to correct
])
def test_correction_shader_selection_and_pixels(name, ex, pq, to_sdr, hdr, mode):
    """m_pPSCorrection choice (Source/DX11VideoProcessor.cpp:1893-1931) and the shader itself, drawn after the resize (:3354-3357)"""
    for (dw, dh, tex) in ((96, 54, O.TEXFMT_AUTOINT), (64, 36, O.TEXFMT_16FLOAT)):        # resized / 1:1 with a final pass behind
        case = H.Case(O.CF[name], 64, 36, dw, dh, ex, up=O.UPSCALE["CatmullRom"], to_sdr=to_sdr, hdr=hdr, tex=tex, correction=mode,
                      upstream_vp=(True, pq), backbuffer=O.SURF_RGB10A2 if hdr else O.SURF_BGRA8, name="corr%d_%s" % (mode, name))
        vp = _check(case, "natural", 90 + mode)
        assert vp.debug_plan().correction == mode


def test_correction_needs_the_upstream_converter_flag():
    plain = H.Case(O.CF["r210"], 64, 36, 96, 54, EX_PQ_TAGS, up=O.UPSCALE["CatmullRom"], name="corr_off")       # tags alone change nothing
    vp = _check(plain, "natural", 99)
    assert vp.debug_plan().correction == 0


# ---- seeded fuzz over whole plans: formats x sizes x rects x settings x orientation, every kernel family vs the oracle ----
def _random_case(rng, i):
    yuv = ["NV12", "P010", "P016", "YUY2", "UYVY", "P210", "Y210", "V210", "AYUV", "Y410", "Y416", "YV12", "YV16", "YV24", "YUV420P10", "YUV422P16", "YUV444P10"]
    rgb = ["RGB24", "XRGB32", "r210", "RGB48", "BGR48", "BGRA64", "B64A", "GBRP8", "GBRP10", "GBRP16", "Y8", "Y10", "Y16"]
    name = str(rng.choice(yuv if rng.random() < 0.7 else rgb))
    is_yuv = name in yuv
    w = int(rng.integers(4, 80)) * 2; h = int(rng.integers(4, 60)) * 2
    exs = [H.EX_709, H.EX_2020_PQ, H.EX_2020_HLG, H.EX_2020_SDR, O.exfmt(chroma=1, range_=1, matrix=2, prim=5, trc=4), O.exfmt(chroma=7, range_=2, matrix=7, prim=2, trc=5)]
    ex = exs[int(rng.integers(0, len(exs)))] if is_yuv else 0
    # source rect: whole frame or an even-aligned crop
    src_rect = None
    if rng.random() < 0.4:
        l = int(rng.integers(0, w // 4)) * 2; t = int(rng.integers(0, h // 4)) * 2
        r = w - int(rng.integers(0, w // 4)) * 2; b = h - int(rng.integers(0, h // 4)) * 2
        if r - l >= 4 and b - t >= 4:
            src_rect = (l, t, r, b)
    sw, sh = (src_rect[2] - src_rect[0], src_rect[3] - src_rect[1]) if src_rect else (w, h)
    rot = int(rng.choice([0, 0, 0, 90, 180, 270])); flip = int(rng.random() < 0.15)
    up = str(rng.choice(["Nearest", "Mitchell", "CatmullRom", "Lanczos2", "Lanczos3", "Jinc2"]))
    if (rot or flip) and up == "Jinc2":
        up = "Lanczos2"
    mode = rng.random()
    tw, th = (sh, sw) if rot in (90, 270) else (sw, sh)
    if mode < 0.2:
        dw, dh = tw, th
    elif mode < 0.4:
        dw, dh = 2 * tw, 2 * th
    else:
        dw = max(2, int(tw * rng.uniform(0.3, 2.6))); dh = max(2, int(th * rng.uniform(0.3, 2.6)))
    win = (dw + int(rng.integers(0, 20)), dh + int(rng.integers(0, 20)))
    ox, oy = int(rng.integers(-6, win[0] - dw + 7)), int(rng.integers(-6, win[1] - dh + 7))
    dst_rect = (ox, oy, ox + dw, oy + dh) if rng.random() < 0.5 else (0, 0, dw, dh)
    if dst_rect == (0, 0, dw, dh):
        win = (dw, dh)
    tex = int(rng.choice([O.TEXFMT_AUTOINT, O.TEXFMT_8INT, O.TEXFMT_10INT, O.TEXFMT_16FLOAT]))
    bb = int(rng.choice([O.SURF_BGRA8, O.SURF_BGRA8, O.SURF_RGB10A2, O.SURF_RGBA16F]))
    procamp = dict(brightness=float(rng.uniform(-20, 20)), contrast=float(rng.uniform(0.7, 1.3)), hue=float(rng.uniform(-30, 30)),
                   saturation=float(rng.uniform(0.6, 1.4))) if rng.random() < 0.3 else None
    return H.Case(O.CF[name], w, h, dw, dh, ex, backbuffer=bb, tex=tex, chroma=int(rng.integers(0, 3)), up=O.UPSCALE[up],
                  down=int(rng.integers(0, 6)), interp50=int(rng.random() < 0.7), dither=int(rng.random() < 0.7), to_sdr=int(rng.random() < 0.8),
                  nits=int(rng.integers(5, 80)) * 5, src_rect=src_rect, dst_rect=dst_rect, win=win, variant=int(rng.random() < 0.2), procamp=procamp,
                  rotation=rot, flip=flip, name="fuzz%d_%s" % (i, name))


@pytest.mark.parametrize("block", range(6))
def test_fuzzed_plans(block):
    rng = np.random.Generator(np.random.PCG64(1000 + block))
    for i in range(20):
        case = _random_case(rng, 20 * block + i)
        try:
            _check(case, str(rng.choice(["fullnoise", "natural", "ramp"])), seed=3000 + 20 * block + i)
        except AssertionError as e:
            raise AssertionError("%s  [%s]" % (e, {k: v for k, v in case.__dict__.items() if k not in ("dovi", "hdr")}))


def test_extreme_downscale_long_convolutions():
    """hundreds of taps per output (the convolution support grows with the ratio): table-driven passes, no tap limit in practice"""
    _check(H.Case(O.CF["NV12"], 2048, 64, 16, 8, H.EX_709, down=O.DOWNSCALE["Lanczos"], name="down128x"), "natural", 101)
    _check(H.Case(O.CF["NV12"], 64, 1024, 40, 6, H.EX_709, down=O.DOWNSCALE["Bicubic"], name="down170y"), "fullnoise", 102)


def test_convert_pass_two_lane_equals_scalar():
    """k_convert2 (two pixels per thread in the packed-fp32 lanes) against k_convert, every transfer chain, odd widths"""
    import videorenderer_b200 as V
    exs = [H.EX_709, H.EX_2020_PQ, H.EX_2020_HLG, H.EX_2020_SDR, O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=14)]
    try:
        for i, ex in enumerate(exs):
            for hdr in (None, dict(out=1)):
                case = H.Case(O.CF["P010"], 66, 34, 66, 34, ex, up=O.UPSCALE["Nearest"], src_rect=(2, 0, 65, 33), dst_rect=(0, 0, 63, 33), win=(63, 33),
                              hdr=hdr, backbuffer=O.SURF_RGBA16F, tex=O.TEXFMT_16FLOAT, dither=0, name="twolane%d" % i)
                codes = H.make_codes(case.cf, case.w, case.h, "fullnoise", 200 + i)
                buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
                outs = []
                for on in (1, 0):
                    vp = case.make_processor(device=0, path=PATH_MULTIPASS)
                    vp.debug_set_convert_two_lane(on)
                    vp.ProcessSample(data=buf, pitch=pitch)
                    outs.append(vp.ReadBackbuffer())
                assert np.array_equal(outs[0], outs[1]), (i, hdr)
                assert H.compare(outs[0], H.run_oracle(case, planes, pitches), case.backbuffer) == (0, 0)
    finally:
        vp.debug_set_convert_two_lane(1)


@pytest.mark.parametrize("cfname,w,h,rect", [("NV12", 150, 38, None), ("P010", 150, 38, None), ("P016", 136, 22, (16, 2, 129, 21)), ("P210", 72, 20, None), ("NV12", 96, 18, (32, 0, 96, 18)),
                                             ("YUY2", 150, 22, None), ("UYVY", 136, 20, (8, 2, 129, 19)), ("YUY2", 64, 10, (56, 0, 64, 10))])
def test_convert_pass_span_kernel_equals_two_lane_and_scalar(cfname, w, h, rect):
    """k_convert_bp (128-bit source spans staged through shared memory, then the two-lane body) against k_convert2 and k_convert:
    ragged last spans, cropped rects on span boundaries, every chroma siting, the three internal formats, transfer chains"""
    exs = [H.EX_709, H.EX_2020_PQ, H.EX_2020_SDR, O.exfmt(chroma=1, range_=2, matrix=1, prim=2, trc=5), O.exfmt(chroma=3, range_=1, matrix=2, prim=2, trc=5)]
    try:
        for i, ex in enumerate(exs):
            for tex, bb in ((O.TEXFMT_10INT, O.SURF_BGRA8), (O.TEXFMT_16FLOAT, O.SURF_RGBA16F), (O.TEXFMT_8INT, O.SURF_RGB10A2)):
                dw, dh = (rect[2] - rect[0], rect[3] - rect[1]) if rect else (w, h)
                case = H.Case(O.CF[cfname], w, h, dw * 3 // 2, dh * 3 // 2, ex, up=O.UPSCALE["CatmullRom"], src_rect=rect, tex=tex, backbuffer=bb, name="span%d" % i)
                codes = H.make_codes(case.cf, case.w, case.h, "fullnoise", 400 + i)
                buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
                outs = []
                for mode in (1, 2, 0):
                    vp = case.make_processor(device=0, path=PATH_MULTIPASS)
                    vp.debug_set_convert_two_lane(mode)
                    vp.ProcessSample(data=buf, pitch=pitch)
                    outs.append(vp.ReadBackbuffer())
                assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2]), (i, tex)
                assert H.compare(outs[0], H.run_oracle(case, planes, pitches), case.backbuffer) == (0, 0)
    finally:
        vp.debug_set_convert_two_lane(1)


@pytest.mark.parametrize("cfg", [("Lanczos3", 0, 133, 77), ("CatmullRom", 0, 96, 54), ("Lanczos3", 1, 133, 77), (None, 0, 50, 31), (None, 1, 23, 90), (None, 0, 31, 17)])
def test_resize_four_outputs_per_thread_equals_one(cfg):
    """k_resize_p and k_resize4 against k_resize: interpolation (incl. the D3D11 duplicate tap), convolution, mixed axes, clipped rects, every surface format"""
    up, variant, dw, dh = cfg
    for tex, bb in ((O.TEXFMT_10INT, O.SURF_BGRA8), (O.TEXFMT_16FLOAT, O.SURF_RGBA16F), (O.TEXFMT_8INT, O.SURF_RGB10A2)):
        case = H.Case(O.CF["NV12"], 62, 34, dw, dh, H.EX_709, up=O.UPSCALE[up or "Mitchell"], down=O.DOWNSCALE["Lanczos"], variant=variant, tex=tex, backbuffer=bb,
                      dst_rect=(-5, 3, dw - 5, dh + 3), win=(dw, dh), interp50=0, name="resize4")
        codes = H.make_codes(case.cf, case.w, case.h, "fullnoise", 300)
        buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
        outs = []
        try:
            for mode in (1, 2, 0):                 # k_resize_p (compile-time taps / formats, two lines per thread), k_resize4, k_resize
                vp = case.make_processor(device=0, path=PATH_MULTIPASS)
                vp.debug_set_resize_four(mode)
                vp.ProcessSample(data=buf, pitch=pitch)
                outs.append(vp.ReadBackbuffer())
        finally:
            vp.debug_set_resize_four(1)
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
        assert H.compare(outs[0], H.run_oracle(case, planes, pitches), case.backbuffer) == (0, 0)


# ------------------------------------------------------------------------------------------------
# persistent warp-specialised 2x kernel (kernel_strip.cuh): TMA-staged source rows, role pipeline P -> A -> B -> C
# ------------------------------------------------------------------------------------------------
def _strip_cases():
    P = O.CF
    cs = []
    # strip geometry: widths around the 128-pixel strip, heights around the 4-row quad / 8-row group, tiny frames (fewer quads than SMs)
    for (w, h) in ((150, 70), (128, 64), (126, 62), (130, 66), (258, 38), (2, 2), (10, 130), (386, 30), (512, 300), (640, 42)):
        cs.append(H.Case(P["P010"], w, h, 2 * w, 2 * h, H.EX_2020_PQ, up=O.UPSCALE["Lanczos3"], name="strip_p010_%dx%d" % (w, h)))
    cs.append(H.Case(P["NV12"], 96, 64, 160, 104, H.EX_709, src_rect=(8, 6, 88, 58), up=O.UPSCALE["Lanczos3"], name="strip_nv12_crop"))
    cs.append(H.Case(P["NV12"], 96, 64, 2 * 77, 2 * 51, H.EX_709, src_rect=(3, 5, 80, 56), up=O.UPSCALE["CatmullRom"], name="strip_nv12_oddcrop"))
    cs.append(H.Case(P["P010"], 64, 36, 128, 72, H.EX_2020_PQ, dst_rect=(16, 11, 144, 83), win=(160, 100), up=O.UPSCALE["Lanczos3"], name="strip_letterbox"))
    cs.append(H.Case(P["P210"], 200, 50, 400, 100, H.EX_2020_PQ, up=O.UPSCALE["Lanczos2"], name="strip_p210"))
    cs.append(H.Case(P["P216"], 136, 40, 272, 80, H.EX_2020_SDR, up=O.UPSCALE["Mitchell"], name="strip_p216_sdr2020"))
    cs.append(H.Case(P["P016"], 140, 44, 280, 88, H.EX_2020_HLG, up=O.UPSCALE["Mitchell"], name="strip_p016_hlg"))
    for siting in (1, 5, 7):
        ex = O.exfmt(chroma=siting, range_=2, matrix=1, prim=2, trc=5)
        cs.append(H.Case(P["NV12"], 140, 52, 280, 104, ex, up=O.UPSCALE["Lanczos3"], variant=O.VARIANT_D3D9, name="strip_nv12_siting%d" % siting))
    cs.append(H.Case(P["P010"], 132, 34, 264, 68, H.EX_2020_PQ, backbuffer=O.SURF_RGB10A2, tex=O.TEXFMT_16FLOAT, up=O.UPSCALE["Lanczos3"], name="strip_16f_rgb10"))
    cs.append(H.Case(P["P010"], 132, 34, 264, 68, H.EX_2020_PQ, backbuffer=O.SURF_RGBA16F, tex=O.TEXFMT_16FLOAT, dither=0, up=O.UPSCALE["Lanczos3"], name="strip_16f_16f"))
    cs.append(H.Case(P["P010"], 132, 34, 264, 68, H.EX_2020_PQ, tex=O.TEXFMT_8INT, dither=0, up=O.UPSCALE["Lanczos3"], name="strip_8int"))
    cs.append(H.Case(P["P010"], 132, 34, 264, 68, H.EX_2020_PQ, nits=400, procamp=dict(brightness=12.0, contrast=1.1, hue=10.0, saturation=0.9), up=O.UPSCALE["Lanczos3"], name="strip_procamp"))
    return cs


def _run_strip(case, buf, pitch, enable, cfg=0):
    vp = case.make_processor(device=0, path=PATH_FUSED)
    vp.debug_strip_status(clear=True)
    vp.debug_set_strip_kernel(enable, cfg)
    try:
        vp.ProcessSample(data=buf, pitch=pitch)
        out = vp.ReadBackbuffer()
    finally:
        vp.debug_set_strip_kernel(1, 0)
    st = vp.debug_strip_status(clear=True)
    assert st[0] == 0, "strip kernel watchdog: %d mbarrier waits timed out (slot %d parity %d iteration %d block %d warp %d)" % tuple(st[:6])
    return vp, out


@pytest.mark.parametrize("case", _strip_cases(), ids=lambda c: c.name)
def test_strip_kernel_equals_tile_kernel_and_oracle(case):
    codes = H.make_codes(case.cf, case.w, case.h, "fullnoise", 5)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    ref = H.run_oracle(case, planes, pitches)
    _, a = _run_strip(case, buf, pitch, 1)
    _, b = _run_strip(case, buf, pitch, 0)
    assert H.compare(a, ref, case.backbuffer) == (0, 0)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("cfg", [1, 2, 3, 6, 7, 8, 9, 10, 11, 12, 13, 14])
def test_strip_kernel_role_splits_give_the_same_frame(cfg):
    """every selectable configuration of the strip kernel (kernel_strip.cu: warp splits 1-3; 6 = gamma through T[] / D[] with the
    read-back quotient formed; 7 = 32 warps; 8 = single-lookup gamma only; 9, 10 = setmaxnreg register budgets; 11 = table read-back in
    role C; 12 = column-per-thread role P; 13 = rolled channel loop; 14 = the 28-warp kernel) renders the oracle's frame -- the measured-and-rejected variants of
    DESIGN.md 4.1 stay honest"""
    case = H.baseline_cases(scale=0.25)[1]
    codes = H.make_codes(case.cf, case.w, case.h, "natural", 77)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    ref = H.run_oracle(case, planes, pitches)
    _, out = _run_strip(case, buf, pitch, 1, cfg)
    assert H.compare(out, ref, case.backbuffer) == (0, 0)


def test_strip_kernel_falls_back_for_unaligned_device_planes():
    """device samples whose planes / pitches are not 16-byte aligned cannot be addressed by a TMA tensor map: the tile kernel takes them"""
    import torch
    case = H.Case(O.CF["P010"], 150, 70, 300, 140, H.EX_2020_PQ, up=O.UPSCALE["Lanczos3"], name="unaligned")
    codes = H.make_codes(case.cf, case.w, case.h, "noise", 9)
    pitch = 150 * 2 + 8                                        # 308 bytes: a multiple of the 4-byte UV texel, not of 16
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes, pitch=pitch)
    ref = H.run_oracle(case, planes, pitches)
    t = torch.from_numpy(buf).cuda()
    vp = case.make_processor(device=0, path=PATH_FUSED)
    vp.ProcessSample(device_planes=[t.data_ptr(), t.data_ptr() + pitch * case.h], device_pitches=[pitch, pitch])
    torch.cuda.synchronize()
    assert H.compare(vp.ReadBackbuffer(), ref, case.backbuffer) == (0, 0)
    # a pitch that is not a multiple of the chroma texel cannot be read with typed loads at all: refused, not a device fault
    import videorenderer_b200 as V
    hr = vp.ProcessSample(device_planes=[t.data_ptr(), t.data_ptr() + 306 * case.h], device_pitches=[306, 306], check=False)
    assert hr == V.HRESULT.E_INVALIDARG


def test_two_instances_on_two_streams_render_their_first_frames_concurrently():
    """round-1 advisor finding: per-device tables must be ready for an instance that never rendered before, whatever another
    instance on another stream is doing (the dither table is a compile-time constant now; the gamma table build is serialised)"""
    import threading
    import torch
    case = H.baseline_cases(scale=0.25)[1]
    codes = H.make_codes(case.cf, case.w, case.h, "natural", 78)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    ref = H.run_oracle(case, planes, pitches)
    outs = [None, None]

    def work(i):
        torch.cuda.set_device(0)
        s = torch.cuda.Stream()
        import videorenderer_b200 as V
        vp = V.VideoProcessor(case.settings(), device=0, stream=s.cuda_stream)
        mt = V.VideoProcessor.media_type(case.cf, case.w, case.h, case.ex)
        assert vp.InitMediaType(mt)
        vp.SetWindowRect((0, 0, case.dst_w, case.dst_h)); vp.SetVideoRect((0, 0, case.dst_w, case.dst_h))
        vp.debug_set_path(PATH_FUSED if i == 0 else PATH_AUTO)
        vp.ProcessSample(data=buf, pitch=pitch)
        outs[i] = vp.ReadBackbuffer()

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for o in outs:
        assert o is not None and H.compare(o, ref, case.backbuffer) == (0, 0)


# ------------------------------------------------------------------------------------------------
# the four single-GPU BASELINE configs at the sizes BASELINE.json names, every pixel of the frame (all four frame edges included)
# ------------------------------------------------------------------------------------------------
def _frame_hash(out, fmt_):
    import hashlib
    rgb = np.ascontiguousarray(H.rgb_view(out, fmt_))
    if rgb.dtype == np.float16:
        rgb = rgb.view(np.uint16)
    return hashlib.sha256(rgb.tobytes()).hexdigest()


@pytest.mark.parametrize("idx", range(4))
def test_full_size_baseline_configs_equal_the_oracle_on_every_pixel(idx):
    """C1 1920x1080 NV12; C2 3840x2160 P010 -> 7680x4320; C3 1920x1080 YUY2 -> 2880x1620 RGB10A2; C4 7680x4320 P016 -> 3840x2160 RGBA16F:
    the oracle renders the whole frame on the host cores and every kernel family must equal it on every pixel; the oracle's own output
    is pinned by the committed hash (tests/golden/frames_full.json, minted by tests/golden/make_golden_fullsize.py)."""
    import hashlib
    import json
    import os
    case = H.baseline_cases(1.0)[idx]
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "frames_full.json")))[case.name]
    codes = H.make_codes(case.cf, case.w, case.h, golden["kind"], golden["seed"])
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    assert hashlib.sha256(buf.tobytes()).hexdigest() == golden["in_sha256"]
    ref = H.run_oracle(case, planes, pitches)
    assert _frame_hash(ref, case.backbuffer) == golden["out_rgb_sha256"], "the oracle's full-size frame drifted from the committed hash"
    for p in (PATH_MULTIPASS, PATH_AUTO, PATH_FUSED):
        vp, out = _run_gpu(case, buf, pitch, p)
        n, d = H.compare(out, ref, case.backbuffer)
        assert n == 0, "%s path %d(%d): %d channel values differ (max delta %d)" % (case.name, p, vp.debug_plan().path, n, d)
        assert _frame_hash(out, case.backbuffer) == golden["out_rgb_sha256"]
    if idx == 1:                                   # C2 also through the tile kernel (the strip kernel's fallback)
        _, out = _run_strip(case, buf, pitch, 0)
        assert H.compare(out, ref, case.backbuffer) == (0, 0)


# ---- MPCVR_SURF_RGB8 (extension): the B8G8R8A8 codes as packed R,G,B bytes -----------------------------------------------------
def _rgb8_pair(case_kw, cf, w, h, dw, dh, ex, **kw):
    a = H.Case(O.CF[cf], w, h, dw, dh, ex, backbuffer=O.SURF_BGRA8, name="bgra", **kw)
    b = H.Case(O.CF[cf], w, h, dw, dh, ex, backbuffer=O.SURF_RGB8, name="rgb8", **kw)
    return a, b


@pytest.mark.parametrize("geom", [("P010", 150, 70, 300, 140, "pq", {}), ("P010", 131, 35, 262, 70, "pq", {}), ("P010", 64, 36, 128, 72, "pq", dict(dst_rect=(16, 11, 144, 83), win=(160, 100))),
                                  ("P010", 64, 36, 128, 72, "pq", dict(dst_rect=(18, 11, 146, 83), win=(164, 100))), ("NV12", 96, 64, 192, 128, "709", {}),
                                  ("NV12", 100, 60, 100, 60, "709", {}), ("YUY2", 96, 54, 144, 81, "hlg", dict(tex=O.TEXFMT_16FLOAT)), ("P016", 120, 68, 60, 34, "sdr2020", dict(tex=O.TEXFMT_10INT))],
                         ids=lambda g: "%s_%dx%d_%dx%d" % g[:5])
def test_rgb8_backbuffer_holds_the_bgra8_codes(geom):
    cf, w, h, dw, dh, exn, kw = geom
    ex = {"pq": H.EX_2020_PQ, "709": H.EX_709, "hlg": H.EX_2020_HLG, "sdr2020": H.EX_2020_SDR}[exn]
    kw = dict(kw); kw.setdefault("up", O.UPSCALE["Lanczos3"])
    win = kw.get("win", (dw, dh))
    a = H.Case(O.CF[cf], w, h, dw, dh, ex, backbuffer=O.SURF_BGRA8, name="bgra", **kw)
    b = H.Case(O.CF[cf], w, h, dw, dh, ex, backbuffer=O.SURF_RGB8, name="rgb8", **kw)
    codes = H.make_codes(a.cf, w, h, "fullnoise", 11)
    buf, pitch, planes, pitches = H.pack(a.cf, w, h, codes)
    ref = H.rgb_view(H.run_oracle(a, planes, pitches), O.SURF_BGRA8)           # the oracle's B8G8R8A8 frame
    for p in (PATH_MULTIPASS, PATH_AUTO):
        vp, out = _run_gpu(b, buf, pitch, p)
        assert out.shape == (win[1], win[0] * 3)
        got = H.rgb_view(out, O.SURF_RGB8)
        assert np.array_equal(got, ref), "path %d(%d): %d values differ" % (p, vp.debug_plan().path, int((got != ref).sum()))
        st = vp.debug_strip_status(clear=True)
        assert st[0] == 0


def test_rgb8_full_size_c2_equals_the_bgra8_frame():
    case = H.baseline_cases(1.0)[1]
    codes = H.make_codes(case.cf, case.w, case.h, "natural", 21)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    _, bgra = _run_gpu(case, buf, pitch, PATH_AUTO)
    c8 = H.Case(case.cf, case.w, case.h, case.dst_w, case.dst_h, case.ex, backbuffer=O.SURF_RGB8, up=case.up, name="c2_rgb8")
    vp, rgb = _run_gpu(c8, buf, pitch, PATH_AUTO)
    assert vp.debug_plan().path == PATH_FUSED and vp.LastFrameKernels() == 1
    assert np.array_equal(H.rgb_view(rgb, O.SURF_RGB8), H.rgb_view(bgra, O.SURF_BGRA8))


# ---- 1:1 plans of bi-planar sources: the 128-bit span kernel (k_fused1x_bp) against the one-pixel-per-thread kernel and the oracle ----
@pytest.mark.parametrize("name,ex", [("NV12", "709"), ("P010", "pq"), ("P016", "hlg"), ("P210", "sdr2020"), ("P216", "709")])
@pytest.mark.parametrize("siting", [1, 5, 7])
def test_fused_1x_span_kernel_equals_the_pixel_kernel(name, ex, siting):
    base = {"709": dict(range_=2, matrix=1, prim=2, trc=5), "pq": dict(range_=2, matrix=4, prim=9, trc=15),
            "hlg": dict(range_=2, matrix=4, prim=9, trc=16), "sdr2020": dict(range_=2, matrix=4, prim=9, trc=5)}[ex]
    exf = O.exfmt(chroma=siting, **base)
    geoms = [dict(w=160, h=46, src_rect=None, dst_rect=None, win=None),                    # aligned spans only
             dict(w=150, h=38, src_rect=None, dst_rect=None, win=None),                    # ragged last span
             dict(w=192, h=40, src_rect=(32, 4, 176, 38), dst_rect=None, win=None),        # crop on a span boundary (16 | 32)
             dict(w=192, h=40, src_rect=(6, 3, 150, 37), dst_rect=None, win=None),         # crop off the boundary: pixel kernel
             dict(w=160, h=46, src_rect=None, dst_rect=(8, 5, 168, 51), win=(180, 60))]    # offset inside a larger window
    for g in geoms:
        sr = g["src_rect"] or (0, 0, g["w"], g["h"])
        dw, dh = sr[2] - sr[0], sr[3] - sr[1]
        for tex, bb, dither in ((O.TEXFMT_AUTOINT, O.SURF_BGRA8, 1), (O.TEXFMT_16FLOAT, O.SURF_RGB10A2, 1), (O.TEXFMT_16FLOAT, O.SURF_RGBA16F, 0)):
            case = H.Case(O.CF[name], g["w"], g["h"], dw, dh, exf, backbuffer=bb, tex=tex, dither=dither, src_rect=g["src_rect"],
                          dst_rect=g["dst_rect"], win=g["win"], name="span1x")
            codes = H.make_codes(case.cf, case.w, case.h, "fullnoise", 13)
            buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
            ref = H.run_oracle(case, planes, pitches)
            outs = []
            try:
                for on in (1, 0, 2):                    # span kernel (8 pixels per thread), pixel kernel, span kernel with 16 pixels per thread
                    vp = case.make_processor(device=0, path=PATH_FUSED)
                    vp.debug_set_fused1x_vector(on)
                    vp.ProcessSample(data=buf, pitch=pitch)
                    outs.append(vp.ReadBackbuffer())
            finally:
                vp.debug_set_fused1x_vector(1)
            assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2]), (name, siting, g, tex, bb)
            assert H.compare(outs[0], ref, bb) == (0, 0), (name, siting, g, tex, bb)
