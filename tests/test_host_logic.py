"""Host-side logic of the product through the C ABI on a planning-only instance (no GPU needed):
colour matrices, resize tables, pass-chain decisions, media-type rules -- against the oracle and the goldens."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import harness as H
import orclib as O
import videorenderer_b200 as V

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return [int(v) for v in np.asarray(a, np.float32).ravel().view(np.uint32)]


def _vp(case):
    return case.make_processor(device=-1)


def test_settings_default_matches_reference_SetDefault():
    s = V.Settings.default()
    # Settings_t::SetDefault, Source/IVideoRenderer.h:140-185 (Windows 10 branch)
    exp = dict(bUseD3D11=1, bShowStats=0, iResizeStats=0, iTexFormat=0, VPFmts_bNV12=1, VPFmts_bP01x=1, VPFmts_bYUY2=1, VPFmts_bOther=1,
               iVPDeinterlacing=1, bDeintDouble=1, bVPScaling=1, iVPSuperRes=0, bVPRTXVideoHDR=0, iChromaScaling=1, iUpscaling=2, iDownscaling=2,
               bInterpolateAt50pct=1, bUseDither=1, bDeintBlend=0, iSwapEffect=1, bExclusiveFS=0, bVBlankBeforePresent=0, bAdjustPresentTime=1,
               bReinitByDisplay=0, bHdrPreferDoVi=0, bHdrPassthrough=1, iHdrToggleDisplay=0, iHdrOsdBrightness=0, bConvertToSdr=1,
               iSDRDisplayNits=125, bHdrLocalToneMapping=0, iHdrLocalToneMappingType=1, iHdrDisplayMaxNits=1000)
    assert {n: getattr(s, n) for n, _ in s._fields_} == exp


def test_color_transform_bits_equal_reference_golden_and_oracle():
    kat = json.load(open(os.path.join(G, "kat_matrices.json")))
    cases = {"bt709_tv_8": (O.CF["NV12"], H.EX_709), "bt2020nc_tv_16": (O.CF["P010"], H.EX_2020_PQ),
             "bt601_tv_8": (O.CF["YUY2"], O.exfmt(range_=2, matrix=2, prim=2, trc=5)),
             "bt709_pc_8": (O.CF["NV12"], O.exfmt(chroma=5, range_=1, matrix=1, prim=2, trc=5)),
             "bt2020nc_tv_10": (O.CF["YUV420P10"], H.EX_2020_PQ), "ycgco_pc_8": (O.CF["YUV444P8"], O.exfmt(range_=1, matrix=7, prim=2, trc=5)),
             "smpte240m_tv_8": (O.CF["NV12"], O.exfmt(chroma=5, range_=2, matrix=3, prim=2, trc=5))}
    for name, (cf, ex) in cases.items():
        vp = _vp(H.Case(cf, 64, 36, 64, 36, ex))
        m, c, g, lum = vp.debug_color_params()
        assert _bits(m) == kat["matrices"][name]["m_bits"], name
        assert _bits(c) == kat["matrices"][name]["c_bits"], name
        assert _bits(g) == kat["gamut_2020_to_709_bits"]
        assert lum == 80.0
    for name, e in kat["procamp_bt709_tv_8"].items():
        vp = _vp(H.Case(O.CF["NV12"], 64, 36, 64, 36, H.EX_709, procamp=dict(brightness=e["brightness"], contrast=e["contrast"], hue=e["hue"], saturation=e["saturation"])))
        m, c, _, _ = vp.debug_color_params()
        assert _bits(m) == e["m_bits"] and _bits(c) == e["c_bits"], name
    # random procamp against the oracle restatement
    rng = np.random.Generator(np.random.PCG64(2))
    for _ in range(20):
        pa = dict(brightness=float(rng.uniform(-100, 100)), contrast=float(rng.uniform(0, 2)), hue=float(rng.uniform(-180, 180)), saturation=float(rng.uniform(0, 2)))
        vp = _vp(H.Case(O.CF["P010"], 64, 36, 64, 36, H.EX_2020_PQ, procamp=pa))
        m, c, _, _ = vp.debug_color_params()
        om, oc = O.csp_matrix(H.EX_2020_PQ, 16, **{k: np.float32(v) for k, v in pa.items()})
        assert _bits(m) == _bits(om) and _bits(c) == _bits(oc)


def test_luminance_scale_follows_nits():
    for nits in (25, 100, 125, 400):
        vp = _vp(H.Case(O.CF["P010"], 64, 36, 64, 36, H.EX_2020_PQ, nits=nits))
        assert vp.debug_color_params()[3] == np.float32(10000.0) / np.float32(nits)


SCALE_CASES = [(1920, 2880, "CatmullRom", "Hamming", 1), (3840, 7680, "Lanczos3", "Hamming", 1), (100, 237, "Mitchell", "Box", 1),
               (100, 233, "Lanczos2", "Box", 1), (7680, 3840, "Lanczos3", "Bicubic", 0), (1000, 333, "CatmullRom", "Lanczos", 1),
               (500, 200, "CatmullRom", "Hamming", 1), (400, 100, "CatmullRom", "Box", 1), (640, 300, "CatmullRom", "Bilinear", 1),
               (640, 320, "Lanczos3", "BicubicSharp", 1), (97, 97, "Lanczos3", "Hamming", 1)]


@pytest.mark.parametrize("src,dst,up,down,i50", SCALE_CASES)
@pytest.mark.parametrize("variant", [O.VARIANT_D3D11, O.VARIANT_D3D9])
def test_axis_tables_equal_oracle(src, dst, up, down, i50, variant):
    case = H.Case(O.CF["NV12"], src, 64, dst, 64, H.EX_709, up=O.UPSCALE[up], down=O.DOWNSCALE[down], interp50=i50, variant=variant)
    vp = _vp(case)
    plan = vp.debug_plan()
    kind = O.lib().orc_scaler_kind(src, dst, i50)
    assert plan.scaler_x == kind and plan.scaler_y == 0
    if kind == 0:
        return
    filt = O.DOWNSCALE[down] if kind == 2 else O.UPSCALE[up]
    for i in list(range(0, min(dst, 40))) + [dst // 2, dst - 2, dst - 1]:
        base, w, wsum = vp.debug_axis_taps(0, i)
        ob, ow, osum, n = O.axis_taps(kind, filt, variant, 0, src, dst, i)
        assert base == ob
        assert _bits(w[:n]) == _bits(ow), (i, w, ow)
        assert np.all(w[n:] == 0)
        if kind == 2:
            assert _bits([wsum]) == _bits([osum])


def test_plan_decisions_follow_reference_rules():
    L = O.lib()
    for cf, depth in ((O.CF["NV12"], 8), (O.CF["P010"], 16), (O.CF["YUV420P10"], 10)):
        for tex in (O.TEXFMT_AUTOINT, O.TEXFMT_8INT, O.TEXFMT_10INT, O.TEXFMT_16FLOAT):
            for bb in (O.SURF_BGRA8, O.SURF_RGB10A2, O.SURF_RGBA16F):
                for dither in (0, 1):
                    vp = _vp(H.Case(cf, 64, 36, 128, 72, H.EX_709, backbuffer=bb, tex=tex, dither=dither))
                    p = vp.debug_plan()
                    assert p.internal_fmt == L.orc_internal_fmt(tex, depth)
                    assert p.final_pass == L.orc_need_final_pass(bb, p.internal_fmt, dither)
    # transfer selection (Source/Shaders.cpp:613-616, :892-905)
    def tr(ex, to_sdr=1):
        return _vp(H.Case(O.CF["P010"], 64, 36, 64, 36, ex, to_sdr=to_sdr)).debug_plan().transfer
    assert tr(H.EX_2020_PQ) == 1 and tr(H.EX_2020_PQ, 0) == 0
    assert tr(O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=16)) == 2 and tr(O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=16), 0) == 3
    assert tr(H.EX_2020_SDR) == 3 and tr(H.EX_709) == 0
    assert tr(O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=1)) == 0       # linear transfer: quirk B4 -> matrix only
    assert tr(O.exfmt(chroma=5, range_=2, matrix=1, prim=2, trc=15)) == 1      # PQ with 709 primaries still tone-maps
    # Nearest upscaling has no shader: point-sampled axis
    p = _vp(H.Case(O.CF["NV12"], 64, 36, 128, 72, H.EX_709, up=O.UPSCALE["Nearest"])).debug_plan()
    assert p.scaler_x == 3 and p.scaler_y == 3
    # variant follows Settings_t::bUseD3D11
    assert _vp(H.Case(O.CF["NV12"], 64, 36, 128, 72, H.EX_709, variant=O.VARIANT_D3D9)).debug_plan().variant == 1


def test_media_type_rules():
    vp = V.VideoProcessor(device=-1)
    mt = V.VideoProcessor.media_type(V.CF.NV12, 1918, 1080)
    assert vp.InitMediaType(mt)
    p = vp.debug_plan()
    assert p.src_pitch == 1920 and p.src_lines == 1620                     # NV12 pitch ALIGN 4, lines = H*3/2 (:1789-1796)
    assert p.src_exfmt == O.lib().orc_specify_extended_format(0, 0, 420, 1918, 1080)
    assert vp.GetSourceRect() == (0, 0, 1918, 1080) and vp.GetVideoSize() == (1918, 1080) and vp.GetAspectRatio() == (959, 540)
    mt = V.VideoProcessor.media_type(V.CF.P010, 3840, 2160, H.EX_2020_PQ, aspect=(16, 9))
    assert vp.InitMediaType(mt)
    p = vp.debug_plan()
    assert p.src_pitch == 7680 and p.src_lines == 3240
    assert (p.src_exfmt & 0xFF) == 3                                       # AMCONTROL_USED | AMCONTROL_COLORINFO_PRESENT
    assert vp.GetAspectRatio() == (16, 9)
    # SD default matrix is BT.601, HD BT.709 (SpecifyExtendedFormat)
    assert vp.InitMediaType(V.VideoProcessor.media_type(V.CF.YUY2, 720, 576))
    assert (vp.debug_plan().src_exfmt >> 15) & 7 == 2
    # VR_Extradata: padded buffer width, true frame size
    mt = V.VideoProcessor.media_type(V.CF.NV12, 2048, 1080, frame_size=(1920, 1080))
    assert vp.InitMediaType(mt)
    assert vp.GetSourceRect() == (0, 0, 1920, 1080) and vp.debug_plan().src_pitch == 2048
    # rejects
    # RGB / v210 samples: pitch rules of InitMediaType (:1789-1803): RGB24 rows align to 4, v210 to 128, BI_RGB bottom-up negates
    assert vp.InitMediaType(V.VideoProcessor.media_type(V.CF.RGB24, 62, 64, bottom_up=1))
    assert vp.debug_plan().src_pitch == -188 and vp.debug_plan().src_lines == 64
    assert vp.InitMediaType(V.VideoProcessor.media_type(V.CF.V210, 1280, 64))
    assert vp.debug_plan().src_pitch == (1280 + 5) // 6 * 16 // 128 * 128 + 128
    assert vp.InitMediaType(V.VideoProcessor.media_type(V.CF.BGR48, 63, 64)) and vp.debug_plan().src_pitch == 380
    assert not vp.VerifyMediaType(V.VideoProcessor.media_type(V.CF.NONE, 64, 64))
    assert not vp.VerifyMediaType(V.VideoProcessor.media_type(V.CF.NV12, 0, 64))
    assert not vp.InitMediaType(V.VideoProcessor.media_type(V.CF.NV12, 64, 64, src_rect=(0, 0, 100, 64)))
    ok, cx, cy = vp.GetAlignmentSize(V.VideoProcessor.media_type(V.CF.P010, 1918, 1080))
    assert ok and (cx * 2) % 256 == 0 and cx >= 1918 and cy == 1080


def test_no_cpu_fallback_and_error_convention():
    vp = V.VideoProcessor(device=-1)
    assert not vp.GetActive() and vp.GetDeviceIndex() == -1
    assert vp.InitMediaType(V.VideoProcessor.media_type(V.CF.NV12, 64, 64))
    hr = vp.ProcessSample(data=np.zeros(64 * 96, np.uint8), check=False)
    assert hr == V.HRESULT.E_UNEXPECTED and "no CPU path" in vp.last_error()
    assert vp._L.mpcvr_render(vp._h, 1, 0) == V.HRESULT.E_UNEXPECTED
    assert vp.SetRotation(90) == 0 and vp.GetVideoSize() == (64, 64) and vp.SetRotation(45) == V.HRESULT.E_INVALIDARG
    assert vp.SetRotation(0) == 0 and vp.SetFlip(True) == 0 and vp.SetFlip(False) == 0
    assert vp._L.mpcvr_process_sample(None, None) == V.HRESULT.E_POINTER
    with pytest.raises(V.MpcvrError):
        V.VideoProcessor(device=12345)
    s = vp.GetSettings(); s.iUpscaling = V.UPSCALE.Lanczos3; s.iSDRDisplayNits = 200; vp.Configure(s)
    assert vp.GetSettings().iUpscaling == V.UPSCALE.Lanczos3
    assert "Lanczos3" in vp.GetVPInfo()


def test_unorm_read_sequence_is_exact():
    """t = c*lo; q = fma(c, hi, t) with hi + lo = 1/N split into two floats equals the IEEE quotient c/N for every code of the
    three UNORM widths (what mpcvr_math.cuh::div_by_const_exact relies on; re-checked on the device in the GPU tests).
    The residual form used until round 2 (q = c*r; q += fma(-q, N, c)*r) is checked to give the same values."""
    from fractions import Fraction
    import math

    def rn32(fr):                                  # exact rational -> nearest float32, ties to even
        if fr == 0:
            return 0.0
        if fr < 0:
            return -rn32(-fr)
        e = math.floor(math.log2(float(fr)))
        while Fraction(2) ** e > fr: e -= 1
        while Fraction(2) ** (e + 1) <= fr: e += 1
        scale = Fraction(2) ** (e - 23)
        m = fr / scale
        fl = m.numerator // m.denominator
        rem = m - fl
        if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and fl % 2 == 1): fl += 1
        return float(Fraction(fl) * scale)

    for n in (255, 1023, 65535):
        hi = float(np.float32(1.0 / n)); lo = float(np.float32(1.0 / n - hi))
        step = 1 if n < 65535 else 7               # every code of the 8/10-bit widths; every 7th + the ends of the 16-bit one (all 65536 on the device)
        codes = sorted(set(list(range(0, n + 1, step)) + [n - 1, n] + [64 * k for k in range(1024)] if n == 65535 else range(n + 1)))
        for c in codes:
            t = rn32(Fraction(c) * Fraction(lo))
            q = rn32(Fraction(c) * Fraction(hi) + Fraction(t))
            assert q == rn32(Fraction(c, n)), (n, c)
        # the residual form
        N = np.float32(n); r = np.float32(1.0) / N
        c = np.arange(n + 1, dtype=np.float32)
        q = c * r
        e = (c.astype(np.float64) - q.astype(np.float64) * np.float64(N)).astype(np.float32)
        assert np.all(e.astype(np.float64) == c.astype(np.float64) - q.astype(np.float64) * np.float64(N)), "residual must be exact"
        q2 = (q.astype(np.float64) + e.astype(np.float64) * np.float64(r)).astype(np.float32)
        assert np.array_equal(q2, c / N)


def test_fused_readback_final_pass_is_exact():
    """kernel_strip.cuh OPT_C_FUSEDQ: for every 10-bit code c and every dither threshold d = k/1024 of the 32x32 table,
    floor(fma(RN(c/1023), 255, d))  ==  floor(fma(c, RN(255/1023), d + 2^-20))  -- the UNORM10 read-back of m_TexsPostScale and
    ps_final_pass (Shaders/d3d11/ps_final_pass.hlsl) as one multiply-add on the integer code.  All 1024 x 1024 pairs; the fma is
    emulated exactly (products of a 10-bit and a 24-bit number plus a 20-bit addend fit float64), rounded once to float32."""
    c = np.arange(1024, dtype=np.float64); k = np.arange(1024, dtype=np.float64)
    p = (c / 1023.0).astype(np.float32).astype(np.float64)
    assert np.array_equal(p.astype(np.float32), (np.arange(1024, dtype=np.float32) / np.float32(1023)))       # the correctly rounded quotient
    d = k / 1024.0
    spec = np.floor((p[:, None] * 255.0 + d[None, :]).astype(np.float32)).astype(np.int32)
    K1 = np.float64(np.float32(255.0 / 1023.0)); delta = 2.0 ** -20
    assert np.all((d + delta).astype(np.float32).astype(np.float64) == d + delta)                              # the shifted thresholds are floats
    fused = np.floor((c[:, None] * K1 + (d[None, :] + delta)).astype(np.float32)).astype(np.int32)
    assert np.array_equal(fused, spec)
    # the window is narrow: without the offset, or with four times the offset, some pair differs
    assert np.any(np.floor((c[:, None] * K1 + d[None, :]).astype(np.float32)).astype(np.int32) != spec)
    assert np.any(np.floor((c[:, None] * K1 + (d[None, :] + 4 * delta)).astype(np.float32)).astype(np.int32) != spec)
    # and the golden (code, threshold) -> u8 table of the final pass (minted by the oracle, tests/golden/make_golden.py) holds the same bytes
    import hashlib, json, os
    J = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "dither_stage.json")))
    assert hashlib.sha256(fused.astype(np.uint8).tobytes()).hexdigest() == J["table_sha256"]


def test_hdr_output_picks_the_convert_type():
    """UpdateConvertColorShader's convertType (Source/DX11VideoProcessor.cpp:2948-2950) with and without an HDR display"""
    vp = V.VideoProcessor(device=-1)                       # defaults: bHdrPassthrough = 1, bConvertToSdr = 1
    pq = V.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=15); hlg = V.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=16)
    assert vp.InitMediaType(V.VideoProcessor.media_type(V.CF.P010, 64, 64, pq)) and vp.debug_plan().transfer == 1      # PQ -> SDR
    assert vp.InitMediaType(V.VideoProcessor.media_type(V.CF.P010, 64, 64, hlg)) and vp.debug_plan().transfer == 2     # HLG -> SDR
    vp.SetDisplayHdr(True)
    assert vp.debug_plan().transfer == 4                                                                               # HLG -> PQ
    assert vp.InitMediaType(V.VideoProcessor.media_type(V.CF.P010, 64, 64, pq)) and vp.debug_plan().transfer == 0      # PQ untouched
    s = vp.GetSettings(); s.bHdrPassthrough = 0; vp.Configure(s)
    assert vp.debug_plan().transfer == 1                   # HDR display, but neither passthrough nor local tone mapping
    s.bHdrLocalToneMapping = 1; vp.Configure(s)
    assert vp.debug_plan().transfer == 0
    vp.SetDisplayHdr(False)
    assert vp.debug_plan().transfer == 1


def test_correction_shader_choice_follows_the_reference_rules():
    """m_pPSCorrection: D3D11 rule (Source/DX11VideoProcessor.cpp:1893-1931) and its D3D9 twin (Source/DX9VideoProcessor.cpp:1225-1238)"""
    pq = V.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=15); hlg = V.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=16)
    sdr2020 = V.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=5); sdr709 = V.exfmt(chroma=5, range_=2, matrix=1, prim=2, trc=5)

    def mode(ex, upstream=(True, False), d3d11=1, to_sdr=1, hdr_display=False, passthrough=1):
        vp = V.VideoProcessor(device=-1)
        s = vp.GetSettings(); s.bUseD3D11 = d3d11; s.bConvertToSdr = to_sdr; s.bHdrPassthrough = passthrough; vp.Configure(s)
        vp.SetDisplayHdr(hdr_display)
        if upstream:
            vp.SetUpstreamConverter(*upstream)
        assert vp.InitMediaType(V.VideoProcessor.media_type(V.CF.r210, 64, 64, ex))
        return vp.debug_plan().correction

    assert mode(pq, upstream=None) == 0 and mode(sdr709) == 0
    assert mode(pq, upstream=(True, True)) == 1 and mode(pq) == 2 and mode(hlg) == 3 and mode(sdr2020) == 4
    assert mode(hlg, to_sdr=0) == 4 and mode(pq, to_sdr=0) == 4          # !bTransFunc22, no HDR display, BT.2020 primaries
    assert mode(hlg, hdr_display=True) == 5 and mode(pq, hdr_display=True) == 0
    assert mode(pq, hdr_display=True, passthrough=0) == 2                  # HDR display but neither passthrough nor local tone mapping
    assert mode(pq, d3d11=0, upstream=(True, True)) == 2 and mode(hlg, d3d11=0) == 3 and mode(sdr2020, d3d11=0) == 4
    assert mode(pq, d3d11=0, to_sdr=0) == 4 and mode(sdr709, d3d11=0) == 0


def test_unorm8_readback_then_requantise_is_identity():
    """The 1:1 span kernel writes Q8(v) directly when an 8-bit convert surface is copied to an 8-bit back buffer
    (kernel_fused.cu fused1x_span, same8): UNORM8 read-back (c/255, correctly rounded) followed by the UNORM8 write
    (floor(fma(p, 255, 0.5))) must return c for every code."""
    import numpy as np
    c = np.arange(256, dtype=np.float32)
    p = (c / np.float32(255.0)).astype(np.float32)
    q = np.floor((p.astype(np.float64) * 255.0 + 0.5).astype(np.float32))
    assert np.array_equal(q, c)
