"""CPU checks of the oracle's RGB / gray / v210 rows: relations that hold by construction of the
reference's sample->texture repacks (Source/Helper.cpp:446-803) and of the no-convert-pass rule
(Source/DX11VideoProcessor.cpp:846-850, :3320-3322)."""
import numpy as np

import harness as H
import orclib as O


def _run(name, w, h, dw, dh, codes_from=None, bottom_up=False, seed=3, kind="fullnoise", **kw):
    cf = O.CF[name]
    case = H.Case(cf, w, h, dw, dh, kw.pop("ex", 0), name=name, **kw)
    codes = H.make_codes(O.CF[codes_from or name], w, h, kind, seed)
    buf, pitch, planes, pitches = H.pack(cf, w, h, codes, bottom_up=bottom_up)
    return H.run_oracle(case, planes, pitches), codes


def test_rgb32_identity_without_convert_pass():
    out, codes = _run("XRGB32", 40, 24, 40, 24, up=O.UPSCALE["Nearest"])
    px = out.reshape(24, 40, 4)
    assert np.array_equal(px[..., 2], codes["Y"]) and np.array_equal(px[..., 1], codes["U"]) and np.array_equal(px[..., 0], codes["V"])


def test_byte_orders_of_one_picture_agree():
    a, _ = _run("RGB24", 38, 22, 57, 33, up=O.UPSCALE["Lanczos3"])
    for other, bu in (("XRGB32", False), ("ARGB32", False), ("RGB24", True), ("XRGB32", True)):
        b, _ = _run(other, 38, 22, 57, 33, codes_from="RGB24", bottom_up=bu, up=O.UPSCALE["Lanczos3"])
        assert np.array_equal(a, b), (other, bu)
    a, _ = _run("RGB48", 38, 22, 57, 33, up=O.UPSCALE["Lanczos3"])
    for other in ("BGR48", "BGRA64", "B64A"):
        b, _ = _run(other, 38, 22, 57, 33, codes_from="RGB48", up=O.UPSCALE["Lanczos3"])
        assert np.array_equal(a, b), other


def test_planar_rgb_matches_packed_rgb_through_the_convert_pass():
    """GBRP always runs the convert pass (swizzled matrix); packed RGB does once ProcAmp is not neutral"""
    pa = dict(brightness=-8.0, contrast=1.1)
    a, codes = _run("RGB48", 32, 20, 48, 30, procamp=pa, up=O.UPSCALE["CatmullRom"])
    g = dict(Y=codes["U"], U=codes["V"], V=codes["Y"], bits=16)            # planes G, B, R  <-  R, G, B codes
    cf = O.CF["GBRP16"]
    buf, pitch, planes, pitches = H.pack(cf, 32, 20, g)
    b = H.run_oracle(H.Case(cf, 32, 20, 48, 30, 0, procamp=pa, up=O.UPSCALE["CatmullRom"]), planes, pitches)
    assert np.array_equal(a, b)


def test_v210_equals_y210():
    a, _ = _run("V210", 66, 20, 99, 30, codes_from="Y210", ex=H.EX_709, chroma=O.CHROMA_CatmullRom, up=O.UPSCALE["Lanczos2"])
    b, _ = _run("Y210", 66, 20, 99, 30, ex=H.EX_709, chroma=O.CHROMA_CatmullRom, up=O.UPSCALE["Lanczos2"])
    assert np.array_equal(a, b)


def test_gray_is_neutral_rgb():
    out, codes = _run("Y8", 40, 24, 40, 24, up=O.UPSCALE["Nearest"])
    px = out.reshape(24, 40, 4)
    assert np.array_equal(px[..., 0], px[..., 1]) and np.array_equal(px[..., 1], px[..., 2])
    assert np.abs(px[..., 0].astype(int) - codes["Y"]).max() <= 1            # full-range identity up to the 10-bit internal format


def test_spec_and_literal_modes_agree_on_rgb_rows():
    for name in ("RGB24", "r210", "GBRP10", "Y16"):
        cf = O.CF[name]
        case = H.Case(cf, 48, 28, 72, 42, 0, up=O.UPSCALE["Lanczos3"], procamp=dict(brightness=5.3, contrast=0.93))
        codes = H.make_codes(cf, 48, 28, "natural", 4)
        buf, pitch, planes, pitches = H.pack(cf, 48, 28, codes)
        a = H.run_oracle(case, planes, pitches, mode=O.MODE_SPEC); b = H.run_oracle(case, planes, pitches, mode=O.MODE_F64)
        n, d = H.compare(a, b, case.backbuffer)
        assert d <= 1 and n <= 0.01 * a.size, (name, n, d)
