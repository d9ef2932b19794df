"""Pins of the CPU oracle (CPU only): the reference ships no tests for this path, so the golden
vectors are minted in tests/golden/make_golden.py from (a) the reference's own csputils.cpp compiled in place and
(b) closed-form facts about the shaders (weights, dither stage).  These tests keep the oracle on those pins."""
import hashlib
import json
import os

import numpy as np
import pytest

import harness as H
import orclib as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return [int(v) for v in np.asarray(a, np.float32).ravel().view(np.uint32)]


def _exfmt_for(space, levels):
    matrix = {1: 2, 2: 1, 3: 3, 4: 4, 8: 7}[space]          # mp_csp -> DXVA2 matrix value (Source/Helper.cpp:966-974)
    return O.exfmt(chroma=5, range_=1 if levels == 2 else 2, matrix=matrix, prim=2, trc=5)


def test_csp_matrices_match_reference_golden():
    kat = json.load(open(os.path.join(G, "kat_matrices.json")))
    for name, e in kat["matrices"].items():
        m, c = O.csp_matrix(_exfmt_for(e["space"], e["levels"]), e["bits"])
        assert _bits(m) == e["m_bits"], name
        assert _bits(c) == e["c_bits"], name
    for name, e in kat["procamp_bt709_tv_8"].items():
        m, c = O.csp_matrix(_exfmt_for(2, 1), 8, brightness=e["brightness"], contrast=e["contrast"], hue=e["hue"], saturation=e["saturation"])
        assert _bits(m) == e["m_bits"] and _bits(c) == e["c_bits"], name
    g = np.zeros((3, 3), np.float32); O.lib().orc_gamut_matrix_2020_to_709(g.ctypes.data)
    assert _bits(g) == kat["gamut_2020_to_709_bits"]


def test_survey_kats_a3_a5():
    """the decimal KATs quoted in SURVEY 8(a3)/(a5)"""
    m, c = O.csp_matrix(H.EX_2020_PQ, 16)
    np.testing.assert_allclose(m, [[1.16893196, 0, 1.68523157], [1.16893196, -0.188057855, -0.652965128], [1.16893196, 2.15013862, 0]], rtol=0, atol=1e-7)
    np.testing.assert_allclose(c, [-0.915687978, 0.347458541, -1.14814508], rtol=0, atol=1e-7)
    m, c = O.csp_matrix(H.EX_709, 8)
    np.testing.assert_allclose(m, [[1.16438353, 0, 1.79274106], [1.16438353, -0.213248596, -0.532909334], [1.16438353, 2.11240172, 0]], rtol=0, atol=1e-7)
    np.testing.assert_allclose(c, [-0.972945094, 0.301482677, -1.13340223], rtol=0, atol=1e-7)
    g = np.zeros((3, 3), np.float32); O.lib().orc_gamut_matrix_2020_to_709(g.ctypes.data)
    np.testing.assert_allclose(g, [[1.66049695, -0.587656736, -0.0728399456], [-0.124547064, 1.13289523, -0.0083479844],
                                   [-0.0181536824, -0.100597292, 1.11875105]], rtol=0, atol=1e-7)


@pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built (reference tree absent)")
def test_csp_matches_reference_library_sweep():
    """restatement vs the reference's own compiled csputils over every matrix/range/depth and random procamp"""
    R = O.ref()
    rng = np.random.Generator(np.random.PCG64(0))
    for space in (1, 2, 3, 4, 8):
        for levels in (1, 2):
            for depth in (8, 10, 16):
                for _ in range(6):
                    b, ct, hue, sat = rng.uniform(-100, 100), rng.uniform(0, 2), rng.uniform(-180, 180), rng.uniform(0, 2)
                    if _ == 0:
                        b, ct, hue, sat = 0.0, 1.0, 0.0, 1.0
                    m, c = O.csp_matrix(_exfmt_for(space, levels), depth, brightness=b, contrast=ct, hue=hue, saturation=sat)
                    rm = np.zeros(9, np.float32); rc = np.zeros(3, np.float32)
                    hue_r = np.float32(np.float32(np.float32(hue) / np.float32(180)) * np.arccos(-1.0))
                    R.ref_csp_matrix(space, levels, float(np.float32(b) / np.float32(255)), float(np.float32(ct)), float(hue_r), float(np.float32(sat)), 0, depth,
                                     rm.ctypes.data, rc.ctypes.data)
                    assert _bits(m) == _bits(rm) and _bits(c) == _bits(rc), (space, levels, depth, b, ct, hue, sat)
    for pin, pout in ((4, 3), (3, 4), (1, 3), (2, 3), (5, 3), (10, 3)):
        g = np.zeros((3, 3), np.float32); O.lib().orc_gamut_matrix(pin, pout, g.ctypes.data)
        rg = np.zeros(9, np.float32); R.ref_gamut_matrix(pin, pout, rg.ctypes.data)
        assert _bits(g) == _bits(rg)


def test_resize_weights_golden_and_survey_kats():
    kat = json.load(open(os.path.join(G, "kat_weights.json")))
    for name, e in kat.items():
        base, w, wsum, n = O.axis_taps(e["kind"], e["filter"], e["variant"], 0, e["src"], e["dst"], e["i"])
        assert base == e["base"] and _bits(w) == e["w_bits"] and _bits([wsum])[0] == e["wsum_bits"], name
    # Lanczos3, t = 1/4 (odd outputs of a 2x upscale) and mirrored for t = 3/4
    _, w1, _, _ = O.axis_taps(1, O.UPSCALE["Lanczos3"], 0, 0, 3840, 7680, 1)
    np.testing.assert_allclose(w1, [-0.010831532, -0.084724804, 0.891053494, 0.239913152, -0.017905185, -0.017505125], atol=5e-8)
    _, w0, _, _ = O.axis_taps(1, O.UPSCALE["Lanczos3"], 0, 0, 3840, 7680, 0)
    np.testing.assert_allclose(w0, w1[::-1], atol=5e-8)
    # Catmull-Rom 1.5x phases 5/6, 1/2, 1/6
    _, w, _, _ = O.axis_taps(1, O.UPSCALE["CatmullRom"], 0, 0, 1920, 2880, 0)
    np.testing.assert_allclose(w, [-.01157407, .13194444, .9375, -.05787037], atol=2e-7)
    _, w, _, _ = O.axis_taps(1, O.UPSCALE["CatmullRom"], 0, 0, 1920, 2880, 1)
    np.testing.assert_array_equal(w, np.array([-.0625, .5625, .5625, -.0625], np.float32))
    # interpolation weights sum to 1 (the wc compensation), convolution weights are normalised by wsum
    for filt in (1, 2, 3, 4):
        for i in range(0, 50):
            _, w, _, _ = O.axis_taps(1, filt, 0, 0, 97, 233, i)
            assert abs(float(np.sum(w.astype(np.float64))) - 1.0) < 3e-7


def test_d3d11_quirks_are_variants():
    """B2: the D3D11 convolution is shifted by half a source texel relative to D3D9"""
    b11, w11, s11, _ = O.axis_taps(2, O.DOWNSCALE["Bilinear"], O.VARIANT_D3D11, 0, 400, 100, 10)
    b9, w9, s9, _ = O.axis_taps(2, O.DOWNSCALE["Bilinear"], O.VARIANT_D3D9, 0, 400, 100, 10)
    c11 = np.sum((b11 + np.arange(len(w11)) + 0.5) * w11) / s11
    c9 = np.sum((b9 + np.arange(len(w9)) + 0.5) * w9) / s9
    assert abs((c11 - c9) - 0.5) < 1e-3
    assert abs(c9 - (10 + 0.5) * 4) < 1e-3


def test_dither_table_and_final_stage_exhaustive():
    gold = json.load(open(os.path.join(G, "dither32x32.json")))
    K = np.ctypeslib.as_array(O.lib().orc_dither_thresholds(), shape=(1024,)).copy()
    assert K.tolist() == gold["thresholds_x1024"]
    assert sorted(K.tolist()) == list(range(1024))
    assert gold["sha256"] == "24b5048f1390879d3435072f6e48b62b042fd4787b33bfac5e6d9bc84ec28e7c"
    # every (code10, threshold) pair through the oracle's real final pass: 32x32 block per code, no resize
    codes = 1024
    surf = np.zeros((32 * codes, 32), np.uint32)
    for c in range(codes):
        surf[32 * c:32 * c + 32, :] = c | (c << 10) | (c << 20) | (3 << 30)
    case = H.Case(O.CF["P010"], 32, 32 * codes, 32, 32 * codes, H.EX_2020_PQ, name="dither_sweep")
    out = O.process(case.orc_config(), None, None, from_convert=surf.view(np.uint8).reshape(32 * codes, 128))
    got = H.rgb_view(out, O.SURF_BGRA8)[..., 0].reshape(codes, 32, 32)
    table = np.zeros((1024, 1024), np.uint8)
    kk = K.reshape(32, 32)
    for c in range(codes):
        table[c, kk] = got[c]
    ds = json.load(open(os.path.join(G, "dither_stage.json")))
    assert hashlib.sha256(table.tobytes()).hexdigest() == ds["table_sha256"]
    exact = ((np.arange(1024, dtype=np.int64)[:, None] * 255 * 1024 + np.arange(1024, dtype=np.int64)[None, :] * 1023) // (1023 * 1024)).astype(np.uint8)
    assert np.argwhere(table != exact).tolist() == [[345, 3], [686, 3], [690, 6], [694, 9]]
    # all three arithmetic modes agree on this integer stage
    for mode in (O.MODE_LITERAL, O.MODE_F64):
        o2 = O.process(case.orc_config(), None, None, from_convert=surf.view(np.uint8).reshape(32 * codes, 128), mode=mode)
        if mode == O.MODE_LITERAL:
            assert np.array_equal(o2, out)


def test_frame_hashes_golden():
    gold = json.load(open(os.path.join(G, "frames.json")))
    for case in H.baseline_cases(scale=0.125):
        for kind in ("ramp", "noise", "natural"):
            codes = H.make_codes(case.cf, case.w, case.h, kind, seed=1)
            buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
            e = gold["%s/%s" % (case.name, kind)]
            assert hashlib.sha256(buf.tobytes()).hexdigest() == e["in_sha256"]
            out = H.run_oracle(case, planes, pitches)
            rgb = np.ascontiguousarray(H.rgb_view(out, case.backbuffer))
            assert hashlib.sha256(rgb.tobytes()).hexdigest() == e["out_rgb_sha256"], (case.name, kind)


def test_math_contract_accuracy_vs_libm():
    """SPEC exp2/log2 are within ~2 ulp of the true functions (orc_math.h header claims LG2 <= 2.1, EX2 < 1)"""
    L = O.lib()
    rng = np.random.Generator(np.random.PCG64(1))
    xs = np.concatenate([np.exp(rng.uniform(-80, 80, 20000)), rng.uniform(0.5, 2.0, 20000), [1.0, 2.0, 0.5, 1.4142135]]).astype(np.float32)
    got = np.array([L.orc_test_lg2(float(v)) for v in xs], np.float32).astype(np.float64)
    ref = np.log2(xs.astype(np.float64))
    ulp = np.abs(np.spacing(ref.astype(np.float32)).astype(np.float64))
    assert np.max(np.abs(got - ref) / ulp) <= 2.2
    assert L.orc_test_lg2(1.0) == 0.0 and L.orc_test_lg2(2.0) == 1.0 and L.orc_test_lg2(0.5) == -1.0
    ts = np.concatenate([rng.uniform(-125.4, 127, 40000), [0.0, -1.0, 1.0, -125.0]]).astype(np.float32)
    got = np.array([L.orc_test_ex2(float(v)) for v in ts], np.float32).astype(np.float64)
    ref = np.exp2(ts.astype(np.float64))
    ulp = np.abs(np.spacing(ref.astype(np.float32)).astype(np.float64))
    assert np.max(np.abs(got - ref) / ulp) <= 1.0
    assert L.orc_test_ex2(0.0) == 1.0 and L.orc_test_ex2(-1.0) == 0.5 and L.orc_test_ex2(-127.0) == 0.0 and L.orc_test_ex2(-125.6) == 0.0
    assert L.orc_test_ex2(-125.0) == 2.0 ** -125 and L.orc_test_ex2(127.0) == 2.0 ** 127 and L.orc_test_ex2(128.0) == np.inf
    assert L.orc_test_pow(0.0, 0.5, 0) == 0.0 and L.orc_test_pow(1.0, 78.84375, 0) == 1.0
    # SURVEY a8 scalar KAT (gray through PQ EOTF at scale 80, Hable, gamma): checked on the transfer golden
    kat = json.load(open(os.path.join(G, "kat_transfer.json")))["pow(v,1/m2)"]
    for v, e in kat.items():
        got = L.orc_test_pow(float(np.float32(float(v))), float(np.float32(1.0) / np.float32(78.84375)), 0)
        assert _bits([got])[0] == e["spec_bits"]
        assert abs(got - e["f64"]) <= 4e-7 * max(abs(e["f64"]), 1e-30) + 1e-30


@pytest.mark.parametrize("idx", range(4))
def test_spec_arithmetic_stays_close_to_literal_and_f64(idx):
    """SPEC (fused mad, polynomial exp2/log2) vs LITERAL (written order, libm) vs F64: the contract the CUDA path
    matches is within a fraction of a percent of one-code differences of the literal restatement."""
    case = H.baseline_cases(scale=0.125)[idx]
    codes = H.make_codes(case.cf, case.w, case.h, "natural", seed=3)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    spec = H.run_oracle(case, planes, pitches)
    lit = H.run_oracle(case, planes, pitches, mode=O.MODE_LITERAL)
    f64 = H.run_oracle(case, planes, pitches, mode=O.MODE_F64)
    total = H.rgb_view(spec, case.backbuffer).size
    for other in (lit, f64):
        n, d = H.compare(spec, other, case.backbuffer)
        assert n <= 0.01 * total
        if case.backbuffer != O.SURF_RGBA16F:
            assert d <= 1


def test_pass_chain_decisions():
    L = O.lib()
    # UpdateTexParams (fact 4)
    assert L.orc_internal_fmt(O.TEXFMT_AUTOINT, 8) == O.SURF_BGRA8 and L.orc_internal_fmt(O.TEXFMT_AUTOINT, 16) == O.SURF_RGB10A2
    assert L.orc_internal_fmt(O.TEXFMT_16FLOAT, 8) == O.SURF_RGBA16F and L.orc_internal_fmt(O.TEXFMT_8INT, 16) == O.SURF_BGRA8
    # m_bFinalPass (fact 6)
    assert L.orc_need_final_pass(O.SURF_BGRA8, O.SURF_BGRA8, 1) == 0
    assert L.orc_need_final_pass(O.SURF_BGRA8, O.SURF_RGB10A2, 1) == 1 and L.orc_need_final_pass(O.SURF_BGRA8, O.SURF_RGB10A2, 0) == 0
    assert L.orc_need_final_pass(O.SURF_RGB10A2, O.SURF_RGB10A2, 1) == 0 and L.orc_need_final_pass(O.SURF_RGB10A2, O.SURF_RGBA16F, 1) == 1
    # scaler choice (fact 8)
    assert L.orc_scaler_kind(100, 100, 1) == 0 and L.orc_scaler_kind(100, 200, 1) == 1
    assert L.orc_scaler_kind(200, 100, 1) == 1 and L.orc_scaler_kind(200, 100, 0) == 2 and L.orc_scaler_kind(201, 100, 1) == 2
    # SpecifyExtendedFormat defaults
    ex = L.orc_specify_extended_format(0, 0, 420, 1920, 1080)
    assert ex == O.exfmt(chroma=5, range_=2, matrix=1, light=3, prim=2, trc=5)
    ex = L.orc_specify_extended_format(0, 0, 422, 720, 576)
    assert ex == O.exfmt(chroma=0, range_=2, matrix=2, light=3, prim=2, trc=5)
    assert L.orc_specify_extended_format(12345, 1, 444, 100, 100) == 0


def test_edge_cases_oracle():
    """tiny frames, 1-pixel crops, clamp addressing: the oracle runs and stays within range"""
    for (w, h, dw, dh) in ((2, 2, 7, 5), (4, 2, 2, 2), (16, 16, 16, 16)):
        case = H.Case(O.CF["NV12"], w, h, dw, dh, H.EX_709, up=O.UPSCALE["Lanczos3"], down=O.DOWNSCALE["Lanczos"], name="tiny")
        codes = H.make_codes(case.cf, w, h, "fullnoise", 1)
        buf, pitch, planes, pitches = H.pack(case.cf, w, h, codes)
        out = H.run_oracle(case, planes, pitches)
        assert out.shape == (dh, dw * 4)
    # constant gray stays constant through Lanczos3 + dither-free 8-bit path (unit DC gain after wc compensation)
    case = H.Case(O.CF["NV12"], 32, 32, 64, 64, H.EX_709, up=O.UPSCALE["Lanczos3"], name="dc")
    codes = H.make_codes(case.cf, 32, 32, "noise", 1)
    codes["Y"][:] = 126; codes["U"][:] = 128; codes["V"][:] = 128
    buf, pitch, planes, pitches = H.pack(case.cf, 32, 32, codes)
    out = H.rgb_view(H.run_oracle(case, planes, pitches), O.SURF_BGRA8)
    assert out.min() == out.max() == 128


@pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built (reference tree absent)")
def test_rgb_and_gray_matrices_match_reference_library():
    """extfmt.value == 0 -> MP_CSP_RGB / PC levels (Source/Helper.cpp:953-957); RGB and gray sources incl. ProcAmp"""
    R = O.ref()
    rng = np.random.Generator(np.random.PCG64(1))
    for gray in (0, 1):
        for depth in (8, 10, 16):
            for k in range(5):
                b, ct = (0.0, 1.0) if k == 0 else (rng.uniform(-100, 100), rng.uniform(0, 2))
                m, c = O.csp_matrix(0, depth, gray=gray, brightness=b, contrast=ct)
                rm = np.zeros(9, np.float32); rc = np.zeros(3, np.float32)
                R.ref_csp_matrix(6, 2, float(np.float32(b) / np.float32(255)), float(np.float32(ct)), 0.0, 1.0, gray, depth, rm.ctypes.data, rc.ctypes.data)
                assert _bits(m) == _bits(rm) and _bits(c) == _bits(rc), (gray, depth, b, ct)


@pytest.mark.skipif(O.ref_copy() is None, reason="oracle/_ref/libref_copy.so not built (reference tree absent)")
@pytest.mark.parametrize("name,which,bpp_src,bpp_dst,ncmp", [
    ("RGB24", 1, 3, 4, 3), ("RGB48", 2, 6, 8, 6), ("BGR48", 3, 6, 8, 6), ("BGRA64", 4, 8, 8, 8), ("B64A", 5, 8, 8, 8),
    ("r210", 7, 4, 4, 4)])
def test_repack_matches_reference_copy_functions(name, which, bpp_src, bpp_dst, ncmp):
    """orc_repack vs the reference's CopyFrame* compiled where they lie (Source/Helper.cpp:446-787), incl. a bottom-up DIB"""
    R = O.ref_copy()
    rng = np.random.Generator(np.random.PCG64(which))
    for w, h in ((16, 5), (37, 4), (64, 3)):
        src_pitch = (w * bpp_src + 3) // 4 * 4 if name in ("RGB24", "BGR48") else w * bpp_src
        src = rng.integers(0, 256, (h + 1) * src_pitch + 64, dtype=np.uint8)        # slack: the reference reads whole words
        for bottom_up in ((0, 1) if name == "RGB24" else (0,)):
            dst_pitch = w * bpp_dst
            mine = O.repack(O.CF[name], w, h, src, -src_pitch if bottom_up else src_pitch, dst_pitch)
            ref = np.zeros((h, dst_pitch + 64), np.uint8)
            start = src.ctypes.data + (src_pitch * (h - 1) if bottom_up else 0)     # MemCopyToTexSrcVideo :1246
            R.ref_copy_frame(which, h, ref.ctypes.data, dst_pitch + 64, start, -src_pitch if bottom_up else src_pitch)
            a = mine.reshape(h, w, bpp_dst)[..., :ncmp]; b = ref[:, :dst_pitch].reshape(h, w, bpp_dst)[..., :ncmp]
            if name == "RGB48":     # CopyFrameRGB48 (:541-565) has no remainder loop: the last w % 4 texels of a row stay unwritten
                a = a[:, :w & ~3]; b = b[:, :w & ~3]
            if name == "r210":                                                       # alpha bits are zero in both
                assert np.array_equal(mine.reshape(h, w, 4), ref[:, :dst_pitch].reshape(h, w, 4))
            assert np.array_equal(a, b), (name, w, h, bottom_up)


@pytest.mark.skipif(O.ref_copy() is None, reason="oracle/_ref/libref_copy.so not built (reference tree absent)")
def test_v210_repack_matches_reference_copy_function():
    R = O.ref_copy()
    rng = np.random.Generator(np.random.PCG64(9))
    for w in (48, 96, 1280 // 4):
        h = 3
        src_pitch = ((w + 5) // 6 * 16 + 127) // 128 * 128
        src = rng.integers(0, 256, h * src_pitch, dtype=np.uint8)
        dst_pitch = w * 4                                                            # Y210 texture row: 8 bytes per 2 pixels
        mine = O.repack(O.CF["V210"], w, h, src, src_pitch, dst_pitch)
        ref = np.zeros((h, dst_pitch), np.uint8)
        R.ref_copy_frame(6, h, ref.ctypes.data, dst_pitch, src.ctypes.data, src_pitch)
        n = (w // 6) * 6 * 4                                                         # whole 6-pixel groups
        assert np.array_equal(mine[:, :n], ref[:, :n]), w
