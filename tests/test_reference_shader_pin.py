"""Pins against reference code EXECUTED here: the reference's own format table, extended-format defaults, set_colorspace and
convert-shader generator (Source/Helper.cpp:283-369, :949-1004, :1169-1211; Source/Shaders.cpp:66-930) are compiled where they
lie into oracle/_ref/libref_shaders.so (oracle/Makefile; CompileShader is stubbed to keep the generated HLSL text).

What is compared:
  * s_FmtConvMapping row by row with the oracle's table (oracle/orc_fmt.c) and the product's (csrc/fmt_table.cpp);
  * SpecifyExtendedFormat over random DXVA2_ExtendedFormat words, oracle and product;
  * set_colorspace over every (range, matrix) field value with the oracle's;
  * for every format x chroma scaler x chroma siting x blend x transfer x convert type: the fetch code and the op sequence after
    "//convert color" that GetShaderConvertColor EMITS, against the branch FN(fetch_yuv) / FN(convert_color) of the oracle TAKE
    (orc_describe_convert), and the matrix_conv_prim literals against the oracle's gamut matrix bit for bit.
The product's kernels are compared with the oracle pixel by pixel elsewhere (tests/test_gpu_parity.py); this file ties the
oracle's structure to the reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import orclib as O

REF = os.path.join(O.ORC_DIR, "_ref", "libref_shaders.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_shaders.so not built (reference tree absent at build time)")

_L = None


def ref():
    global _L
    if _L is None:
        L = C.CDLL(REF)
        L.ref_fmt_params.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.ref_fmt_count.restype = C.c_int
        L.ref_specify_extended_format.restype = C.c_uint32
        L.ref_specify_extended_format.argtypes = [C.c_uint32, C.c_int, C.c_uint, C.c_uint]
        L.ref_set_colorspace.argtypes = [C.c_uint32, C.POINTER(C.c_int)]
        L.ref_convert_shader.argtypes = [C.c_int, C.c_uint, C.c_long, C.c_long, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int,
                                         C.c_int, C.c_char_p, C.c_int]
        _L = L
    return _L


def orc():
    L = O.lib()
    L.orc_describe_convert.restype = C.c_int
    L.orc_describe_convert.argtypes = [C.POINTER(O.OrcConfig), C.c_void_p, C.c_char_p, C.c_int]
    return L


DXGI = {56: ("R16", 2), 61: ("R8", 1), 35: ("R16G16", 2), 49: ("R8G8", 1), 28: ("R8G8B8A8", 1), 24: ("R10G10B10A2", 4), 11: ("R16G16B16A16", 2),
        87: ("B8G8R8A8", 1), 88: ("B8G8R8X8", 1)}


def test_format_table_rows_equal_the_references():
    import videorenderer_b200 as V
    PL = V.lib()
    n = ref().ref_fmt_count()
    assert n == 40
    for cf in range(1, n):
        r = (C.c_int * 13)(); assert ref().ref_fmt_params(cf, r) == 0
        cformat, packsize, pitchcoeff, cstype, sub, cdepth, planes, f1, f2, f3, dw, dh, dx11 = list(r)
        f = O.lib().orc_get_fmt(cf).contents
        name = f.name.decode()
        assert cformat == cf == f.cformat
        assert (f.cstype, f.subsampling, f.cdepth, f.pitchcoeff) == (cstype, sub, cdepth, pitchcoeff), name
        if name != "v210":                       # the reference stores 8/3 = 2 for v210 and computes its pitch separately (Helper.cpp:364 remark 3)
            assert f.packsize == packsize, name
        # plane structure: 2 planes = bi-planar, 3 = planar; chroma divisors from the DX11 plane config
        if planes == 2:
            assert f.layout == 1 and (f.div_w, f.div_h) == (dw, dh), name
            assert DXGI[f1][0] in ("R8", "R16") and DXGI[f2][0] == {"R8": "R8G8", "R16": "R16G16"}[DXGI[f1][0]]
            assert f.bytes == DXGI[f1][1]
        elif planes == 3:
            assert f.layout == 2 and (f.div_w, f.div_h) == (dw, dh) and f.bytes == DXGI[f1][1], name
            assert f1 == f2 == f3
        else:
            assert f.layout in (3, 4, 5, 6), name
            if f.layout == 6:
                assert DXGI[dx11][0] in ("R8", "R16") and f.bytes == DXGI[dx11][1]
        # the product's table
        p = (C.c_int32 * 12)(); assert PL.mpcvr_debug_fmt_params(cf, p) == 0
        p = list(p)
        assert p[0] == cf and (p[3], p[4], p[5], p[2]) == (cstype, sub, cdepth, pitchcoeff), name
        if name != "v210":
            assert p[1] == packsize, name
        if planes >= 2:
            assert (p[8], p[9]) == (dw, dh) and p[6] == (1 if planes == 2 else 2) and p[7] == f.bytes, name
        assert (p[10], p[11]) == (f.lsb_shift, f.vu_swapped), name


def test_specify_extended_format_equals_the_references():
    import videorenderer_b200 as V
    PL = V.lib()
    rng = np.random.Generator(np.random.PCG64(5))
    words = [0, 0xFFFFFFFF] + [int(x) for x in rng.integers(0, 2 ** 32, 3000, dtype=np.uint64)]
    # plus words with single fields unknown (the defaults are what is being tested)
    for trc in (0, 5, 15, 16):
        for prim in (0, 2, 9):
            for mat in (0, 1, 2, 4, 7):
                for rg in (0, 1, 2):
                    for ch in (0, 1, 5, 7):
                        words.append(O.exfmt(chroma=ch, range_=rg, matrix=mat, prim=prim, trc=trc, sample=3))
    sizes = [(720, 576), (1024, 576), (1025, 576), (1024, 577), (1920, 1080), (3840, 2160), (2, 2)]
    for cf in range(1, 40):
        f = O.lib().orc_get_fmt(cf).contents
        for i, w in enumerate(words):
            sw, sh = sizes[i % len(sizes)]
            want = ref().ref_specify_extended_format(w, cf, sw, sh)
            assert O.lib().orc_specify_extended_format(w, f.cstype, f.subsampling, sw, sh) == want, (cf, hex(w))
            assert PL.mpcvr_debug_specify_extended_format(w, cf, sw, sh) == want, (cf, hex(w))


def test_set_colorspace_equals_the_references():
    for rg in range(8):
        for mat in range(8):
            for prim in (0, 2, 9):
                for trc in (0, 5, 15):
                    w = O.exfmt(chroma=5, range_=rg, matrix=mat, prim=prim, trc=trc, sample=3)
                    r = (C.c_int * 4)(); ref().ref_set_colorspace(w, r)
                    sp, lv = C.c_int(), C.c_int(); O.lib().orc_set_colorspace(w, C.byref(sp), C.byref(lv))
                    assert (sp.value, lv.value) == (r[0], r[1]), hex(w)
    r = (C.c_int * 4)(); ref().ref_set_colorspace(0, r)
    sp, lv = C.c_int(), C.c_int(); O.lib().orc_set_colorspace(0, C.byref(sp), C.byref(lv))
    assert (sp.value, lv.value) == (r[0], r[1])


# ---- the HLSL the reference emits -> the same vocabulary orc_describe_convert prints -------------------------------------
BODY = [
    (r"color\.rgb = float3\(mul\(cm_r, color\.rgb\), mul\(cm_g, color\.rgb\), mul\(cm_b, color\.rgb\)\) \+ cm_c;", "matrix"),
    (r"color = saturate\(color\);", "sat"),
    (r"color\.rgb = HLGtoLinear\(color\.rgb\);", "hlg2lin"),
    (r"color = LinearToST2084\(color, 1000\.0\);", "lin2pq(1000)"),
    (r"color = LinearToST2084\(color, 1\.0\);", "lin2pq(1)"),
    (r"color = ST2084ToLinear\(color, LuminanceScale\);", "pq2lin(LuminanceScale)"),
    (r"color = ST2084ToLinear\(color, 1\.0\);", "pq2lin(1)"),
    (r"color = max\(color, 0\.0\);", "max0"),
    (r"color\.rgb = mul\(mat, color\.rgb\);", "mat"),
    (r"if \(L2Enabled\) color = DolbyVisionTrims\(color\);", "trims?"),
    (r"color\.rgb = ToneMappingHable\(color\.rgb\);", "hable"),
    (r"color\.rgb = mul\(matrix_conv_prim, color\.rgb\);", "gamut"),
    (r"color = pow\(color, 1\.0/2\.2\);", "pow(1/2.2)"),
    (r"color = pow\(color, ([0-9.]+)\);", "pow(%s)"),
    (r"return color;", None), (r"\}", None), (r"float3x3 mat = \{", None), (r"[-0-9.e]+, [-0-9.e]+, [-0-9.e]+,?", None), (r"\};", None),
]


def parse_body(text):
    tail = text.split("//convert color\n", 1)[1]
    ops = []
    for line in tail.split("\n"):
        line = line.strip()
        if not line:
            continue
        for pat, tok in BODY:
            m = re.fullmatch(pat, line)
            if m:
                if tok:
                    ops.append(tok % m.groups() if "%s" in tok else tok)
                break
        else:
            raise AssertionError("unrecognised line in the reference's convert body: %r" % line)
    return ";".join(ops)


def parse_fetch(text, sub):
    head = text.split("//convert color\n", 1)[0]
    main = head.split("float4 main(PS_INPUT input) : SV_Target\n{\n", 1)[1]
    nplanes = 1 + ("Texture2D texUV" in head) + 2 * ("Texture2D texU " in head)
    if nplanes == 1:
        if ".zyxw" in main: return "packed444:zyxw"
        if ".yxzw" in main: return "packed444:yxzw"
        if "fmod(input.Tex.x*w, 2)" in main:
            uyvy = "color = float4(color.yxz, 1);" in main
            return "packed422:uyvy=%d:%s" % (uyvy, "catmull" if "CATMULLROM_05" in main else "linear")
        assert main.strip() == "float4 color = tex.Sample(samp, input.Tex);"
        return "single"                                       # RGB and gray sources share this text; the view format tells them apart
    out = "planes=%d" % nplanes
    if nplanes == 3 and re.search(r"Texture2D texV : register\(t1\)", head):
        out += ":vu"
    if "colorY = (colorY * 2 + y1 + y2) / 4;" in main:
        out += ":blend420"
    if "CATMULLROM_05(c0,c1,c2,c3);" in main:
        return out + ":catmull422"
    m = re.search(r"float2 t = frac\(input\.Tex \* \(wh\*0\.5\)\)\+float2\(([-0-9.]+),([-0-9.]+)\);", main)
    if m:
        return out + ":catmull420:off2=(%s,%s)" % m.groups()
    m = re.search(r"float2 pos = input\.Tex(\+float2\(([^,]+),([^)]+)\))?;", main)
    if m:
        # luma-texel offset (ox, oy) of the sample point -> quarter chroma texels: q = (W/cw) * ((s + .5 + o) / div) * 4 - 2 = a*s + b
        val = {"dx*0.5": 0.5, "dy*0.5": 0.5, "0": 0.0, None: 0.0}
        oxl, oyl = val[m.group(2)], val[m.group(3)]
        divh = 2 if sub == 420 else 1
        ax, bx = 2, int(round(2 * (0.5 + oxl) - 2))
        ay, by = 4 // divh, int(round((4 / divh) * (0.5 + oyl) - 2))
        return out + ":bilinear:q=(%dx%+d,%dy%+d)" % (ax, bx, ay, by)
    assert ("colorUV = texUV.Sample(samp, input.Tex).rg;" in main) or ("colorUV[0] = texU.Sample(samp, input.Tex).r;" in main)
    return out + ":nearest"


def parse_gamut(text):
    m = re.search(r"static const float3x3 matrix_conv_prim = \{\n(.*?)\};", text, re.S)
    if not m:
        return None
    vals = [np.float32(x) for x in re.findall(r"[-0-9.e]+", m.group(1))]
    assert len(vals) == 9
    return " ".join("%08x" % np.array(v, np.float32).view(np.uint32) for v in vals) + " "


def describe(cf, ex, chroma, to_sdr, hdr_out, blend, dovi=None, w=1920, h=1080):
    cfg = O.make_config(cf, w, h, w, h, ex, chroma=chroma, to_sdr=to_sdr, deint_blend=blend, hdr={"out": 1} if hdr_out else None)
    buf = C.create_string_buffer(2048)
    rc = orc().orc_describe_convert(C.byref(cfg), C.byref(dovi) if dovi is not None else None, buf, len(buf))
    assert rc == 0, rc
    return dict(line.split("=", 1) for line in buf.value.decode().strip().split("\n"))


def reference_shader(cf, ex_specified, chroma, convert_type, blend, dovi_mmr=-1, lms=None, w=1920, h=1080):
    buf = C.create_string_buffer(1 << 16)
    arr = (C.c_double * 9)(*lms) if lms is not None else None
    # packed 4:2:2 textures are half as wide (Source/Helper.cpp:305, DX11Helper.h:127-130); only "#define w" depends on it
    n = ref().ref_convert_shader(1, w, w, h, cf, ex_specified, dovi_mmr, arr, chroma, convert_type, blend, buf, len(buf))
    assert 0 < n < len(buf)
    return buf.value.decode()


SITINGS = (1, 5, 7, 0, 6)
TRCS = (0, 1, 2, 3, 4, 5, 6, 7, 8, 14, 15, 16)


def test_fetch_branch_equals_the_shader_the_reference_generates():
    """every format x chroma scaler x chroma siting x blend-deinterlace"""
    seen = set()
    for cf in range(1, 40):
        f = O.lib().orc_get_fmt(cf).contents
        for chroma in (O.CHROMA_Nearest, O.CHROMA_Bilinear, O.CHROMA_CatmullRom):
            for siting in SITINGS:
                for blend in (0, 1):
                    ex = O.exfmt(chroma=siting, range_=2, matrix=1, prim=2, trc=5, sample=3)
                    d = describe(cf, ex, chroma, 1, 0, blend)
                    exs = int(d["exfmt"], 16)
                    text = reference_shader(cf, exs, chroma, 1, blend)
                    got = parse_fetch(text, f.subsampling)
                    want = d["fetch"]
                    if want in ("rgb", "gray"):
                        want = "single"
                    assert got == want, (f.name.decode(), chroma, siting, blend, got, want)
                    seen.add(got)
    assert len(seen) >= 20, sorted(seen)


def test_convert_body_equals_the_shader_the_reference_generates():
    """every transfer function x primaries x convert type (to SDR / none / to PQ) on YUV and RGB sources: the op sequence after
    "//convert color" and the matrix_conv_prim literals"""
    bodies = set()
    for cf in (O.CF["P010"], O.CF["NV12"], O.CF["YUY2"], O.CF["Y416"], O.CF["RGB48"], O.CF["GBRP10"], O.CF["Y16"]):
        for trc in TRCS:
            for prim in (0, 2, 9, 11):
                for to_sdr, hdr_out in ((1, 0), (0, 0), (1, 1), (0, 1)):
                    if trc == 1 and prim == 9:
                        # DXVA2_VideoTransFunc_10 with BT.2020 primaries: the reference emits a literal "\\\\nothing" line (Shaders.cpp:895),
                        # text fxc rejects, so that convert shader never exists; not a case
                        continue
                    ex = O.exfmt(chroma=5, range_=2, matrix=4 if prim == 9 else 1, prim=prim, trc=trc, sample=3)
                    d = describe(cf, ex, O.CHROMA_Bilinear, to_sdr, hdr_out, 0)
                    exs = int(d["exfmt"], 16)
                    # convertType as CDX11VideoProcessor picks it (:2948-2950): to PQ on an HDR display, else to SDR when enabled
                    convert_type = 2 if hdr_out else (1 if to_sdr else 0)
                    text = reference_shader(cf, exs, O.CHROMA_Bilinear, convert_type, 0)
                    assert parse_body(text) == d["body"], (cf, trc, prim, to_sdr, hdr_out, parse_body(text), d["body"])
                    g = parse_gamut(text)
                    if "gamut" in d["body"]:
                        assert g == d["gamut"], (g, d["gamut"])
                    bodies.add(d["body"])
    assert len(bodies) >= 6, sorted(bodies)


def test_dolby_vision_body_equals_the_shader_the_reference_generates():
    import harness as H
    for mmr in (0, 1):
        dv = H.make_dovi(seed=3, mmr=mmr)
        for to_sdr, hdr_out in ((1, 0), (0, 1)):
            ex = O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=15, sample=3)
            d = describe(O.CF["P010"], ex, O.CHROMA_Bilinear, to_sdr, hdr_out, 0, dovi=dv)
            text = reference_shader(O.CF["P010"], int(d["exfmt"], 16), O.CHROMA_Bilinear, 2 if hdr_out else 1, 0, dovi_mmr=mmr,
                                    lms=list(dv.rgb_to_lms_matrix))
            assert parse_body(text) == d["body"], (mmr, to_sdr, hdr_out, parse_body(text), d["body"])
            assert ("reshape_mmr" in text) == bool(mmr)


def test_convert_color_constants_equal_the_references():
    """SetShaderConvertColorParams (Source/DX11VideoProcessor.cpp:813-873) compiled where it lies (ref_convert_color_params): the
    PS_COLOR_TRANSFORM rows -- colour matrix with ProcAmp, the GBRP column rotation, the gray rows, the Dolby Vision matrix with its
    offsets -- and m_PSConvColorData.bEnable, against the oracle's build_cparams, bit for bit, for every format x every
    (range, matrix) tag x a grid of ProcAmp values that are exact in DXVA2's 16.16 fixed point."""
    import harness as H
    R = ref(); L = O.lib()
    R.ref_convert_color_params.restype = C.c_int
    R.ref_convert_color_params.argtypes = [C.c_int, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                           (C.c_float * 3) * 3, C.c_float * 3, C.POINTER(C.c_int32)]
    L.orc_test_convert_params.restype = C.c_int
    L.orc_test_convert_params.argtypes = [C.POINTER(O.OrcConfig), C.c_void_p, (C.c_float * 3) * 3, C.c_float * 3, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]
    rng = np.random.Generator(np.random.PCG64(99))
    n = ref().ref_fmt_count()
    checked = enabled = 0
    for cf in range(1, n):
        for range_ in (0, 1, 2):
            for matrix in (0, 1, 2, 3, 4, 7):
                for trial in range(3):
                    if trial == 0: b, c, h, s = 0.0, 1.0, 0.0, 1.0
                    else:
                        b = float(rng.integers(-100, 101)); c = float(rng.integers(0, 513)) / 256.0
                        h = float(rng.integers(-180, 181)); s = float(rng.integers(0, 513)) / 256.0
                    ex = O.exfmt(chroma=5, range_=range_, matrix=matrix, prim=2, trc=5)
                    cfg = O.make_config(cf, 64, 32, 64, 32, ex, brightness=b, contrast=c, hue=h, saturation=s)
                    ocm, occ, oen, oex = ((C.c_float * 3) * 3)(), (C.c_float * 3)(), C.c_int32(), C.c_uint32()
                    assert L.orc_test_convert_params(C.byref(cfg), None, ocm, occ, C.byref(oen), C.byref(oex)) == 0, cf
                    rcm, rcc, ren = ((C.c_float * 3) * 3)(), (C.c_float * 3)(), C.c_int32()
                    assert R.ref_convert_color_params(cf, oex.value, b, c, h, s, None, None, rcm, rcc, C.byref(ren)) == 0, (cf, b, c, h, s)
                    assert bytes(rcm) == bytes(ocm) and bytes(rcc) == bytes(occ), (cf, range_, matrix, b, c, h, s, np.array(rcm), np.array(ocm), np.array(rcc), np.array(occ))
                    assert ren.value == oen.value, (cf, b, c)
                    checked += 1; enabled += ren.value
                    # ... and the PRODUCT's SetShaderConvertColorParams (csrc/video_processor.cpp) on a planning-only instance (no GPU needed)
                    vp = H.Case(cf, 64, 32, 64, 32, ex, procamp=dict(brightness=b, contrast=c, hue=h, saturation=s)).make_processor(device=-1)
                    pm, pc, _, _ = vp.debug_color_params()
                    assert np.asarray(pm, np.float32).tobytes() == bytes(rcm) and np.asarray(pc, np.float32).tobytes() == bytes(rcc), (cf, range_, matrix, b, c, h, s)
    assert checked > 2000 and 0 < enabled < checked        # RGB sources with neutral ProcAmp run without a convert pass
    # Dolby Vision: the RPU's own matrix and offsets, with ProcAmp brightness / contrast (:817-834)
    for seed in range(20):
        dv = H.make_dovi(seed=seed, mmr=seed & 1)
        b = float(rng.integers(-100, 101)); c = float(rng.integers(0, 513)) / 256.0
        ex = O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=15)
        cfg = O.make_config(O.CF["P010"], 64, 32, 64, 32, ex, brightness=b, contrast=c)
        ocm, occ, oen, oex = ((C.c_float * 3) * 3)(), (C.c_float * 3)(), C.c_int32(), C.c_uint32()
        assert L.orc_test_convert_params(C.byref(cfg), C.byref(dv), ocm, occ, C.byref(oen), C.byref(oex)) == 0
        rcm, rcc, ren = ((C.c_float * 3) * 3)(), (C.c_float * 3)(), C.c_int32()
        m = (C.c_double * 9)(*dv.ycc_to_rgb_matrix); off = (C.c_double * 3)(*dv.ycc_to_rgb_offset)
        assert R.ref_convert_color_params(O.CF["P010"], oex.value, b, c, 0.0, 1.0, m, off, rcm, rcc, C.byref(ren)) == 0
        assert bytes(rcm) == bytes(ocm) and bytes(rcc) == bytes(occ) and ren.value == oen.value == 1, (seed, np.array(rcm), np.array(ocm), np.array(rcc), np.array(occ))
