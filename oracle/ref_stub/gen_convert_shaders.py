#!/usr/bin/env python3
"""oracle/ref_stub/gen_convert_shaders.py <REF> <libref_shaders.so> -- TEST INFRASTRUCTURE ONLY.

Asks the reference's own GetShaderConvertColor (Source/Shaders.cpp:593-930, compiled where it lies into libref_shaders.so; its
CompileShader is a stub that keeps the text) for the HLSL convert shader of each case below, and prints one C++ unit in which every
generated shader sits in its own namespace and can be called per pixel.  The text is the reference's; the edits are syntax only:
  * the three "//#include x.hlsl" marker lines the resource stub leaves are replaced by the files themselves (Shaders/convert/*.hlsl),
  * `cbuffer NAME : register(bN) {` -> `namespace {`, register / semantic annotations off, `[unroll(n)]` off, main -> ps_main,
  * hlg.hlsl's vector `c ? a : b` -> hlsl_select(c, a, b) (C++ has no overloadable ?:),
  * #undef of the text's own macros (w, dx, dy, ...) after each shader.
The texture register each Texture2D was declared with is recorded before the annotation is stripped, so the runner binds plane k of the
frame to register tk the way CDX11VideoProcessor binds its shader resource views.  Nothing of this output is kept in the repository."""
import ctypes as C
import re
import sys

REF, LIB = sys.argv[1], sys.argv[2]
R = C.CDLL(LIB)
R.ref_convert_shader.argtypes = [C.c_int, C.c_uint, C.c_long, C.c_long, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int,
                                 C.c_char_p, C.c_int]
R.ref_specify_extended_format.restype = C.c_uint32
R.ref_specify_extended_format.argtypes = [C.c_uint32, C.c_int, C.c_uint, C.c_uint]
R.ref_fmt_params.argtypes = [C.c_int, C.POINTER(C.c_int)]

CF = dict(NV12=1, P010=2, P016=3, YUY2=4, UYVY=5, P210=6, P216=7, Y210=8, Y216=9, AYUV=11, Y410=12, Y416=13, YV12=14, YV16=15, YV24=16,
          YUV420P8=17, YUV422P8=18, YUV444P8=19, YUV420P16=21, YUV422P16=23, YUV444P16=25, XRGB32=30, RGB48=33, Y8=37, Y16=39)
PACKED422 = ("YUY2", "UYVY", "Y210", "Y216")
W, H = 64, 32                                     # powers of two: texel centre / size and texcoord * size are exact


def exfmt(chroma=0, range_=0, matrix=0, prim=0, trc=0, sample=3):
    return (sample | (chroma << 8) | (range_ << 12) | (matrix << 15) | (prim << 22) | (trc << 27)) & 0xFFFFFFFF


SDR709, SDR2020, PQ, HLG = dict(range_=2, matrix=1, prim=2, trc=5), dict(range_=2, matrix=4, prim=9, trc=5), dict(range_=2, matrix=4, prim=9, trc=15), \
    dict(range_=2, matrix=4, prim=9, trc=16)
NEAREST, BILINEAR, CATMULL = 0, 1, 2
NONE, TO_SDR, TO_PQ = 0, 1, 2
# the RGB -> LMS matrix of the Dolby Vision cases (the text embeds dovi_lms2rgb x this as a literal): the ICtCp one scaled by 1/4096
DOVI_LMS = [1688 / 4096, 2146 / 4096, 262 / 4096, 683 / 4096, 2951 / 4096, 462 / 4096, 99 / 4096, 309 / 4096, 3688 / 4096]

CASES = []      # (name, format, exfmt fields, chroma siting, chroma scaler, convert type, blend deinterlace, dovi: -1 none / 0 poly / 1 mmr)
for fmt in ("NV12", "P010", "YV12", "YUV420P8", "YUV420P16"):
    for siting in (1, 5, 7):
        for scaler in (NEAREST, BILINEAR, CATMULL):
            CASES.append(("%s_s%d_c%d" % (fmt, siting, scaler), fmt, SDR709 if fmt != "P010" else PQ, siting, scaler, NONE, 0, -1))
for fmt in ("P210", "P216", "YV16", "YUV422P8", "YUV422P16", "YUY2", "UYVY", "Y210", "Y216"):
    for scaler in (NEAREST, BILINEAR, CATMULL):
        CASES.append(("%s_c%d" % (fmt, scaler), fmt, SDR709, 5, scaler, NONE, 0, -1))
for fmt in ("YV24", "YUV444P8", "YUV444P16", "AYUV", "Y410", "Y416", "XRGB32", "RGB48", "Y8", "Y16"):
    CASES.append(("%s" % fmt, fmt, SDR709, 5, BILINEAR, NONE, 0, -1))
for fmt in ("NV12", "YV12"):
    CASES.append(("%s_blend" % fmt, fmt, SDR709, 5, BILINEAR, NONE, 1, -1))
    CASES.append(("%s_blend_catmull" % fmt, fmt, SDR709, 5, CATMULL, NONE, 1, -1))
for fmt in ("P010", "Y410", "YUV420P16"):
    CASES.append(("%s_pq_to_sdr" % fmt, fmt, PQ, 5, BILINEAR, TO_SDR, 0, -1))
    CASES.append(("%s_hlg_to_sdr" % fmt, fmt, HLG, 5, BILINEAR, TO_SDR, 0, -1))
    CASES.append(("%s_hlg_to_pq" % fmt, fmt, HLG, 5, BILINEAR, TO_PQ, 0, -1))
    CASES.append(("%s_hlg_as_sdr" % fmt, fmt, HLG, 5, BILINEAR, NONE, 0, -1))
    CASES.append(("%s_pq_passthrough" % fmt, fmt, PQ, 5, BILINEAR, TO_PQ, 0, -1))
CASES.append(("YUY2_hlg_to_sdr", "YUY2", HLG, 5, CATMULL, TO_SDR, 0, -1))
CASES.append(("P016_2020_sdr", "P016", SDR2020, 5, BILINEAR, NONE, 0, -1))
for trc in (2, 3, 4, 5, 6, 7, 8, 14):
    CASES.append(("P010_2020_trc%d" % trc, "P010", dict(range_=2, matrix=4, prim=9, trc=trc), 5, BILINEAR, NONE, 0, -1))
CASES.append(("NV12_601_full", "NV12", dict(range_=1, matrix=2, prim=2, trc=5), 5, BILINEAR, NONE, 0, -1))
CASES.append(("P010_pq_to_sdr_1024x512", "P010", PQ, 5, BILINEAR, TO_SDR, 0, -1))      # a C2-shaped frame of some size for the whole-chain test
SIZES = {"P010_pq_to_sdr_1024x512": (1024, 512)}
for mmr in (0, 1):
    CASES.append(("P010_dovi%d_to_sdr" % mmr, "P010", PQ, 5, BILINEAR, TO_SDR, 0, mmr))
    CASES.append(("P010_dovi%d_to_pq" % mmr, "P010", PQ, 5, BILINEAR, TO_PQ, 0, mmr))


def include(name):
    s = open("%s/Shaders/convert/%s" % (REF, name)).read()
    if name == "hlg.hlsl":      # rgb = (cond) ? a : b;  on float3 -> hlsl_select(cond, a, b)
        s, n = re.subn(r"(\w+)\s*=\s*(\([^;?]*?\))\s*\?\s*([^;:]*?)\s*:\s*([^;]*?);", r"\1 = hlsl_select(\2, \3, \4);", s, count=1, flags=re.S)
        assert n == 1, "hlg.hlsl: the vector conditional was not found"
    return s


def translate(text):
    regs = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"Texture2D\s+(\w+)\s*:\s*register\(t(\d+)\)", text))
    text = re.sub(r"^//#include (\S+)$", lambda m: include(m.group(1)), text, flags=re.M)
    text = re.sub(r'^#include "\.\./convert/(\S+)"\s*$', lambda m: include(m.group(1)), text.replace("\r\n", "\n"), flags=re.M)
    text = re.sub(r"cbuffer\s+\w+\s*:\s*register\(\w+\)\s*\{", "namespace {", text)
    text = re.sub(r"\s*:\s*register\(\w+\)", "", text)
    text = re.sub(r"\s*:\s*(SV_POSITION|TEXCOORD|SV_Target)\b", "", text)
    text = re.sub(r"\[unroll\(\d+\)\]", "", text)
    text = re.sub(r"\bfloat4 main\(", "float4 ps_main(", text)
    macros = sorted(set(re.findall(r"^\s*#define\s+(\w+)", text, flags=re.M)))
    return text, regs, macros


print("// GENERATED by oracle/ref_stub/gen_convert_shaders.py from the text the reference's GetShaderConvertColor emits -- build product of")
print("// oracle/Makefile, never committed.  (hlsl_vec.h precedes this text.)")
print("struct ConvArgs { const float* tex[3]; int tw[3], th[3]; float cm[4][3]; float lum_scale; int W, H;")
print("                  const float* piv; const float* coef; const float* mmr; const int* par; const float* trims; int l2; };")
rows = []
for i, (name, fmt, exf, siting, scaler, conv, blend, dovi) in enumerate(CASES):
    cf = CF[fmt]
    W, H = SIZES.get(name, (64, 32))
    row = (C.c_int * 13)()
    assert R.ref_fmt_params(cf, row) == 0
    # only YUV media types carry their dwControlFlags into m_decExFmt (InitMediaType, DX11VideoProcessor.cpp:1763-1766): RGB and gray start from 0
    ex = R.ref_specify_extended_format(exfmt(chroma=siting, **exf) if row[3] == 0 else 0, cf, W, H)
    texw = W // 2 if fmt in PACKED422 else W
    buf = C.create_string_buffer(1 << 17)
    lms = (C.c_double * 9)(*DOVI_LMS)
    n = R.ref_convert_shader(1, W, texw, H, cf, ex, dovi, lms if dovi >= 0 else None, scaler, conv, blend, buf, len(buf))
    assert 0 < n < len(buf) - 1, name
    text, regs, macros = translate(buf.value.decode())
    ns = "cs_%s" % name
    print("// ======== %s: format %s exfmt %#010x chroma scaler %d convert type %d blend %d dovi %d" % (name, fmt, ex, scaler, conv, blend, dovi))
    print("namespace %s {" % ns)
    print(text)
    for m in macros:
        print("#undef %s" % m)
    print("static void setup(const ConvArgs& A) {              // constant buffers and bindings: once per draw")
    for t, r in sorted(regs.items(), key=lambda kv: kv[1]):
        print("    %s.bind(A.tex[%d], A.tw[%d], A.th[%d]);" % (t, r, r, r))
    print("    samp.linear = 0; sampL.linear = 1;")
    for k, row in enumerate(("cm_r", "cm_g", "cm_b", "cm_c")):
        print("    %s = float3(A.cm[%d][0], A.cm[%d][1], A.cm[%d][2]);" % (row, k, k, k))
    if "LuminanceScale" in text:
        print("    LuminanceScale = A.lum_scale; param2 = 0;")
    if dovi >= 0:
        print("    for (int c = 0; c < 3; c++) {")
        print("        for (int k = 0; k < 7; k++) curves[c].pivots_data[k] = A.piv[c * 7 + k];")
        print("        for (int k = 0; k < 8; k++) curves[c].coeffs_data[k] = float4(A.coef[(c * 8 + k) * 4], A.coef[(c * 8 + k) * 4 + 1], A.coef[(c * 8 + k) * 4 + 2], A.coef[(c * 8 + k) * 4 + 3]);")
        if dovi == 1:
            print("        for (int k = 0; k < 48; k++) curves[c].mmr_data[k] = float4(A.mmr[(c * 48 + k) * 4], A.mmr[(c * 48 + k) * 4 + 1], A.mmr[(c * 48 + k) * 4 + 2], A.mmr[(c * 48 + k) * 4 + 3]);")
            print("        curves[c].params.methods = A.par[c * 4]; curves[c].params.mmr_single = A.par[c * 4 + 1];")
            print("        curves[c].params.min_order = A.par[c * 4 + 2]; curves[c].params.max_order = A.par[c * 4 + 3];")
        print("    }")
        print("    ChromaWeight = A.trims[0]; SaturationGain = A.trims[1]; TrimSlope = A.trims[2]; TrimOffset = A.trims[3]; TrimPower = A.trims[4]; L2Enabled = A.l2;")
    print("}")
    print("static float4 shade(float u, float v) { PS_INPUT in; in.Tex = float2(u, v); return ps_main(in); }   // reads the globals only: callable from many threads")
    print("}  // namespace %s" % ns)
    rows.append('{ "%s", %d, %#010xu, %d, %d, %d, %d, %d, %d, %d, %s::setup, %s::shade }' % (name, cf, ex, siting, scaler, conv, blend, dovi, W, H, ns, ns))
# ---- Shaders/d3d11/ps_hdr10_tonemap.hlsl (the HDR10 local tone-mapping post-scale step), a file of the tree rather than generated text
text, regs, macros = translate(open("%s/Shaders/d3d11/ps_hdr10_tonemap.hlsl" % REF, encoding="utf-8", errors="replace").read())
print("// ======== ps_hdr10_tonemap.hlsl")
print("namespace tm_shader {")
print(text)
print("static float4 run(const float* rgba, int tw, int th, const float* hp, const float* l2, float u, float v) {")
print("    tex.bind(rgba, tw, th); samp.linear = 0;")
print("    MasteringMinLuminanceNits = hp[0]; MasteringMaxLuminanceNits = hp[1]; maxCLL = hp[2]; maxFALL = hp[3]; displayMaxNits = hp[4]; selection = (uint)hp[5];")
print("    L2Enabled = l2 ? 1u : 0u;")
print("    if (l2) { ChromaWeight = l2[0]; SaturationGain = l2[1]; TrimSlope = l2[2]; TrimOffset = l2[3]; TrimPower = l2[4]; }")
print("    PS_INPUT in; in.Tex = float2(u, v);")
print("    return ps_main(in);")
print("}")
print("}  // namespace tm_shader")
# ---- Shaders/examples/ps_resize_onepass_jinc2.hlsl (the "Jinc2m" upscaler, resource IDF_PS_11_INTERP_JINC2)
src = open("%s/Shaders/examples/ps_resize_onepass_jinc2.hlsl" % REF, encoding="utf-8", errors="replace").read()
src, n = re.subn(r"return \((x == float4\(0, 0, 0, 0\))\)\s*\?\s*(.*?)\s*:\s*(.*?);", r"return hlsl_select(\1, \2, \3);", src, count=1, flags=re.S)
assert n == 1, "jinc2: the vector conditional of resampler() was not found"
text, regs, macros = translate(src)
print("// ======== ps_resize_onepass_jinc2.hlsl")
print("namespace jinc2_shader {")
print(text)
for m in macros:
    print("#undef %s" % m)
print("static float4 run(const float* rgba, int tw, int th, float u, float v) {")
print("    tex.bind(rgba, tw, th); samp.linear = 0; wh = float2((float)tw, (float)th); dxdy = float2(1.0f / (float)tw, 1.0f / (float)th);")
print("    PS_INPUT in; in.Tex = float2(u, v);")
print("    return ps_main(in);")
print("}")
print("}  // namespace jinc2_shader")
# ---- the post-scale correction shaders of the D3D11-video-processor path (m_pPSCorrection, DX11VideoProcessor.cpp:1893-1931)
CORR = [("corr_convert_pq_to_sdr", "ps_convert_pq_to_sdr.hlsl", 1), ("corr_fixconvert_pq_to_sdr", "ps_fixconvert_pq_to_sdr.hlsl", 2),
        ("corr_fixconvert_hlg_to_sdr", "ps_fixconvert_hlg_to_sdr.hlsl", 3), ("corr_fix_bt2020", "ps_fix_bt2020.hlsl", 4),
        ("corr_convert_hlg_to_pq", "ps_convert_hlg_to_pq.hlsl", 5), ("corr_fix_ycgco", "ps_fix_ycgco.hlsl", 6)]
for ns, fname, mode in CORR:
    text, regs, macros = translate(open("%s/Shaders/d3d11/%s" % (REF, fname), encoding="utf-8", errors="replace").read())
    print("// ======== %s  <- Shaders/d3d11/%s" % (ns, fname))
    print("namespace %s {" % ns)
    print(text)
    for m in macros:
        print("#undef %s" % m)
    print("static float4 run(const float* rgba, int tw, int th, float lum, float u, float v) {")
    print("    tex.bind(rgba, tw, th); samp.linear = 0;")
    if "LuminanceScale" in text:
        print("    LuminanceScale = lum; param2 = 0;")
    print("    PS_INPUT in; in.Tex = float2(u, v);")
    print("    return ps_main(in);")
    print("}")
    # the matrices the text folds at shader-compile time, for a side-by-side look at the constants
    print("static void constants(float* o) {")
    if "fix_bt2020_matrix" in text:
        print("    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[i * 3 + j] = fix_bt2020_matrix[i][j];")
    elif "fix_ycgco_matrix" in text:
        print("    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[i * 3 + j] = fix_ycgco_matrix[i][j];")
    else:
        print("    for (int k = 0; k < 9; k++) o[k] = 0;")
    if "convert_matrix_2020_to_709" in text:
        print("    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[9 + i * 3 + j] = convert_matrix_2020_to_709[i][j];")
    else:
        print("    for (int k = 0; k < 9; k++) o[9 + k] = 0;")
    print("}")
    # ... and a way to run the per-pixel text with GIVEN matrices (the test separates "which constants" from "which operations")
    print("static void set_constants(const float* o) {")
    if "fix_bt2020_matrix" in text:
        print("    float4x4& F = const_cast<float4x4&>(fix_bt2020_matrix); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F[i][j] = o[i * 3 + j];")
    elif "fix_ycgco_matrix" in text:
        print("    float4x4& F = const_cast<float4x4&>(fix_ycgco_matrix); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F[i][j] = o[i * 3 + j];")
    if "convert_matrix_2020_to_709" in text:
        print("    float3x3& G = const_cast<float3x3&>(convert_matrix_2020_to_709); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) G[i][j] = o[9 + i * 3 + j];")
    print("}")
    print("}  // namespace %s" % ns)
print("struct CorrEntry { int mode; float4 (*run)(const float*, int, int, float, float, float); void (*constants)(float*); void (*set_constants)(const float*); };")
print("static const CorrEntry g_corr[] = { %s };" % ", ".join("{ %d, %s::run, %s::constants, %s::set_constants }" % (mode, ns, ns, ns) for ns, _, mode in CORR))
print("static const int g_ncorr = %d;" % len(CORR))

# ---- the Direct3D 9 twins of the resize shaders (Shaders/d3d9/*.hlsl): same filters, texel-space texcoords that carry -0.5
# (CDX9VideoProcessor::TextureResizeShader, DX9VideoProcessor.cpp:2853-2937), no duplicate Lanczos3 tap, no half-texel shift
def adapt9(src):
    src = src.replace('#include "../resize/convolution_filters.hlsl"', open("%s/Shaders/resize/convolution_filters.hlsl" % REF, encoding="utf-8", errors="replace").read())
    src = src.replace("\r\n", "\n")
    src = re.sub(r"\s*:\s*register\(\w+\)", "", src)
    src = re.sub(r"\s*:\s*(TEXCOORD0|COLOR)\b", "", src)
    src = src.replace("[loop] ", "")
    src = re.sub(r"\bfloat4 main\s*\(", "float4 ps_main(", src)
    src = re.sub(r"^static const float (\w+) = (.*);$", r"#define \1 (\2)", src, flags=re.M)
    return src


D9 = [("d9_mitchell4_x", "interpolation_spline4.hlsl", {"METHOD": "0", "AXIS": "0"}), ("d9_mitchell4_y", "interpolation_spline4.hlsl", {"METHOD": "0", "AXIS": "1"}),
      ("d9_catmull4_x", "interpolation_spline4.hlsl", {"METHOD": "1", "AXIS": "0"}), ("d9_catmull4_y", "interpolation_spline4.hlsl", {"METHOD": "1", "AXIS": "1"}),
      ("d9_lanczos2_x", "interpolation_lanczos2.hlsl", {"AXIS": "0"}), ("d9_lanczos2_y", "interpolation_lanczos2.hlsl", {"AXIS": "1"}),
      ("d9_lanczos3_x", "interpolation_lanczos3.hlsl", {"AXIS": "0"}), ("d9_lanczos3_y", "interpolation_lanczos3.hlsl", {"AXIS": "1"})]
for fname, fid, extra in (("box", "0", {}), ("bilinear", "1", {}), ("hamming", "2", {}), ("bicubic05", "3", {"A": "-0.5"}), ("bicubic15", "3", {"A": "-1.5"}), ("lanczos", "4", {})):
    for ax in ("0", "1"):
        dd = {"FILTER": fid, "AXIS": ax}; dd.update(extra)
        D9.append(("d9_conv_%s_%s" % (fname, "xy"[int(ax)]), "convolution.hlsl", dd))
SYMS9 = ["AXIS", "METHOD", "FILTER", "A", "PI", "filter_support", "support", "ss"]
for ns, fname, defs in D9:
    print("// ======== %s  <- Shaders/d3d9/%s  %s" % (ns, fname, " ".join("/D%s=%s" % kv for kv in defs.items())))
    for sym in SYMS9:
        print("#undef %s" % sym)
    for kv in defs.items():
        print("#define %s %s" % kv)
    print("namespace %s {" % ns)
    print(adapt9(open("%s/Shaders/d3d9/%s" % (REF, fname), encoding="utf-8", errors="replace").read()))
    print("static float4 run(const float* rgba, int tw, int th, float tx, float ty, float sx, float sy) {")
    print("    s0.bind(rgba, tw, th); dxdy = float2(1.0f / (float)tw, 1.0f / (float)th);")
    if fname == "convolution.hlsl":
        print("    scale = float2(sx, sy);")
    print("    return ps_main(float2(tx, ty));")
    print("}")
    print("}  // namespace %s" % ns)
for sym in SYMS9:
    print("#undef %s" % sym)
print("struct D9Entry { const char* name; float4 (*run)(const float*, int, int, float, float, float, float); };")
print("static const D9Entry g_d9[] = { %s };" % ", ".join('{ "%s", %s::run }' % (ns, ns) for ns, _, _ in D9))
print("static const int g_nd9 = %d;" % len(D9))
print("struct ConvCase { const char* name; int cformat; unsigned exfmt; int siting, scaler, convert_type, blend, dovi, W, H; void (*setup)(const ConvArgs&); float4 (*shade)(float, float); };")
print("static const ConvCase g_cases[] = {\n    %s\n};" % ",\n    ".join(rows))
print("static const int g_ncases = %d;" % len(rows))
print("static const double g_dovi_lms[9] = { %s };" % ", ".join(repr(float(x)) for x in DOVI_LMS))
