#!/usr/bin/env python3
"""oracle/ref_stub/gen_hlsl_shaders.py <reference root> -- TEST INFRASTRUCTURE ONLY.
Writes to stdout ONE C++ translation unit that contains the reference's pixel-shader files of the resize and final passes, each in a
namespace of its own and with the preprocessor symbols its build script gives it (Shaders/compile_shaders.cmd:82-107), so that g++ can
compile them (with oracle/ref_stub/hlsl_compat.h in front) and tests can execute them.  The shader text is read from the reference tree at
build time and passed through; the only edits are to HLSL's binding syntax, which has no C++ spelling and carries no arithmetic:
    `: register(..)`, `: SV_POSITION`, `: TEXCOORD`, `: SV_Target`   removed
    `cbuffer NAME`                                                    -> `namespace` (its members become globals, as in HLSL)
    `[loop]`                                                          removed
    `float4 main(`                                                    -> `float4 ps_main(`
    `static const float NAME = EXPR;` at file scope (ps_convolution)  -> `#define NAME (EXPR)`  (HLSL evaluates them per invocation from the
                                                                         constant buffer; a C++ static would be initialised once, at load)
    `#include "../resize/convolution_filters.hlsl"`                   -> the file's text, in place
Nothing is written into the repository."""
import os, re, sys

REF = sys.argv[1]
D11 = os.path.join(REF, "Shaders", "d3d11")


def text(path):
    return open(path, encoding="utf-8", errors="replace").read().replace("\r\n", "\n")


def adapt(src):
    src = src.replace('#include "../resize/convolution_filters.hlsl"', text(os.path.join(REF, "Shaders", "resize", "convolution_filters.hlsl")))
    src = re.sub(r"\s*:\s*register\([a-z0-9]+\)", "", src)
    src = re.sub(r"\s*:\s*(SV_POSITION|TEXCOORD|SV_Target)\b", "", src)
    src = re.sub(r"^cbuffer\s+\w+\s*$", "namespace", src, flags=re.M)
    src = src.replace("[loop] ", "")
    src = src.replace("float4 main(", "float4 ps_main(")
    src = re.sub(r"^static const float (\w+) = (.*);$", r"#define \1 (\2)", src, flags=re.M)
    return src


SHADERS = [  # (namespace, file, defines)  -- Shaders/compile_shaders.cmd:82-107
    ("mitchell4_x", "ps_interpolation_spline4.hlsl", {"METHOD": "0", "AXIS": "0"}), ("mitchell4_y", "ps_interpolation_spline4.hlsl", {"METHOD": "0", "AXIS": "1"}),
    ("catmull4_x", "ps_interpolation_spline4.hlsl", {"METHOD": "1", "AXIS": "0"}), ("catmull4_y", "ps_interpolation_spline4.hlsl", {"METHOD": "1", "AXIS": "1"}),
    ("lanczos2_x", "ps_interpolation_lanczos2.hlsl", {"AXIS": "0"}), ("lanczos2_y", "ps_interpolation_lanczos2.hlsl", {"AXIS": "1"}),
    ("lanczos3_x", "ps_interpolation_lanczos3.hlsl", {"AXIS": "0"}), ("lanczos3_y", "ps_interpolation_lanczos3.hlsl", {"AXIS": "1"}),
]
for i, (name, extra) in enumerate((("box", {}), ("bilinear", {}), ("hamming", {}), ("bicubic05", {"A": "-0.5"}), ("bicubic15", {"A": "-1.5"}), ("lanczos", {}))):
    f = {"box": "0", "bilinear": "1", "hamming": "2", "bicubic05": "3", "bicubic15": "3", "lanczos": "4"}[name]
    for ax in ("0", "1"):
        d = {"FILTER": f, "AXIS": ax}; d.update(extra)
        SHADERS.append(("conv_%s_%s" % (name, "xy"[int(ax)]), "ps_convolution.hlsl", d))
SHADERS += [("final_8", "ps_final_pass.hlsl", {"QUANTIZATION": "255"}), ("final_10", "ps_final_pass.hlsl", {"QUANTIZATION": "1023"})]

SYMS = ["AXIS", "METHOD", "FILTER", "A", "QUANTIZATION", "PI", "filter_support", "support", "ss"]
out = []
for ns, fname, defs in SHADERS:
    out.append("// ======== %s  <- Shaders/d3d11/%s  %s" % (ns, fname, " ".join("/D%s=%s" % kv for kv in defs.items())))
    out += ["#undef %s" % s for s in SYMS]
    out += ["#define %s %s" % kv for kv in defs.items()]
    out.append("namespace %s {" % ns)
    out.append(adapt(text(os.path.join(D11, fname))))
    out.append("}  // namespace %s" % ns)
out += ["#undef %s" % s for s in SYMS]
# the dispatch table the wrapper uses
out.append("struct ShaderEntry { const char* name; void (*setup)(const float* tex, int w, int h, int wrap, const float* tex2, int w2, int h2, float sx, float sy); float4 (*shade)(float u, float v); };")
for ns, fname, defs in SHADERS:
    final = fname == "ps_final_pass.hlsl"
    body = ("%s::tex.bind(tex, w, h); %s::samp.wrap = wrap; " % (ns, ns))
    if final:
        body += "%s::texDither.bind(tex2, w2, h2); %s::sampDither.wrap = 1; %s::ditherCoordScale = float2(sx, sy); " % (ns, ns, ns)
    else:
        body += "%s::wh = float2((float)w, (float)h); %s::dxdy = float2(1.0f / (float)w, 1.0f / (float)h); %s::scale = float2(sx, sy); " % (ns, ns, ns)
    out.append("static void setup_%s(const float* tex, int w, int h, int wrap, const float* tex2, int w2, int h2, float sx, float sy) { %s }" % (ns, body))
    out.append("static float4 shade_%s(float u, float v) { %s::PS_INPUT in; in.Tex = float2(u, v); return %s::ps_main(in); }" % (ns, ns, ns))
out.append("static const ShaderEntry g_shaders[] = { %s };" % ", ".join('{ "%s", setup_%s, shade_%s }' % (ns, ns, ns) for ns, _, _ in SHADERS))
out.append("static const int g_nshaders = %d;" % len(SHADERS))
sys.stdout.write("\n".join(out) + "\n")
