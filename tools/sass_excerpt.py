#!/usr/bin/env python3
"""tools/sass_excerpt.py -> profiles/r02b_sass_kernels.txt
Mnemonic census (cuobjdump -sass of the built library) of the kernels the bench configs run, with the first lines around the
TMA / mbarrier / 128-bit memory instructions, so the instruction-level claims in DESIGN.md can be checked without a GPU."""
import collections, hashlib, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "videorenderer_b200", "libmpcvr_b200.so")
WANT = [("k_strip2x (C2 frame kernel, product configuration: 32 warps, OPT 19 = gamma table + fused final pass + setmaxnreg)", r"k_strip2xINS_8StripCfgILi128ELi8ELi8ELi12ELi4ELi8ELi19EEELi6ELb1ELi1ELi0E"), ("k_fused1x_bp<1, 8> (C1 frame kernel)", r"k_fused1x_bpILi1ELi8E"),
        ("k_fused2x (round-1 tile kernel, 6 taps)", r"k_fused2xILi6E"), ("k_convert_bp<2,false> (C4 convert pass)", r"k_convert_bpILi2ELb0E"), ("k_convert_bp<1,true> (C3 convert pass, YUY2)", r"k_convert_bpILi1ELb1E"),
        ("k_resize_p<9,16F,16F,X,conv> (C4 X pass)", r"k_resize_pILi9ELi2ELi2ELi0ELb1E"), ("k_resize_p<4,16F,16F,Y,interp> (C3 Y pass)", r"k_resize_pILi4ELi2ELi2ELi1ELb0E"), ("k_final_p<16F,RGB10A2> (C3 final pass)", r"k_final_pILi2ELi1E")]
KEY = ["UTMALDG", "UTMAPF", "USETMAXREG", "SYNCS", "NANOSLEEP", "BAR.SYNC", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "PRMT", "LDG.E.128", "LDG.E.64", "LDG.E.U16", "LDG.E.U8", "STG.E.128", "STG.E.64", "STG.E",
       "LDS.128", "LDS.64", "LDS", "STS.128", "STS.64", "STS", "SHFL", "PMTRIG", "I2F", "F2I", "FRND", "BRA"]

def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = collections.OrderedDict(); cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m: cur = m.group(1); funcs[cur] = []; continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): funcs[cur].append(ln.rstrip())
    out = ["# cuobjdump -sass videorenderer_b200/libmpcvr_b200.so (sm_100a), sha256 " + hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16], ""]
    for title, pat in WANT:
        hits = [f for f in funcs if re.search(pat, f)]
        if not hits: out.append(f"## {title}: no function matches {pat}\n"); continue
        f = hits[0]
        body = funcs[f]
        ops = collections.Counter()
        for ln in body:
            m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", ln)
            if m: ops[m.group(1)] += 1
        out.append(f"## {title}\n{f}\n{len(body)} SASS instructions ({len(body) * 16} bytes)")
        row = []
        for k in KEY:
            n = sum(v for o, v in ops.items() if o == k or o.startswith(k + "."))
            if n: row.append(f"{k} {n}")
        out.append("census: " + ", ".join(row))
        for cls in ["UTMALDG", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "NANOSLEEP", "BAR.SYNC", "PMTRIG", "LDG.E.128", "STG.E.128", "STG.E.64", "FFMA2", "FFMA.SAT", "PRMT"]:
            shown = 0
            for ln in body:
                if cls in ln and shown < 3:
                    out.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", ln)); shown += 1
        out.append("")
    open(os.path.join(ROOT, "profiles", "r02b_sass_kernels.txt"), "w").write("\n".join(out))
    print("\n".join(out)[:6000])

if __name__ == "__main__":
    main()
