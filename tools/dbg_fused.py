import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, harness as H, orclib as O
case = H.Case(O.CF["P010"], 150, 70, 300, 140, H.EX_2020_PQ, up=O.UPSCALE["CatmullRom"], dither=0, tex=O.TEXFMT_10INT, backbuffer=O.SURF_RGB10A2, name="dbg")
codes = H.make_codes(case.cf, case.w, case.h, "noise", 31)
buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
outs = {}
for path in (1, 2):
    vp = case.make_processor(device=0, path=path)
    vp.ProcessSample(data=buf, pitch=pitch)
    outs[path] = H.rgb_view(vp.ReadBackbuffer(), case.backbuffer)
d = np.abs(outs[1] - outs[2])
print("mismatch", (d != 0).sum(), "of", d.size, "max", d.max())
ys, xs, cs = np.nonzero(d)
print("x parity hist", np.bincount(xs % 4, minlength=4), "y%4", np.bincount(ys % 4, minlength=4), "chan", np.bincount(cs, minlength=3))
print("x hist /16:", np.bincount(xs // 16))
print("first", list(zip(ys[:10], xs[:10], cs[:10])), outs[1][ys[:5], xs[:5], cs[:5]], outs[2][ys[:5], xs[:5], cs[:5]])
