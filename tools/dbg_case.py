"""debug helper: one case on multipass / fused vs the oracle, prints where they differ"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import harness as H, orclib as O

case = H.Case(O.CF["P010"], 66, 34, 132, 68, H.EX_2020_PQ, backbuffer=O.SURF_BGRA8, tex=O.TEXFMT_8INT, dither=0, up=O.UPSCALE["Lanczos3"])
codes = H.make_codes(case.cf, case.w, case.h, "natural", 33)
buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
ref = H.run_oracle(case, planes, pitches)
for path in (1, 2, 0):
    vp = case.make_processor(device=0, path=path)
    vp.ProcessSample(data=buf, pitch=pitch)
    out = vp.ReadBackbuffer()
    a = H.rgb_view(out, case.backbuffer); b = H.rgb_view(ref, case.backbuffer)
    d = (a != b)
    print("path", path, "plan path", vp.debug_plan().path, "kernels", vp.LastFrameKernels(), "diff", int(d.sum()), "of", d.size)
    if d.sum():
        ys, xs, cs = np.nonzero(d)
        print("  rows", ys.min(), ys.max(), "cols", xs.min(), xs.max(), "out[0,:4]", a[0, :4].tolist(), "ref[0,:4]", b[0, :4].tolist())
        print("  out mean", a.mean(), "ref mean", b.mean())
