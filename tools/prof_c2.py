#!/usr/bin/env python3
"""tools/prof_c2.py -- a few C2 frames through the fused path with device-resident input, for ncu captures.
Usage: python tools/prof_c2.py [--cfg N] [--frames K] [--tile] [--config 1|2|3|4]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import harness as H
import videorenderer_b200 as V

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", type=int, default=0); ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--tile", action="store_true"); ap.add_argument("--config", type=int, default=2)
ap.add_argument("--path", default="auto")
a = ap.parse_args()
case = H.baseline_cases(1.0)[a.config - 1]
codes = H.make_codes(case.cf, case.w, case.h, "natural", seed=100)
buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
vp = case.make_processor(device=0, path={"auto": 0, "multipass": 1, "fused": 2}[a.path])
vp.debug_set_strip_kernel(0 if a.tile else 1, a.cfg)
for i in range(a.frames):
    vp.ProcessSample(data=buf, pitch=pitch)
torch.cuda.synchronize()
print("done: path", vp.debug_plan().path, "kernels/frame", vp.LastFrameKernels())
