#!/usr/bin/env python3
"""Mint tests/golden/frames_full.json: sha256 of the oracle's output (SPEC mode) for the four single-GPU BASELINE configs at the
sizes BASELINE.json names (1920x1080 NV12; 3840x2160 P010 -> 7680x4320; 1920x1080 YUY2 -> 2880x1620; 7680x4320 P016 -> 3840x2160),
"natural" frame, seed 21.  The GPU tests compare the product's full frames with the oracle pixel by pixel AND with these hashes, so a
drift of the oracle itself between machines / compilers shows up too.  Run here: python tests/golden/make_golden_fullsize.py"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H   # noqa: E402


def frame_hash(out, fmt_):
    rgb = np.ascontiguousarray(H.rgb_view(out, fmt_))
    if rgb.dtype == np.float16:
        rgb = rgb.view(np.uint16)
    return hashlib.sha256(rgb.tobytes()).hexdigest(), list(rgb.shape), str(rgb.dtype)


def main():
    frames = {}
    for case in H.baseline_cases(1.0):
        t0 = time.time()
        codes = H.make_codes(case.cf, case.w, case.h, "natural", seed=21)
        buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
        out = H.run_oracle(case, planes, pitches)
        h, shape, dt = frame_hash(out, case.backbuffer)
        frames[case.name] = {"kind": "natural", "seed": 21, "in_sha256": hashlib.sha256(buf.tobytes()).hexdigest(), "out_rgb_sha256": h, "shape": shape, "dtype": dt}
        print(case.name, shape, "%.1f s" % (time.time() - t0), flush=True)
    json.dump(frames, open(os.path.join(HERE, "frames_full.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
