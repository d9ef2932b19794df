#!/usr/bin/env python3
"""Device-resident frames/s of the four single-GPU BASELINE configs on both kernel families (supplement to bench.py,
which reports the headline config C2).  Output: one JSON line per (config, path)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import harness as H, orclib as O, videorenderer_b200 as V

ALGO = {0: 1920*1080*1.5 + 1920*1080*4, 1: 3840*2160*3 + 7680*4320*4, 2: 1920*1080*2 + 2880*1620*4, 3: 7680*4320*3 + 3840*2160*8}
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for idx, case in enumerate(H.baseline_cases(1.0)):
    frames = []
    for i in range(4):
        codes = H.make_codes(case.cf, case.w, case.h, "noise", seed=i)
        buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
        frames.append((torch.from_numpy(buf).cuda(), pitch))
    f = H.fmt(case.cf)
    for path in (V.PATH.MULTIPASS, V.PATH.AUTO):
        vp = V.VideoProcessor(case.settings(), device=0, stream=stream.cuda_stream)
        assert vp.InitMediaType(V.VideoProcessor.media_type(case.cf, case.w, case.h, case.ex))
        vp.SetBackbufferFormat(case.backbuffer); vp.SetWindowRect((0, 0, case.dst_w, case.dst_h)); vp.SetVideoRect((0, 0, case.dst_w, case.dst_h))
        vp.debug_set_path(path)
        def step(i):
            t, pitch = frames[i % 4]; base = t.data_ptr()
            if f.layout == 1: vp.ProcessSample(device_planes=[base, base + pitch * case.h], device_pitches=[pitch, pitch])
            else: vp.ProcessSample(device_planes=[base], device_pitches=[pitch])
        for i in range(5): step(i)
        n = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        for i in range(n): step(i)
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        # the same frame re-rendered from C (no Python between launches): the kernels' own rate
        vp.debug_render_n(20); torch.cuda.synchronize()
        e0.record(stream); vp.debug_render_n(200); e1.record(stream); torch.cuda.synchronize()
        ms_c = e0.elapsed_time(e1) / 200
        gbs = ALGO[idx] / (ms * 1e-3) / 1e9
        print(json.dumps({"config": case.name, "path": {1: "multipass", 2: "fused"}[vp.debug_plan().path], "kernels_per_frame": vp.LastFrameKernels(),
                          "ms_per_frame": round(ms, 4), "frames_per_s": round(1e3 / ms, 1), "ms_per_frame_c_loop": round(ms_c, 4), "frames_per_s_c_loop": round(1e3 / ms_c, 1), "algorithmic_GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 4)}))
