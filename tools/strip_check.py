#!/usr/bin/env python3
"""tools/strip_check.py -- development probe for the persistent strip kernel (kernel_strip.cuh) on a GPU box:
(1) small/odd geometries against the oracle and against the tile kernel, (2) C2 frame time per warp-role split.
Usage: python tools/strip_check.py [--no-time] [--cfgs 0,1,2]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness as H
import orclib as O
import videorenderer_b200 as V

PATH_FUSED = 2
MASKS = []
WAITS = []
REPEAT = 1


def run(case, buf, pitch, strip, cfg=0):
    vp = case.make_processor(device=0, path=PATH_FUSED)
    vp.debug_set_strip_kernel(strip, cfg)
    vp.ProcessSample(data=buf, pitch=pitch)
    out = vp.ReadBackbuffer()
    vp.debug_set_strip_kernel(1, 0)
    return out


def correctness():
    cases = []
    P = O.CF
    for (w, h) in ((150, 70), (128, 64), (126, 62), (130, 66), (258, 38), (64, 36), (2, 2), (10, 130), (386, 30), (512, 300)):
        cases.append(H.Case(P["P010"], w, h, 2 * w, 2 * h, H.EX_2020_PQ, up=O.UPSCALE["Lanczos3"], name="p010_%dx%d" % (w, h)))
    cases.append(H.Case(P["NV12"], 96, 64, 160, 104, H.EX_709, src_rect=(8, 6, 88, 58), up=O.UPSCALE["Lanczos3"], name="nv12_crop"))
    cases.append(H.Case(P["NV12"], 96, 64, 2 * 77, 2 * 51, H.EX_709, src_rect=(3, 5, 80, 56), up=O.UPSCALE["CatmullRom"], name="nv12_oddcrop"))
    cases.append(H.Case(P["P010"], 64, 36, 128, 72, H.EX_2020_PQ, dst_rect=(16, 11, 144, 83), win=(160, 100), up=O.UPSCALE["Lanczos3"], name="letterbox"))
    cases.append(H.Case(P["P210"], 200, 50, 400, 100, H.EX_2020_PQ, up=O.UPSCALE["Lanczos2"], name="p210"))
    cases.append(H.Case(P["P016"], 140, 44, 280, 88, H.EX_2020_HLG, up=O.UPSCALE["Mitchell"], name="p016_hlg"))
    for siting in (1, 5, 7):
        ex = O.exfmt(chroma=siting, range_=2, matrix=1, prim=2, trc=5)
        cases.append(H.Case(P["NV12"], 140, 52, 280, 104, ex, up=O.UPSCALE["Lanczos3"], variant=O.VARIANT_D3D9, name="nv12_siting%d" % siting))
    cases.append(H.Case(P["P010"], 132, 34, 264, 68, H.EX_2020_PQ, backbuffer=O.SURF_RGB10A2, tex=O.TEXFMT_16FLOAT, up=O.UPSCALE["Lanczos3"], name="p010_16f_rgb10"))
    cases.append(H.Case(P["P010"], 132, 34, 264, 68, H.EX_2020_PQ, backbuffer=O.SURF_RGBA16F, tex=O.TEXFMT_16FLOAT, dither=0, up=O.UPSCALE["Lanczos3"], name="p010_16f_16f"))
    cases.append(H.Case(P["P010"], 132, 34, 264, 68, H.EX_2020_PQ, tex=O.TEXFMT_8INT, dither=0, up=O.UPSCALE["Lanczos3"], name="p010_8int"))
    bad = 0
    for c in cases:
        codes = H.make_codes(c.cf, c.w, c.h, "fullnoise", 5)
        buf, pitch, planes, pitches = H.pack(c.cf, c.w, c.h, codes)
        ref = H.run_oracle(c, planes, pitches)
        try:
            a = run(c, buf, pitch, 1)
        except Exception as e:
            print("CASE %-18s strip FAILED: %s" % (c.name, e)); bad += 1
            if "llegal" in str(e) or "create" in str(e):
                break
            continue
        st = V.VideoProcessor(c.settings(), device=0).debug_strip_status()
        if st[0]:
            print("   WATCHDOG: %d waits timed out; first: barrier slot %d parity %d iteration %d block %d warp %d" % tuple(st[:6]))
        b = run(c, buf, pitch, 0)
        n1, d1 = H.compare(a, ref, c.backbuffer); n2, d2 = H.compare(b, ref, c.backbuffer)
        print("CASE %-18s strip vs oracle: %d (max %d)   tile vs oracle: %d" % (c.name, n1, d1, n2))
        if n1:
            bad += 1
            ra, rb = H.rgb_view(a, c.backbuffer), H.rgb_view(ref, c.backbuffer)
            ys, xs, _ = np.nonzero(ra.astype(np.int64) != rb.astype(np.int64)) if ra.dtype != np.float16 else np.nonzero(ra.view(np.int16) != rb.view(np.int16))
            print("   mismatch rows %d..%d cols %d..%d; first: y=%d x=%d got %s want %s" % (ys.min(), ys.max(), xs.min(), xs.max(), ys[0], xs[0], ra[ys[0], xs[0]], rb[ys[0], xs[0]]))
    print("correctness: %d failing cases of %d" % (bad, len(cases)))
    return bad


def timing(cfgs, steps=300):
    import torch
    case = H.baseline_cases(1.0)[1]
    frames = []
    for i in range(4):
        codes = H.make_codes(case.cf, case.w, case.h, "natural", seed=100 + i)
        buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
        frames.append((torch.from_numpy(buf).cuda(), pitch, buf))
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    vp = V.VideoProcessor(case.settings(), device=0, stream=stream.cuda_stream)
    mt = V.VideoProcessor.media_type(case.cf, case.w, case.h, case.ex)
    assert vp.InitMediaType(mt)
    vp.SetWindowRect((0, 0, case.dst_w, case.dst_h)); vp.SetVideoRect((0, 0, case.dst_w, case.dst_h))
    vp.debug_set_path(PATH_FUSED)
    ref_out = None
    res = {}
    runs = [("tile", 0, 0)] + [("strip%d" % c, 1, c) for c in cfgs]
    runs += [("wait%d" % (16 << w), 1, (w << 12)) for w in WAITS]          # sleep-between-polls wait mode, ns
    runs += [("skip%x" % m, 1, (m << 8)) for m in MASKS if m < 16]      # role-balance experiment: bit set = that role (P, A, B, C) only handshakes
    rounds = {}
    for rep in range(REPEAT):
        for name, strip, cfg in runs:
            vp.debug_set_strip_kernel(strip, cfg)
            def step(i):
                t, pitch, _ = frames[i % 4]
                vp.ProcessSample(device_planes=[t.data_ptr(), t.data_ptr() + pitch * case.h], device_pitches=[pitch, pitch])
            try:
                for i in range(8): step(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for i in range(steps): step(i)
                e1.record(stream); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / steps * 1e3
                step(0); out = vp.ReadBackbuffer()
                if name.startswith("skip"): same = "(partial work)"
                elif ref_out is None or name == "tile": ref_out = out; same = "ref"
                elif np.array_equal(out, ref_out): same = "equal"
                else:
                    d = np.abs(out.astype(np.int16) - ref_out.astype(np.int16))
                    same = "DIFFERENT: %d of %d bytes (%.4f %%), |diff| 1: %d, 2: %d, >2: %d, max %d" % (int((d != 0).sum()), d.size, 100.0 * float((d != 0).mean()),
                                                                                                 int((d == 1).sum()), int((d == 2).sum()), int((d > 2).sum()), int(d.max()))
                res[name] = min(us, res.get(name, 1e9)); rounds.setdefault(name, []).append(round(us, 1))
                print("TIME %-8s %8.1f us/frame  %8.1f fps   output %s" % (name, us, 1e6 / us, same), flush=True)
            except Exception as e:
                print("TIME %-8s FAILED: %s" % (name, e), flush=True)
                break
    vp.debug_set_strip_kernel(1, 0)
    if REPEAT > 1:
        for k, v in rounds.items(): print("ROUNDS %-8s min %.1f  all %s" % (k, min(v), v))
    print(json.dumps(res))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-time", action="store_true"); ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--cfgs", default="0,1,2,3")
    ap.add_argument("--waits", default="", help="comma separated wait codes w (sleep 16<<w ns between polls)")
    ap.add_argument("--masks", default="", help="comma separated hex role-skip masks to time (bit0 P, bit1 A, bit2 B, bit3 C)")
    ap.add_argument("--one", action="store_true", help="run one small case on the strip kernel and print the watchdog record")
    ap.add_argument("--repeat", type=int, default=1, help="time every configuration this many times (interleaved rounds), report the minimum")
    a = ap.parse_args()
    globals()["REPEAT"] = a.repeat
    WAITS[:] = [int(x) for x in a.waits.split(",") if x]
    MASKS[:] = [int(x, 16) for x in a.masks.split(",") if x]
    if a.one:
        c = H.Case(O.CF["P010"], 150, 70, 300, 140, H.EX_2020_PQ, up=O.UPSCALE["Lanczos3"], name="one")
        codes = H.make_codes(c.cf, c.w, c.h, "fullnoise", 5)
        buf, pitch, planes, pitches = H.pack(c.cf, c.w, c.h, codes)
        ref = H.run_oracle(c, planes, pitches)
        out = run(c, buf, pitch, 1)
        print("one: mismatches", H.compare(out, ref, c.backbuffer), "status", V.VideoProcessor(c.settings(), device=0).debug_strip_status())
        sys.exit(0)
    bad = 0 if a.no_check else correctness()
    if not a.no_time:
        timing([int(x) for x in a.cfgs.split(",")])
    sys.exit(1 if bad else 0)
