#!/usr/bin/env python3
"""Mint the golden vectors the reference itself lacks (it ships no tests).

Run HERE (needs /root/reference for the matrices); commits small JSON fixtures:
  kat_matrices.json    mp_get_csp_matrix / GetColorspaceGamutConversionMatrix outputs of the REFERENCE's own
                       csputils.cpp (compiled in place into oracle/_ref/libref_csputils.so), as fp32 bit patterns
  kat_weights.json     resize weight vectors from the oracle, as fp32 bit patterns
  kat_transfer.json    scalar PQ->SDR curve samples in SPEC and F64 arithmetic
  dither_stage.json    sha256 of the exhaustive (code10 x threshold) -> u8 table of the final pass + the 4 pairs where
                       fp32 differs from the exact rational floor
  frames.json          sha256 of oracle outputs (SPEC mode) for the scaled BASELINE configs x {ramp, noise, natural}
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H   # noqa: E402
import orclib as O    # noqa: E402


def bits(a):
    return [int(v) for v in np.asarray(a, np.float32).ravel().view(np.uint32)]


def main():
    R = O.ref()
    assert R is not None, "build oracle/_ref first (make -C oracle)"
    mats = {}
    # (mp_csp, levels, bits) as set_colorspace would produce them; default procamp
    for name, (space, levels, depth) in {"bt601_tv_8": (1, 1, 8), "bt709_tv_8": (2, 1, 8), "bt709_pc_8": (2, 2, 8), "smpte240m_tv_8": (3, 1, 8),
                                         "bt2020nc_tv_16": (4, 1, 16), "bt2020nc_tv_10": (4, 1, 10), "bt2020nc_pc_16": (4, 2, 16),
                                         "ycgco_pc_8": (8, 2, 8), "bt709_tv_16": (2, 1, 16)}.items():
        m = np.zeros(9, np.float32); c = np.zeros(3, np.float32)
        R.ref_csp_matrix(space, levels, 0.0, 1.0, 0.0, 1.0, 0, depth, m.ctypes.data, c.ctypes.data)
        mats[name] = {"space": space, "levels": levels, "bits": depth, "m_bits": bits(m), "c_bits": bits(c), "m": [float(x) for x in m], "c": [float(x) for x in c]}
    # procamp variants (brightness/255, contrast, hue rad, saturation as the reference call site passes them)
    pa = {}
    for name, (b, ct, hue_deg, sat) in {"pa1": (12.0, 1.15, 25.0, 0.8), "pa2": (-30.0, 0.7, -90.0, 1.6)}.items():
        m = np.zeros(9, np.float32); c = np.zeros(3, np.float32)
        hue = np.float32(np.float32(hue_deg) / np.float32(180)) * np.arccos(-1.0)
        R.ref_csp_matrix(2, 1, float(np.float32(b) / np.float32(255)), ct, float(np.float32(hue)), sat, 0, 8, m.ctypes.data, c.ctypes.data)
        pa[name] = {"brightness": b, "contrast": ct, "hue": hue_deg, "saturation": sat, "m_bits": bits(m), "c_bits": bits(c)}
    g = np.zeros(9, np.float32); R.ref_gamut_matrix(4, 3, g.ctypes.data)
    json.dump({"source": "reference Source/csputils.cpp compiled in place (oracle/Makefile)", "matrices": mats, "procamp_bt709_tv_8": pa,
               "gamut_2020_to_709_bits": bits(g), "gamut_2020_to_709": [float(x) for x in g]}, open(os.path.join(HERE, "kat_matrices.json"), "w"), indent=1)

    w = {}
    for name, (kind, filt, src, dst, idx) in {"lanczos3_2x_i0": (1, 4, 3840, 7680, 0), "lanczos3_2x_i1": (1, 4, 3840, 7680, 1),
                                              "catmull_1.5x_i0": (1, 2, 1920, 2880, 0), "catmull_1.5x_i1": (1, 2, 1920, 2880, 1), "catmull_1.5x_i2": (1, 2, 1920, 2880, 2),
                                              "mitchell_2x_i0": (1, 1, 100, 200, 0), "lanczos2_2x_i1": (1, 3, 100, 200, 1),
                                              "bicubic_0.5x_i10": (2, 3, 7680, 3840, 10), "hamming_0.4x_i7": (2, 2, 500, 200, 7),
                                              "lanczos_0.5x_i3": (2, 5, 400, 200, 3), "box_0.25x_i5": (2, 0, 400, 100, 5)}.items():
        for variant in (0, 1):
            base, ws, wsum, n = O.axis_taps(kind, filt, variant, 0, src, dst, idx)
            w["%s_%s" % (name, "d3d11" if variant == 0 else "d3d9")] = {"kind": kind, "filter": filt, "src": src, "dst": dst, "i": idx, "variant": variant,
                                                                        "base": base, "w_bits": bits(ws), "w": [float(x) for x in ws], "wsum_bits": bits([wsum])[0]}
    json.dump(w, open(os.path.join(HERE, "kat_weights.json"), "w"), indent=1)

    # scalar transfer samples through the whole chain: gray P010-like input codes
    L = O.lib()
    pq = {}
    for v in (0.0, 0.1, 0.25, 0.5, 0.508, 0.58, 0.75, 1.0):
        pq[str(v)] = {"spec_bits": bits([L.orc_test_pow(float(np.float32(v)), float(np.float32(1.0) / np.float32(78.84375)), 0)])[0],
                      "f64": float(L.orc_test_pow(float(np.float32(v)), 1.0 / 78.84375, 2))}
    json.dump({"pow(v,1/m2)": pq}, open(os.path.join(HERE, "kat_transfer.json"), "w"), indent=1)

    # final-pass stage, exhaustive: out[c][k] = floor(fl(c/1023)*255 + k/1024)  (fp32, fused multiply-add == mul-then-add)
    c = (np.arange(1024, dtype=np.float32) / np.float32(1023))[:, None]
    k = (np.arange(1024, dtype=np.float32) / np.float32(1024))[None, :]
    f32 = np.floor(c * np.float32(255) + k).astype(np.uint8)
    exact = ((np.arange(1024, dtype=np.int64)[:, None] * 255 * 1024 + np.arange(1024, dtype=np.int64)[None, :] * 1023) // (1023 * 1024)).astype(np.uint8)
    diff = np.argwhere(f32 != exact)
    json.dump({"table_sha256": hashlib.sha256(f32.tobytes()).hexdigest(), "fp32_vs_exact_rational_pairs": diff.tolist(),
               "fp32_values_at_pairs": [int(f32[a, b]) for a, b in diff]}, open(os.path.join(HERE, "dither_stage.json"), "w"), indent=1)

    frames = {}
    for case in H.baseline_cases(scale=0.125):
        for kind in ("ramp", "noise", "natural"):
            codes = H.make_codes(case.cf, case.w, case.h, kind, seed=1)
            buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
            out = H.run_oracle(case, planes, pitches)
            rgb = np.ascontiguousarray(H.rgb_view(out, case.backbuffer))
            frames["%s/%s" % (case.name, kind)] = {"in_sha256": hashlib.sha256(buf.tobytes()).hexdigest(),
                                                  "out_rgb_sha256": hashlib.sha256(rgb.tobytes()).hexdigest(), "shape": list(rgb.shape), "dtype": str(rgb.dtype)}
    json.dump(frames, open(os.path.join(HERE, "frames.json"), "w"), indent=1)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
