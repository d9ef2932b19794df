"""SPEC (the contract the CUDA path matches bit for bit) against the literal fp32 and the float64 restatements, for the
"next" rows of SURVEY 8(f): the contract's fused steps / own sin / pow must not move the picture by more than rounding noise."""
import ctypes as C

import numpy as np
import pytest

import harness as H
import orclib as O


def _modes(case, kind="natural", seed=5):
    codes = H.make_codes(case.cf, case.w, case.h, kind, seed)
    buf, pitch, planes, pitches = H.pack(case.cf, case.w, case.h, codes)
    return [H.run_oracle(case, planes, pitches, mode=m) for m in (O.MODE_SPEC, O.MODE_LITERAL, O.MODE_F64)]


def _close(case, frac=0.02, maxd=1, **kw):
    spec, lit, f64 = _modes(case, **kw)
    for other, tag in ((lit, "literal"), (f64, "f64")):
        n, d = H.compare(spec, other, case.backbuffer)
        if case.backbuffer == O.SURF_RGBA16F:
            assert d <= 4, (case.name, tag, d)                       # fp16 ulps
        else:
            assert d <= maxd and n <= frac * spec.size, (case.name, tag, n, d)


@pytest.mark.parametrize("tm", [1, 2, 3, 4, 5, 6])
def test_hdr10_tonemap_curves(tm):
    hdr = dict(out=1, tonemap=1, min=50.0, max=4000.0, cll=1200.0, fall=300.0, display=600.0, type=tm)
    _close(H.Case(O.CF["P010"], 64, 36, 96, 54, H.EX_2020_PQ, backbuffer=O.SURF_RGB10A2, up=O.UPSCALE["Lanczos3"], hdr=hdr, name="tm%d" % tm), maxd=2)


def test_jinc2_and_rotation():
    _close(H.Case(O.CF["NV12"], 64, 36, 128, 72, H.EX_709, up=O.UPSCALE["Jinc2"], name="jinc2"))
    _close(H.Case(O.CF["NV12"], 64, 36, 54, 96, H.EX_709, up=O.UPSCALE["Lanczos3"], rotation=90, flip=1, name="rot90"))
    _close(H.Case(O.CF["NV12"], 160, 90, 31, 50, H.EX_709, down=O.DOWNSCALE["Hamming"], rotation=270, name="rot270_down"))


@pytest.mark.parametrize("mmr", [0, 2])
def test_dolby_vision(mmr):
    ex = O.exfmt(chroma=5, range_=1, matrix=0, prim=0, trc=0)
    lo = (2081, 2100, 2030, 2120, 2010, 2075); hi = (3388, 2070, 2040, 2090, 2060, 2000)
    _close(H.Case(O.CF["P010"], 64, 36, 96, 54, ex, up=O.UPSCALE["CatmullRom"], dovi=H.make_dovi(7, mmr=mmr, l2=[lo, hi], source_max_pq=3696),
                  hdr=dict(out=0, display=1000), name="dovi%d" % mmr), frac=0.03, maxd=2)


@pytest.mark.parametrize("mode", [1, 2, 3, 4, 5])
def test_correction_shaders(mode):
    bb = O.SURF_RGB10A2 if mode == 5 else O.SURF_BGRA8
    _close(H.Case(O.CF["r210"], 64, 36, 96, 54, 0, up=O.UPSCALE["CatmullRom"], correction=mode, backbuffer=bb, name="corr%d" % mode), frac=0.03, maxd=2)


def test_rgb_and_gray_rows():
    for name in ("RGB24", "RGB48", "GBRP10", "Y8", "V210"):
        ex = H.EX_709 if name == "V210" else 0
        _close(H.Case(O.CF[name], 66, 36, 99, 54, ex, up=O.UPSCALE["Lanczos2"], procamp=dict(brightness=4.3, contrast=1.07), name=name))
