"""Sixth pin against reference code EXECUTED here: the per-stream planning decisions.

`CDX11VideoProcessor::UpdateTexParams` (Source/DX11VideoProcessor.cpp:1143-1155: internal texture format from iTexFormat and the source's
component depth), `UpdatePostScaleTexures` (:2894-2912: the final-pass decision m_bFinalPass, the 8 / 10-bit final-pass shader, the number of
post-scale textures), `GetPostScaleSteps` (:778-795) and the `convertType` expression of `UpdateConvertColorShader` (:2948-2950) are compiled
where they lie (oracle/ref_stub/plan_prelude.h, plan_wrap.inc -> oracle/_ref/libref_plan.so; nothing is copied) and compared with the
oracle's `orc_internal_fmt` / `orc_need_final_pass` and with the PRODUCT's plan on a planning-only instance (no GPU)."""
import ctypes as C
import os

import numpy as np
import pytest

import orclib as O
import harness as H

REF = os.path.join(O.ORC_DIR, "_ref", "libref_plan.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_plan.so not built (reference tree absent at build time)")

DXGI_TO_SURF = {87: O.SURF_BGRA8, 24: O.SURF_RGB10A2, 10: O.SURF_RGBA16F}
SURF_TO_DXGI = {v: k for k, v in DXGI_TO_SURF.items()}


def test_internal_format_and_final_pass_equal_the_references():
    R = C.CDLL(REF); L = O.lib()
    L.orc_internal_fmt.restype = C.c_int; L.orc_need_final_pass.restype = C.c_int
    n = 0
    for texfmt in (O.TEXFMT_AUTOINT, O.TEXFMT_8INT, O.TEXFMT_10INT, O.TEXFMT_16FLOAT):
        for fmt, cdepth in (("NV12", 8), ("P010", 10), ("P016", 16), ("YUV444P10", 10), ("XRGB32", 8), ("RGB48", 16)):
            for bb in (O.SURF_BGRA8, O.SURF_RGB10A2):
                for dither in (0, 1):
                    out = (C.c_int * 5)()
                    assert R.ref_plan(texfmt, cdepth, SURF_TO_DXGI[bb], dither, 1, 0, 0, 640, 360, out) == 0
                    ref_internal, ref_final, ref_resid, ref_ring, ref_ring_fmt = out[0], out[1], out[2], out[3], out[4]
                    # oracle
                    oi = L.orc_internal_fmt(texfmt, cdepth)
                    assert oi == DXGI_TO_SURF[ref_internal], (texfmt, cdepth, oi, ref_internal)
                    assert L.orc_need_final_pass(bb, oi, dither) == ref_final, (texfmt, cdepth, bb, dither)
                    # the final-pass shader variant follows the swap chain format, the ring holds one texture per post-scale step
                    if ref_final: assert ref_resid == (2 if bb == O.SURF_RGB10A2 else 1)
                    assert ref_ring == ref_final and (ref_ring == 0 or ref_ring_fmt == ref_internal)
                    # product (planning-only instance)
                    ex = H.EX_2020_PQ if fmt.startswith(("P0", "YUV")) and cdepth > 8 else (H.EX_709 if fmt == "NV12" else 0)
                    vp = H.Case(O.CF[fmt], 64, 36, 128, 72, ex, backbuffer=bb, tex=texfmt, dither=dither).make_processor(device=-1)
                    p = vp.debug_plan()
                    assert (p.internal_fmt, p.final_pass) == (DXGI_TO_SURF[ref_internal], ref_final), (fmt, texfmt, bb, dither, p.internal_fmt, p.final_pass)
                    n += 1
    assert n == 4 * 6 * 2 * 2
    # without the dither texture there is no final pass, whatever the formats; more post-scale steps -> a ring of two
    out = (C.c_int * 5)()
    R.ref_plan(O.TEXFMT_16FLOAT, 10, 87, 1, 0, 0, 0, 640, 360, out); assert out[1] == 0 and out[3] == 0
    R.ref_plan(O.TEXFMT_16FLOAT, 10, 87, 1, 1, 1, 1, 640, 360, out); assert out[1] == 1 and out[3] == 3     # CheckCreate is asked for 3, creates min(3, 2)


def test_convert_type_equals_the_references():
    """convertType (SHADER_CONVERT_NONE / TO_SDR / TO_PQ) over every combination of bConvertToSdr, HDR display support, bHdrPassthrough,
    bHdrLocalToneMapping and a PQ / HLG source, against the transfer the product's planner picks"""
    import videorenderer_b200 as V
    R = C.CDLL(REF)
    pq = V.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=15); hlg = V.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=16)
    expect = {(15, 0): 0, (15, 1): 1, (16, 0): 3, (16, 1): 2, (16, 2): 4}     # (source trc, convertType) -> the planner's transfer; an HLG stream left alone
    #                                                                          # takes the BT.2020 SDR branch of the shader (Source/Shaders.cpp:892-905)
    seen = set()
    for to_sdr in (0, 1):
        for support in (0, 1):
            for passthrough in (0, 1):
                for ltm in (0, 1):
                    for trc, ex in ((15, pq), (16, hlg)):
                        ct = R.ref_convert_type(to_sdr, support, passthrough, ltm, trc)
                        vp = V.VideoProcessor(device=-1)
                        s = vp.GetSettings(); s.bConvertToSdr = to_sdr; s.bHdrPassthrough = passthrough; s.bHdrLocalToneMapping = ltm; vp.Configure(s)
                        vp.SetDisplayHdr(bool(support))
                        assert vp.InitMediaType(V.VideoProcessor.media_type(V.CF.P010, 64, 64, ex))
                        t = vp.debug_plan().transfer
                        e = expect[(trc, ct)]
                        assert t == e, (to_sdr, support, passthrough, ltm, trc, ct, t)
                        seen.add((trc, ct))
    assert seen == {(15, 0), (15, 1), (16, 0), (16, 1), (16, 2)}
