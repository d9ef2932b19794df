"""Twelfth pin: the buffer geometry InitMediaType negotiates.

CDX11VideoProcessor::InitMediaType (Source/DX11VideoProcessor.cpp:1786-1819) derives, from the BITMAPINFOHEADER and the format table, how the
decoder's buffer is laid out: m_srcLines = |biHeight| * PitchCoeff / 2, m_srcPitch = biWidth * Packsize with the per-format alignment rules
(NV12 / Y8 / RGB24 / BGR48 to 4 bytes, v210 to 128 bytes of 6-pixel groups), negated for bottom-up DIBs; the true frame size out of the
VR_Extradata block CMpcVideoRenderer::CalcImageSize appends; the default source rectangle.  Those lines compile where they lie (with the
reference's own ALIGN macro, VR_Extradata and format table; oracle/ref_stub/mt_wrap_*.inc) into oracle/_ref/libref_shaders.so, and the
product's mpcvr_init_media_type (a planning-only instance, no GPU) must negotiate the same numbers for every format x a spread of sizes,
with and without extradata, top-down and bottom-up."""
import ctypes as C
import os

import numpy as np
import pytest

import orclib as O

REF = os.path.join(O.ORC_DIR, "_ref", "libref_shaders.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_shaders.so not built (reference tree absent at build time)")


def test_buffer_geometry_equals_the_references():
    import videorenderer_b200 as V
    R = C.CDLL(REF)
    R.ref_init_media_type_geometry.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int * 4, C.POINTER(C.c_int), C.c_int * 6]
    rng = np.random.Generator(np.random.PCG64(8))
    checked = 0
    for cf in range(1, 40):
        f = O.lib().orc_get_fmt(cf).contents
        rgb = f.cstype == 1 and f.layout == 5
        for trial in range(24):
            w = int(rng.choice([2, 6, 64, 66, 126, 720, 1278, 1280, 1920, 3840])) if trial < 16 else int(rng.integers(1, 2000)) * 2
            h = int(rng.choice([2, 36, 64, 480, 1080, 2160])) if trial < 16 else int(rng.integers(1, 1200)) * 2
            bottom_up = int(rgb and trial % 3 == 0)
            extra = None; frame = None
            if trial % 4 == 1:                                   # the buffer is wider / taller than the frame (pitch negotiation, Helper.h:57-65)
                fw, fh = max(2, w - 2 * int(rng.integers(0, 8))), max(2, h - 2 * int(rng.integers(0, 4)))
                frame = (fw, fh)
            rect = (0, 0, 0, 0)
            if trial % 5 == 2:
                rect = (2, 2, max(4, w - 2), max(4, h - 2))
            vp = V.VideoProcessor(V.Settings(), device=-1)
            mt = V.VideoProcessor.media_type(cf, w, h, src_rect=rect if any(rect) else None, frame_size=frame, bottom_up=bottom_up)
            ok = vp.InitMediaType(mt)
            # the reference side: biHeight carries the bottom-up sign for BI_RGB; the extradata block repeats the queried size
            biH = h if bottom_up else (-h if rgb else h)
            comp = 0 if rgb else 0x3231564E
            if frame:
                ex = (C.c_int * 5)(w, biH, frame[0], frame[1] if not bottom_up else frame[1], comp)
                extra = C.cast(ex, C.POINTER(C.c_int))
            out = (C.c_int * 6)()
            assert R.ref_init_media_type_geometry(cf, w, biH, comp, (C.c_int * 4)(*rect), extra, out) == 0
            if not ok:
                continue                                          # sizes the product refuses (odd chroma etc.) are not part of this comparison
            p = vp.debug_plan()
            assert (p.src_lines, p.src_pitch) == (out[0], out[1]), (f.name.decode(), w, h, bottom_up, frame, (p.src_lines, p.src_pitch), list(out))
            l, t, r, b = vp.GetSourceRect()
            assert (r - l, b - t) == (out[2], out[3]), (f.name.decode(), w, h, frame, rect, (l, t, r, b), list(out))
            checked += 1
    assert checked > 700, checked
