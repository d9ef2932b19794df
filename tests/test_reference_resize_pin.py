"""Fifth pin against reference code EXECUTED here: the resize stage's draws.

`CDX11VideoProcessor::ResizeShaderPass` (Source/DX11VideoProcessor.cpp:3103-3187: which shader per direction, one pass or two through
m_TexResize, the quarter-turn rules), `TextureResizeShader` / `TextureCopyRect` (:300-377: viewport, the `scale` constants) and
`FillVertices` (:130-179: where the quad's corners sample the texture for each rotation / flip) are compiled where they lie against a
recording stand-in for the D3D11 device context (oracle/ref_stub/resize_prelude.h, resize_wrap.inc -> oracle/_ref/libref_resize.so;
nothing is copied).  For each configuration the reference's recorded draws are compared with the draws the ORACLE's resize stage
performs (orc_test_trace_*: records written by the oracle's pass functions from their own variables):

  * the number of draws, the pixel shader of each (upscale X / Y, downscale X / Y, the one-pass 2-D scaler, ps_simple), its source
    texture and render target (pass input, m_TexResize, target) and the source texture's size;
  * the viewport = the oracle's output rectangle;
  * the `scale` constant of convolution draws, bit for bit (TextureResizeShader divides same-named axes, turned or not);
  * the geometry: the texture coordinate the rasteriser interpolates from the reference's four vertices at the centres of the
    viewport's corner pixels, against the coordinate the oracle computed for those pixels (<= 2e-3 texel: both sides are fp32,
    in different operation orders).

The product's kernels are compared with the oracle pixel by pixel on the GPU (`test_rotation_*`, `test_jinc2*`, the scaler sweeps)."""
import ctypes as C
import os

import numpy as np
import pytest

import orclib as O

REF = os.path.join(O.ORC_DIR, "_ref", "libref_resize.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_resize.so not built (reference tree absent at build time)")


class TraceEntry(C.Structure):          # orc_trace_entry (oracle/orc_types.h)
    _fields_ = [("type", C.c_int32), ("in_id", C.c_int32), ("out_id", C.c_int32), ("in_w", C.c_int32), ("in_h", C.c_int32),
                ("in_rect", C.c_int32 * 4), ("out_rect", C.c_int32 * 4), ("rot", C.c_int32), ("flip", C.c_int32), ("in_axis", C.c_int32),
                ("kind", C.c_int32), ("filter", C.c_int32), ("u_from_x", C.c_int32), ("scale", C.c_float), ("cu", C.c_float * 2), ("cv", C.c_float * 2)]


def _libs():
    R = C.CDLL(REF)
    R.ref_resize_pass.restype = C.c_int
    R.ref_resize_pass.argtypes = [C.c_int, C.c_int, C.c_int * 4, C.c_int * 4, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int]
    L = O.lib()
    L.orc_test_trace_begin.argtypes = [C.POINTER(TraceEntry), C.c_int]
    L.orc_test_trace_end.restype = C.c_int
    return R, L


def _oracle_draws(L, cfg, sw, sh, resize_stage_only=True):
    conv = np.zeros((sh, sw * 4), np.uint8)                     # m_TexConvertOutput (RGB10A2), contents irrelevant here
    buf = (TraceEntry * 8)()
    L.orc_test_trace_begin(buf, 8)
    try:
        O.process(cfg, None, None, from_convert=conv, threads=1)
    finally:
        n = L.orc_test_trace_end()
    return [buf[i] for i in range(n) if buf[i].type <= 5 or not resize_stage_only]


def _ref_draws(R, sw, sh, dst, win, rot, flip, interp50, up_kind):
    out = (C.c_float * (38 * 8))()
    n = R.ref_resize_pass(sw, sh, (C.c_int * 4)(0, 0, sw, sh), (C.c_int * 4)(*dst), win[0], win[1], rot, flip, interp50, up_kind, out, 8)
    assert 0 <= n <= 8
    return np.array(out, np.float32).reshape(8, 38)[:n]


def _shader_id(e):
    if e.type == 2: return 5
    if e.type in (4, 5) or e.in_axis < 0: return 6
    return {(0, 1): 1, (1, 1): 2, (0, 2): 3, (1, 2): 4}[(e.in_axis, e.kind)]


def _ref_coord(d, x, y):
    """texture coordinate (texels) the rasteriser interpolates at the centre of viewport pixel (x, y) from the four recorded vertices"""
    W, H = float(d[8]), float(d[9]); sw, sh = float(d[2]), float(d[3])
    v = d[18:38].astype(np.float64).reshape(4, 5)
    X = -1.0 + 2.0 * (x + 0.5) / W; Y = 1.0 - 2.0 * (y + 0.5) / H
    A = np.array([[1.0, v[i][0], v[i][1]] for i in range(3)])
    cu = np.linalg.solve(A, v[:3, 3]); cv = np.linalg.solve(A, v[:3, 4])
    for c, col in ((cu, 3), (cv, 4)):                                # the fourth vertex lies on the same affine map (a parallelogram)
        assert abs(c[0] + c[1] * v[3][0] + c[2] * v[3][1] - v[3][col]) < 1e-6
    return (cu[0] + cu[1] * X + cu[2] * Y) * sw, (cv[0] + cv[1] * X + cv[2] * Y) * sh


UP = O.UPSCALE; DOWN = O.DOWNSCALE


def _cases():
    rng = np.random.Generator(np.random.PCG64(7))
    cases = []
    sizes = [(64, 48)]
    for rot in (0, 90, 180, 270):
        for flip in (0, 1):
            for up in ("CatmullRom", "Lanczos3", "Jinc2", "Nearest"):
                for (dw, dh) in ((96, 72), (64, 48), (48, 72), (40, 30), (64, 90), (20, 48), (96, 20), (31, 23), (130, 25), (48, 64), (200, 150)):
                    for interp50 in (1, 0):
                        if interp50 == 0 and (dw, dh) not in ((40, 30), (31, 23), (48, 72)): continue
                        ox, oy = int(rng.integers(0, 9)), int(rng.integers(0, 9))
                        cases.append(dict(sw=64, sh=48, dst=(ox, oy, ox + dw, oy + dh), win=(ox + dw + 5, oy + dh + 3), rot=rot, flip=flip, up=up,
                                          down="Hamming" if (dw + dh) % 2 else "Bicubic", interp50=interp50))
    return cases


def test_resize_stage_draws_equal_the_references():
    R, L = _libs()
    seen = {"one_pass_quirk": 0, "two_pass": 0, "jinc": 0, "copy": 0, "stretch": 0, "conv_scale": 0, "turned": 0}
    for cs in _cases():
        sw, sh = cs["sw"], cs["sh"]
        ex = O.exfmt(chroma=5, range_=2, matrix=1, prim=2, trc=5)
        cfg = O.make_config(O.CF["P010"], sw, sh, cs["dst"][2] - cs["dst"][0], cs["dst"][3] - cs["dst"][1], ex, dst_rect=cs["dst"], win=cs["win"],
                            up=UP[cs["up"]], down=DOWN[cs["down"]], interp50=cs["interp50"], rotation=cs["rot"], flip=cs["flip"])
        od = _oracle_draws(L, cfg, sw, sh)
        up_kind = 0 if cs["up"] == "Nearest" else (2 if cs["up"] == "Jinc2" else 1)
        rd = _ref_draws(R, sw, sh, cs["dst"], cs["win"], cs["rot"], cs["flip"], cs["interp50"], up_kind)
        # the oracle's "nothing to scale, upright" shortcut (the final pass reads the convert output itself, :3350-3352) is taken before
        # ResizeShaderPass would run in the reference as well; it only exists when a post-scale step follows, which is the case here
        same = (cs["dst"][2] - cs["dst"][0], cs["dst"][3] - cs["dst"][1]) == (sw, sh) and cs["dst"][0] == 0 and cs["dst"][1] == 0 and cs["rot"] == 0
        if same and not od:
            continue
        assert len(od) == len(rd), (cs, len(od), len(rd), [(_shader_id(e), e.type) for e in od], rd[:, 0])
        for e, d in zip(od, rd):
            assert _shader_id(e) == int(d[0]), (cs, _shader_id(e), int(d[0]))
            # ResizeShaderPass alone knows "the input" (0) and "the target" (2); the oracle names them as Process() does (see test_process_*)
            norm = {0: 0, 3: 0, 1: 1, 2: 2, 4: 2, 5: 2}
            assert (norm[e.in_id], norm[e.out_id]) == (int(d[1]), int(d[4])), (cs, e.in_id, e.out_id, d[1], d[4])
            assert (e.in_w, e.in_h) == (int(d[2]), int(d[3])), (cs, e.in_w, e.in_h, d[2], d[3])
            ow, oh = e.out_rect[2] - e.out_rect[0], e.out_rect[3] - e.out_rect[1]
            assert (e.out_rect[0], e.out_rect[1], ow, oh) == tuple(int(v) for v in d[6:10]), (cs, list(e.out_rect), d[6:10])
            if e.kind == 2 and e.type in (0, 1):
                ref_scale = d[10 + 4] if int(d[0]) == 3 else d[10 + 5]
                assert np.float32(e.scale).tobytes() == np.float32(ref_scale).tobytes(), (cs, e.scale, ref_scale)
                seen["conv_scale"] += 1
            for (x, y) in ((0, 0), (ow - 1, 0), (0, oh - 1), (ow - 1, oh - 1)):
                ui, vi = (x, y) if e.u_from_x else (y, x)
                ulen, vlen = (ow, oh) if e.u_from_x else (oh, ow)
                su = e.cu[0] if ui == 0 else e.cu[1]; sv = e.cv[0] if vi == 0 else e.cv[1]
                if ulen == 1: su = e.cu[1] if e.cu[0] == 0 and e.cu[1] != 0 else e.cu[0]
                ru, rv = _ref_coord(d, x, y)
                assert abs(su - ru) <= 2e-3 and abs(sv - rv) <= 2e-3, (cs, e.type, (x, y), (su, sv), (ru, rv))
        turned = cs["rot"] in (90, 270)
        seen["turned"] += int(turned)
        seen["two_pass"] += int(len(od) == 2)
        seen["jinc"] += int(any(e.type == 2 for e in od))
        seen["copy"] += int(any(e.type == 4 or (e.type == 1 and e.in_axis < 0) for e in od))
        seen["stretch"] += int(any(e.type == 5 for e in od))
        seen["one_pass_quirk"] += int(turned and len(od) == 1 and od[0].type == 1 and od[0].in_axis == 1 and
                                      (cs["dst"][2] - cs["dst"][0]) != sh and (cs["dst"][3] - cs["dst"][1]) != sw)
    assert all(v > 0 for v in seen.values()), seen


# ---- the whole of Process() (Source/DX11VideoProcessor.cpp:3285-3424): convert pass or not, the resize stage or the "Hmm" shortcut, the
# post-scale steps through the m_TexsPostScale ring, the last step into the back buffer ----
def _oracle_process_draws(L, cfg, planes, pitches):
    buf = (TraceEntry * 8)()
    L.orc_test_trace_begin(buf, 8)
    try:
        O.process(cfg, planes, pitches, threads=1)
    finally:
        n = L.orc_test_trace_end()
    return [buf[i] for i in range(n)]


def test_process_chain_equals_the_references():
    """every draw of a frame: CDX11VideoProcessor::Process, ResizeShaderPass, FinalPass, GetPostScaleSteps and the reference's own
    CTex2DRing compiled in place (ref_process) against the oracle's chain (orc_process with the draw trace): the same draws in the same
    order, each with the same shader, source texture, target texture and viewport -- including the shortcut that lets the first
    post-scale step read the convert output when nothing is scaled (and loses a flip on the way, :3350-3352), the ping-pong of the two
    post-scale textures, and RGB sources without a convert pass (the resize stage then samples m_TexSrcVideo inside m_srcRect)."""
    import harness as H
    R, L = _libs()
    R.ref_process.restype = C.c_int
    R.ref_process.argtypes = [C.c_int, C.c_int, C.c_int * 4, C.c_int, C.c_int * 4, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.POINTER(C.c_float), C.c_int]
    seen = {"hmm": 0, "lost_flip": 0, "no_convert": 0, "ring3": 0, "direct": 0, "two_pass": 0}
    rng = np.random.Generator(np.random.PCG64(11))
    W, Hh = 64, 48
    for fmt, convert in (("P010", 1), ("XRGB32", 0)):
        cf = O.CF[fmt]
        codes = H.make_codes(cf, W, Hh, "noise", 3)
        buf, pitch, planes, pitches = H.pack(cf, W, Hh, codes)
        for src in ((0, 0, W, Hh), (4, 2, 60, 46)):
            sw, sh = src[2] - src[0], src[3] - src[1]
            for (dw, dh) in ((sw, sh), (sw * 2, sh * 2), (sw, sh * 2), (sw // 2, sh // 2), (sw * 3 // 2, sh // 2)):
                for rot, flip in ((0, 0), (0, 1), (90, 0), (180, 1)):
                    for steps in ("none", "final", "tonemap+final", "correction+tonemap+final", "tonemap"):
                        at_origin = bool(rng.integers(0, 2)) or (dw, dh) == (sw, sh)
                        ox, oy = ((src[0], src[1]) if (dw, dh) == (sw, sh) and not convert else (0, 0)) if at_origin else (int(rng.integers(1, 9)), int(rng.integers(1, 9)))
                        dst = (ox, oy, ox + dw, oy + dh); win = (ox + dw + 3, oy + dh + 2) if not at_origin else (max(ox + dw, W), max(oy + dh, Hh))
                        final = "final" in steps; tone = "tonemap" in steps; corr = "correction" in steps
                        if corr and convert: continue                   # the correction shaders belong to RGB input from an upstream converter
                        ex = O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=15) if convert else 0
                        hdr = dict(out=1, tonemap=1, min=0.005, max=1000.0, cll=800.0, fall=200.0, display=600.0, type=2) if tone else None
                        # final pass <=> dither on and an internal format wider than the back buffer (UpdatePostScaleTexures); 10-bit source, 8-bit target
                        cfg = O.make_config(cf, W, Hh, dw, dh, ex, src_rect=src, dst_rect=dst, win=win, dither=1 if final else 0,
                                            tex=O.TEXFMT_16FLOAT if final or tone or corr else O.TEXFMT_8INT, to_sdr=0 if tone else 1,
                                            rotation=rot, flip=flip, hdr=hdr, correction=4 if corr else 0)
                        od = _oracle_process_draws(L, cfg, planes, pitches)
                        out = (C.c_float * (38 * 8))()
                        n = R.ref_process(W, Hh, (C.c_int * 4)(*src), convert, (C.c_int * 4)(*dst), win[0], win[1], rot, flip, 1, 1, int(corr), int(tone), int(final), out, 8)
                        rd = np.array(out, np.float32).reshape(8, 38)[:n]
                        tag = (fmt, src, dst, win, rot, flip, steps)
                        ops = [{6: 7, 7: 9, 8: 10, 9: 8}.get(e.type, _shader_id(e)) for e in od]
                        assert ops == [int(d[0]) for d in rd], (tag, ops, [int(d[0]) for d in rd])
                        for e, d in zip(od, rd):
                            assert (e.in_id, e.out_id) == (int(d[1]), int(d[4])), (tag, e.type, (e.in_id, e.out_id), (int(d[1]), int(d[4])))
                            assert (e.in_w, e.in_h) == (int(d[2]), int(d[3])), (tag, e.type)
                            ow, oh = e.out_rect[2] - e.out_rect[0], e.out_rect[3] - e.out_rect[1]
                            assert (e.out_rect[0], e.out_rect[1], ow, oh) == tuple(int(v) for v in d[6:10]), (tag, e.type, list(e.out_rect), d[6:10])
                        nres = sum(1 for e in od if e.type <= 5)
                        seen["hmm"] += int(nres == 0 and (final or tone or corr))
                        seen["lost_flip"] += int(nres == 0 and flip == 1)
                        seen["no_convert"] += int(not any(e.type == 6 for e in od))
                        seen["ring3"] += int(sum(1 for e in od if e.type >= 7) == 3)
                        seen["direct"] += int(steps == "none" and nres >= 1)
                        seen["two_pass"] += int(nres == 2)
    assert all(v > 0 for v in seen.values()), seen
