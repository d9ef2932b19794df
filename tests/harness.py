"""Shared test harness: synthetic frames, one config description driving both the CPU
oracle (tests/orclib.py) and the product (videorenderer_b200 through the C ABI)."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import orclib as O  # noqa: E402


def fmt(cf):
    f = O.lib().orc_get_fmt(cf)
    return f.contents if f else None


# ------------------------------------------------------------------------------------------------
# synthetic content
# ------------------------------------------------------------------------------------------------
def _legal_ranges(bits):
    sh = bits - 8
    return (16 << sh, 235 << sh), (16 << sh, 240 << sh)


def make_codes(cf, w, h, kind="noise", seed=0):
    """Y/U/V code planes in the format's nominal bit depth (8, 10 or 16), before memory packing."""
    f = fmt(cf)
    name = f.name.decode()
    bits = 8 if f.bytes == 1 else (10 if (name in ("P010", "Y210", "Y410") or f.lsb_shift) else 16)
    if name in ("P016", "P216", "Y216", "Y416") or (f.bytes == 2 and not f.lsb_shift and name.endswith("16")):
        bits = 16
    if name in ("v210", "r210"):
        bits = 10
    cw, ch = w // f.div_w, h // f.div_h
    maxc = (1 << bits) - 1
    (ylo, yhi), (clo, chi) = _legal_ranges(bits)
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "ramp":
        Y = np.broadcast_to(np.round(np.linspace(0, maxc, w)).astype(np.int64)[None, :], (h, w)).copy()
        U = np.broadcast_to(np.round(np.linspace(0, maxc, ch)).astype(np.int64)[:, None], (ch, cw)).copy()
        V = np.broadcast_to(np.round(np.linspace(maxc, 0, ch)).astype(np.int64)[:, None], (ch, cw)).copy()
    elif kind == "noise":
        Y = rng.integers(ylo, yhi + 1, (h, w)); U = rng.integers(clo, chi + 1, (ch, cw)); V = rng.integers(clo, chi + 1, (ch, cw))
    elif kind == "fullnoise":
        Y = rng.integers(0, maxc + 1, (h, w)); U = rng.integers(0, maxc + 1, (ch, cw)); V = rng.integers(0, maxc + 1, (ch, cw))
    else:  # natural-ish: 5 seeded sinusoids + 5 % noise
        def field(hh, ww, lo, hi):
            yy, xx = np.mgrid[0:hh, 0:ww].astype(np.float64)
            acc = np.zeros((hh, ww))
            for _ in range(5):
                fx, fy = rng.uniform(0.5, 6, 2) * 2 * np.pi / max(ww, hh)
                acc += rng.uniform(0.3, 1) * np.sin(fx * xx + fy * yy + rng.uniform(0, 6.28))
            acc = acc / 5 * 0.5 + 0.5 + rng.normal(0, 0.05, (hh, ww))
            return np.clip(np.round(lo + acc * (hi - lo)), lo, hi).astype(np.int64)
        Y = field(h, w, ylo, yhi); U = field(ch, cw, clo, chi); V = field(ch, cw, clo, chi)
    return dict(Y=Y, U=U, V=V, bits=bits)


def _pack_rgb_sample(name, w, h, R, G, B, bottom_up):
    """RGB sample buffers as a DirectShow source hands them over (byte layouts per Source/Helper.cpp:345-354 comments and the
    FOURCC definitions); returns (buf, pitch>0)."""
    A8, A16 = 255, 65535
    if name == "RGB24":
        pitch = (w * 3 + 3) // 4 * 4
        buf = np.zeros((h, pitch), np.uint8)
        px = buf[:, :w * 3].reshape(h, w, 3); px[..., 0] = B; px[..., 1] = G; px[..., 2] = R
    elif name in ("RGB32", "ARGB32"):
        pitch = w * 4
        buf = np.zeros((h, pitch), np.uint8)
        px = buf.reshape(h, w, 4); px[..., 0] = B; px[..., 1] = G; px[..., 2] = R; px[..., 3] = A8 if name == "RGB32" else 0x80
    elif name == "r210":
        pitch = w * 4
        word = (R.astype(np.uint32) << 20) | (G.astype(np.uint32) << 10) | B.astype(np.uint32)
        buf = np.ascontiguousarray(word.astype(">u4")).view(np.uint8).reshape(h, pitch).copy()
    elif name in ("RGB48", "BGR48"):
        pitch = (w * 6 + 3) // 4 * 4 if name == "BGR48" else w * 6
        buf = np.zeros((h, pitch), np.uint8)
        px = buf[:, :w * 6].view("<u2").reshape(h, w, 3)
        first, last = (R, B) if name == "RGB48" else (B, R)
        px[..., 0] = first; px[..., 1] = G; px[..., 2] = last
    elif name == "BGRA64":
        pitch = w * 8
        buf = np.zeros((h, pitch), np.uint8)
        px = buf.view("<u2").reshape(h, w, 4); px[..., 0] = B; px[..., 1] = G; px[..., 2] = R; px[..., 3] = A16
    elif name == "b64a":
        pitch = w * 8
        buf = np.zeros((h, pitch), np.uint8)
        px = buf.view(">u2").reshape(h, w, 4); px[..., 0] = A16; px[..., 1] = R; px[..., 2] = G; px[..., 3] = B
    else:
        raise ValueError(name)
    if bottom_up:
        buf = buf[::-1].copy()
    return buf.reshape(-1), pitch


def _pack_v210(w, h, Y, U, V):
    """SMPTE v210: 6 pixels in 4 little-endian words, 3 x 10 bit each; rows padded to 128 bytes"""
    pitch = ((w + 5) // 6 * 16 + 127) // 128 * 128
    groups = (w + 5) // 6
    Yp = np.zeros((h, groups * 6), np.uint32); Yp[:, :w] = Y
    Up = np.zeros((h, groups * 3), np.uint32); Up[:, :w // 2] = U
    Vp = np.zeros((h, groups * 3), np.uint32); Vp[:, :w // 2] = V
    words = np.zeros((h, groups, 4), np.uint32)
    y = Yp.reshape(h, groups, 6); u = Up.reshape(h, groups, 3); v = Vp.reshape(h, groups, 3)
    words[..., 0] = u[..., 0] | (y[..., 0] << 10) | (v[..., 0] << 20)
    words[..., 1] = y[..., 1] | (u[..., 1] << 10) | (y[..., 2] << 20)
    words[..., 2] = v[..., 1] | (y[..., 3] << 10) | (u[..., 2] << 20)
    words[..., 3] = y[..., 4] | (v[..., 2] << 10) | (y[..., 5] << 20)
    buf = np.zeros((h, pitch), np.uint8)
    buf[:, :groups * 16] = words.astype("<u4").view(np.uint8).reshape(h, groups * 16)
    return buf.reshape(-1), pitch


def pack(cf, w, h, codes, pitch=None, bottom_up=False):
    """Pack code planes into the caller's frame buffer exactly as the reference expects it
    (layout walked by MemCopyToTexSrcVideo, Source/DX11VideoProcessor.cpp:1213-1252).
    Returns (buffer uint8[...], pitch, planes[list of np views], pitches)."""
    f = fmt(cf)
    name = f.name.decode()
    Y, U, V, bits = codes["Y"], codes["U"], codes["V"], codes["bits"]
    layout = f.layout
    cw, ch = w // f.div_w, h // f.div_h
    if layout == 5:      # RGB sample: the oracle sees the texture the reference's CPU repack produces (orc_repack)
        buf, spitch = _pack_rgb_sample(name, w, h, Y, U, V, bottom_up)      # code planes "Y,U,V" carry R,G,B here
        tex_pitch = w * (8 if f.packed_kind == 2 else 4)
        tex = O.repack(cf, w, h, buf, -spitch if bottom_up else spitch, tex_pitch)
        return buf, (-spitch if bottom_up else spitch), [tex.reshape(-1)], [tex_pitch]
    if name == "v210":
        buf, spitch = _pack_v210(w, h, Y, U, V)
        tex = O.repack(cf, w, h, buf, spitch, w * 4)
        return buf, spitch, [tex.reshape(-1)], [w * 4]
    if layout == 6:      # gray plane
        dtg = np.uint8 if f.bytes == 1 else np.uint16
        pitch = pitch or ((w + 3) // 4 * 4 if name == "Y8" else w * f.packsize)
        buf = np.zeros((h, pitch), np.uint8)
        buf[:, :w * f.bytes].view(dtg)[:] = Y.astype(dtg)
        return buf.reshape(-1), pitch, [buf.reshape(-1)], [pitch]
    if f.bytes == 1:
        dt = np.uint8; up = 0
    else:
        dt = np.uint16
        up = 6 if name in ("P010", "P210", "Y210") else 0          # MSB-aligned 10-bit containers
        if name in ("P210",):
            up = 6
    if layout == 1:      # biplanar
        base_pitch = w * f.packsize
        if name == "NV12":
            base_pitch = (base_pitch + 3) // 4 * 4
        pitch = pitch or base_pitch
        buf = np.zeros(pitch * (h + ch), np.uint8)
        yv = buf[:pitch * h].reshape(h, pitch)[:, :w * f.bytes].view(dt)
        yv[:] = (Y << up).astype(dt)
        cv = buf[pitch * h:].reshape(ch, pitch)[:, :cw * 2 * f.bytes].view(dt)
        cv[:, 0::2] = (U << up).astype(dt); cv[:, 1::2] = (V << up).astype(dt)
        planes = [buf[:pitch * h], buf[pitch * h:]]
        return buf, pitch, planes, [pitch, pitch]
    if layout == 2:      # planar3
        pitch = pitch or w * f.packsize
        cp = pitch // f.div_w
        buf = np.zeros(pitch * h + 2 * cp * ch, np.uint8)
        buf[:pitch * h].reshape(h, pitch)[:, :w * f.bytes].view(dt)[:] = Y.astype(dt)
        p1 = buf[pitch * h:pitch * h + cp * ch]; p2 = buf[pitch * h + cp * ch:]
        first, second = (V, U) if f.vu_swapped else (U, V)
        p1.reshape(ch, cp)[:, :cw * f.bytes].view(dt)[:] = first.astype(dt)
        p2.reshape(ch, cp)[:, :cw * f.bytes].view(dt)[:] = second.astype(dt)
        return buf, pitch, [buf[:pitch * h], p1, p2], [pitch, cp, cp]
    if layout == 3:      # packed 4:2:2
        pitch = pitch or w * f.packsize
        buf = np.zeros(pitch * h, np.uint8)
        row = buf.reshape(h, pitch)[:, :w * 2 * f.bytes].view(dt)     # [h, w*2] components
        if f.packed_kind == 0:
            row[:, 0::4] = (Y[:, 0::2] << up).astype(dt); row[:, 1::4] = (U << up).astype(dt)
            row[:, 2::4] = (Y[:, 1::2] << up).astype(dt); row[:, 3::4] = (V << up).astype(dt)
        else:
            row[:, 0::4] = U.astype(dt); row[:, 1::4] = Y[:, 0::2].astype(dt); row[:, 2::4] = V.astype(dt); row[:, 3::4] = Y[:, 1::2].astype(dt)
        return buf, pitch, [buf], [pitch]
    # packed 4:4:4
    pitch = pitch or w * f.packsize
    buf = np.zeros(pitch * h, np.uint8)
    if f.packed_kind == 0:       # AYUV bytes V U Y A
        r = buf.reshape(h, pitch)[:, :w * 4].reshape(h, w, 4)
        r[..., 0] = V; r[..., 1] = U; r[..., 2] = Y; r[..., 3] = 255
    elif f.packed_kind == 1:     # Y410
        r = buf.reshape(h, pitch)[:, :w * 4].view(np.uint32)
        r[:] = (U.astype(np.uint32) | (Y.astype(np.uint32) << 10) | (V.astype(np.uint32) << 20) | (3 << 30))
    else:                        # Y416
        r = buf.reshape(h, pitch)[:, :w * 8].view(np.uint16).reshape(h, w, 4)
        r[..., 0] = U; r[..., 1] = Y; r[..., 2] = V; r[..., 3] = 65535
    return buf, pitch, [buf], [pitch]


# ------------------------------------------------------------------------------------------------
# one config description for both implementations
# ------------------------------------------------------------------------------------------------
class Case:
    def __init__(self, cf, w, h, dst_w, dst_h, ex, backbuffer=O.SURF_BGRA8, tex=O.TEXFMT_AUTOINT, chroma=O.CHROMA_Bilinear,
                 up=O.UPSCALE["CatmullRom"], down=O.DOWNSCALE["Hamming"], interp50=1, dither=1, to_sdr=1, nits=125,
                 src_rect=None, dst_rect=None, win=None, variant=O.VARIANT_D3D11, procamp=None, interlaced=0, deint_blend=0,
                 name=None, bottom_up=0, hdr=None, rotation=0, flip=0, dovi=None, correction=0, upstream_vp=None):
        self.__dict__.update(locals()); del self.__dict__["self"]
        self.name = name or "case"

    def orc_config(self):
        pa = self.procamp or {}
        return O.make_config(self.cf, self.w, self.h, self.dst_w, self.dst_h, self.ex, backbuffer=self.backbuffer,
                             src_rect=self.src_rect, dst_rect=self.dst_rect, win=self.win, tex=self.tex, chroma=self.chroma,
                             up=self.up, down=self.down, interp50=self.interp50, dither=self.dither, to_sdr=self.to_sdr,
                             nits=self.nits, deint_blend=int(bool(self.deint_blend and self.interlaced)),
                             brightness=pa.get("brightness", 0.0), contrast=pa.get("contrast", 1.0), hue=pa.get("hue", 0.0),
                             saturation=pa.get("saturation", 1.0), variant=self.variant, hdr=self.hdr, rotation=self.rotation, flip=self.flip, correction=self.correction)

    def settings(self):
        import videorenderer_b200 as V
        s = V.Settings.default()
        s.bUseD3D11 = 1 if self.variant == O.VARIANT_D3D11 else 0
        s.iTexFormat = self.tex; s.iChromaScaling = self.chroma; s.iUpscaling = self.up; s.iDownscaling = self.down
        s.bInterpolateAt50pct = self.interp50; s.bUseDither = self.dither; s.bConvertToSdr = self.to_sdr
        s.iSDRDisplayNits = self.nits; s.bDeintBlend = self.deint_blend
        if self.hdr:
            self.hdr_settings(s)
        return s

    def make_processor(self, device=0, path=None):
        import videorenderer_b200 as V
        vp = V.VideoProcessor(self.settings(), device=device)
        if self.hdr:
            vp.SetDisplayHdr(True)
        if self.upstream_vp:                                   # (present, pq_supported): RGB samples from a fixed-function converter
            vp.SetUpstreamConverter(*self.upstream_vp)
        mt = V.VideoProcessor.media_type(self.cf, self.w, self.h, self.ex, src_rect=self.src_rect, interlaced=self.interlaced,
                                         bottom_up=self.bottom_up)
        assert vp.InitMediaType(mt), "InitMediaType failed: " + vp.last_error()
        if self.procamp:
            vp.SetProcAmpValues(**self.procamp)
        win = self.win or (self.dst_w, self.dst_h)
        vp.SetBackbufferFormat(self.backbuffer)
        vp.SetWindowRect((0, 0, win[0], win[1]))
        vp.SetVideoRect(self.dst_rect or (0, 0, self.dst_w, self.dst_h))
        if self.rotation or self.flip:
            assert vp.SetRotation(self.rotation) == 0 and vp.SetFlip(self.flip) == 0
        if path is not None:
            vp.debug_set_path(path)
        return vp

    def hdr_settings(self, s):
        """HDR output cases: Settings_t fields + the display capability (mpcvr_set_display_hdr)"""
        h = self.hdr or {}
        s.bHdrPassthrough = 1 if h.get("out", 0) else 0
        s.bHdrLocalToneMapping = 1 if h.get("tonemap", 0) else 0
        # "type" is what SetHDR10ShaderParams receives (oracle side); "settings_type" what the user selected, when they differ
        s.iHdrLocalToneMappingType = int(h.get("settings_type", h.get("type", 1))); s.iHdrDisplayMaxNits = int(h.get("display", 1000))
        return s


def run_oracle(case, planes, pitches, mode=O.MODE_SPEC, want_debug=False, threads=0):
    return O.process(case.orc_config(), planes, pitches, mode=mode, want_debug=want_debug, threads=threads, dovi=case.dovi)


def make_dovi(seed=0, mmr=0, bits=10, denom=23, l2=None, source_max_pq=3079, l1=None, l3=None):
    d = _make_dovi(seed, mmr, bits, denom)
    if l1:                                          # (min_pq, max_pq, avg_pq); l3: offsets around 2048
        d.ext[20].level = 1
        for k, v in enumerate(l1):
            d.ext[20].v[k] = int(v)
    if l3:
        d.ext[4].level = 3
        for k, v in enumerate(l3):
            d.ext[4].v[k] = int(v)
    d.source_min_pq = 62; d.source_max_pq = source_max_pq
    for i, blk in enumerate(l2 or []):              # (target_max_pq, slope, offset, power, chroma_weight, saturation_gain)
        e = d.ext[2 * i + 1]                         # leave gaps: the reference scans all 32 slots by level
        e.level = 2
        for k, v in enumerate(blk):
            e.v[k] = int(v)
    return d


def _make_dovi(seed=0, mmr=0, bits=10, denom=23):
    """Synthetic Dolby Vision reshaping data (OrcDovi; the product's mpcvr_dovi_metadata has the same layout): per component a
    few polynomial pieces close to identity, optionally MMR pieces of order 1..3 for the chroma components (as profile 5 streams
    have), the BT.2020-like YCC->RGB matrix with full-range offsets and an RGB->LMS matrix near the ICtCp one.
    mmr: 0 none, 1 one MMR piece per chroma curve, 2 several MMR pieces of different orders."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d = O.OrcDovi()
    d.coef_log2_denom = denom; d.bl_bit_depth = bits
    one = 1 << denom
    top = (1 << bits) - 1
    for c in range(3):
        cv = d.curves[c]
        if c == 0 or not mmr:
            npiv = int(rng.integers(2, 10))
            piv = np.sort(rng.choice(np.arange(1, top), npiv - 2, replace=False)) if npiv > 2 else np.array([], int)
            piv = np.concatenate([[0], piv, [top]])
            cv.num_pivots = npiv
            for i in range(npiv):
                cv.pivots[i] = int(piv[i])
            for i in range(npiv - 1):
                cv.mapping_idc[i] = 0
                cv.poly_order[i] = int(rng.integers(1, 3))
                cv.poly_coef[i][0] = int(rng.uniform(-0.02, 0.02) * one)
                cv.poly_coef[i][1] = int(rng.uniform(0.9, 1.1) * one)
                cv.poly_coef[i][2] = int(rng.uniform(-0.08, 0.08) * one)
        else:
            npiv = 2 if mmr == 1 else 4
            piv = [0, top] if mmr == 1 else [0, top // 3, 2 * top // 3, top]
            cv.num_pivots = npiv
            for i in range(npiv):
                cv.pivots[i] = piv[i]
            for i in range(npiv - 1):
                if mmr == 2 and i == 1 and c == 2:      # a polynomial piece between MMR pieces: methods == POLY + MMR
                    cv.mapping_idc[i] = 0; cv.poly_order[i] = 2
                    cv.poly_coef[i][0] = int(0.01 * one); cv.poly_coef[i][1] = int(0.97 * one); cv.poly_coef[i][2] = int(0.02 * one)
                    continue
                cv.mapping_idc[i] = 1
                order = 3 if mmr == 1 else int(1 + (i + c) % 3)
                cv.mmr_order[i] = order
                cv.mmr_constant[i] = int(rng.uniform(-0.02, 0.02) * one)
                for j in range(order):
                    for k in range(7):
                        base = 1.0 if (j == 0 and k == c) else 0.0       # first-order term of the own component ~ identity
                        cv.mmr_coef[i][j][k] = int((base + rng.uniform(-0.05, 0.05)) * one)
    ycc = np.array([[1.0, 0.0, 1.4746], [1.0, -0.16455, -0.57135], [1.0, 1.8814, 0.0]]) * rng.uniform(0.98, 1.02)
    for i in range(9):
        d.ycc_to_rgb_matrix[i] = float(ycc.ravel()[i])
    for i, v in enumerate((0.0, 0.5, 0.5)):
        d.ycc_to_rgb_offset[i] = v
    lms = np.array([[0.4121, 0.5239, 0.0640], [0.1667, 0.7204, 0.1129], [0.0242, 0.0755, 0.9004]]) + rng.uniform(-0.01, 0.01, (3, 3))
    for i in range(9):
        d.rgb_to_lms_matrix[i] = float(lms.ravel()[i])
    return d


def rgb_view(buf, fmt_):
    """decode a packed surface [h, w*bpp] uint8 into integer/float RGB [h, w, 3] (alpha dropped)"""
    h = buf.shape[0]
    if fmt_ == O.SURF_BGRA8:
        a = buf.reshape(h, -1, 4)
        return a[..., [2, 1, 0]].astype(np.int32)
    if fmt_ == O.SURF_RGB8:
        return buf.reshape(h, -1, 3).astype(np.int32)
    if fmt_ == O.SURF_RGB10A2:
        v = buf.view(np.uint32)
        return np.stack([v & 1023, (v >> 10) & 1023, (v >> 20) & 1023], -1).astype(np.int32)
    a = buf.view(np.float16).reshape(h, -1, 4)
    return a[..., :3]


def compare(a, b, fmt_):
    """returns (mismatching channel values, max abs code delta)"""
    ra, rb = rgb_view(a, fmt_), rgb_view(b, fmt_)
    if fmt_ == O.SURF_RGBA16F:
        ia = ra.view(np.int16).astype(np.int32); ib = rb.view(np.int16).astype(np.int32)
        d = np.abs(ia - ib)
        same = (ra == rb) | (d == 0)
        return int((~same).sum()), int(d.max())
    d = np.abs(ra - rb)
    return int((d != 0).sum()), int(d.max())


EX_709 = O.exfmt(chroma=5, range_=2, matrix=1, prim=2, trc=5)
EX_2020_PQ = O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=15)
EX_2020_HLG = O.exfmt(chroma=0, range_=2, matrix=4, prim=9, trc=16)
EX_2020_SDR = O.exfmt(chroma=5, range_=2, matrix=4, prim=9, trc=5)


def baseline_cases(scale=1.0):
    """The four single-GPU BASELINE.json configs, optionally scaled down (scale divides both source axes)."""
    def s(v, m=2):
        return max(m, int(round(v * scale / m)) * m)
    return [
        Case(O.CF["NV12"], s(1920), s(1080), s(1920), s(1080), EX_709, up=O.UPSCALE["Nearest"], name="C1_nv12_709_copy"),
        Case(O.CF["P010"], s(3840), s(2160), 2 * s(3840), 2 * s(2160), EX_2020_PQ, up=O.UPSCALE["Lanczos3"], name="C2_p010_pq_lanczos3_2x"),
        Case(O.CF["YUY2"], s(1920, 4), s(1080, 4), s(1920, 4) * 3 // 2, s(1080, 4) * 3 // 2, EX_2020_HLG, backbuffer=O.SURF_RGB10A2,
             tex=O.TEXFMT_16FLOAT, up=O.UPSCALE["CatmullRom"], name="C3_yuy2_hlg_catmull_1.5x"),
        Case(O.CF["P016"], s(7680, 4), s(4320, 4), s(7680, 4) // 2, s(4320, 4) // 2, EX_2020_SDR, backbuffer=O.SURF_RGBA16F,
             tex=O.TEXFMT_16FLOAT, down=O.DOWNSCALE["Bicubic"], interp50=0, dither=0, name="C4_p016_bicubic_0.5x"),
    ]
