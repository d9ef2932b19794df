"""ctypes bindings of the CPU oracle (oracle/liborc.so) and, when built, the reference's own
csputils (oracle/_ref/libref_csputils.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_DIR = os.path.join(ROOT, "oracle")

# enums (oracle/orc_types.h == reference numeric values)
CF = dict(NONE=0, NV12=1, P010=2, P016=3, YUY2=4, UYVY=5, P210=6, P216=7, Y210=8, Y216=9, V210=10, AYUV=11,
          Y410=12, Y416=13, YV12=14, YV16=15, YV24=16, YUV420P8=17, YUV422P8=18, YUV444P8=19, YUV420P10=20,
          YUV420P16=21, YUV422P10=22, YUV422P16=23, YUV444P10=24, YUV444P16=25,
          GBRP8=26, GBRP10=27, GBRP16=28, RGB24=29, XRGB32=30, ARGB32=31, r210=32, RGB48=33, BGR48=34, BGRA64=35,
          B64A=36, Y8=37, Y10=38, Y16=39)
SURF_BGRA8, SURF_RGB10A2, SURF_RGBA16F = 0, 1, 2
SURF_RGB8 = 4        # product extension (packed R,G,B bytes of the B8G8R8A8 codes); the oracle renders B8G8R8A8 for it
MODE_SPEC, MODE_LITERAL, MODE_F64 = 0, 1, 2
VARIANT_D3D11, VARIANT_D3D9 = 0, 1
TEXFMT_AUTOINT, TEXFMT_8INT, TEXFMT_10INT, TEXFMT_16FLOAT = 0, 8, 10, 16
CHROMA_Nearest, CHROMA_Bilinear, CHROMA_CatmullRom = 0, 1, 2
UPSCALE = dict(Nearest=0, Mitchell=1, CatmullRom=2, Lanczos2=3, Lanczos3=4, Jinc2=5)
DOWNSCALE = dict(Box=0, Bilinear=1, Hamming=2, Bicubic=3, BicubicSharp=4, Lanczos=5)
CHROMA_SUB = dict(Unknown=0, MPEG1=1, MPEG2=5, DV_PAL=6, Cosited=7)
RANGE = dict(Unknown=0, R0_255=1, R16_235=2)
MATRIX = dict(Unknown=0, BT709=1, BT601=2, SMPTE240M=3, BT2020=4, YCgCo=7)
PRIM = dict(Unknown=0, BT709=2, BT470M=3, BT470BG=4, SMPTE170M=5, SMPTE240M=6, BT2020=9, DCI_P3=11)
TRC = dict(Unknown=0, L10=1, G18=2, G20=3, G22=4, BT709=5, S240M=6, sRGB=7, G28=8, G26=14, PQ=15, HLG=16)


def exfmt(chroma=0, range_=0, matrix=0, light=0, prim=0, trc=0, sample=0):
    """DXVA2_ExtendedFormat.value"""
    return (sample | (chroma << 8) | (range_ << 12) | (matrix << 15) | (light << 18) | (prim << 22) | (trc << 27)) & 0xFFFFFFFF


class OrcConfig(C.Structure):
    _fields_ = [("cformat", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("src_rect", C.c_int32 * 4), ("exfmt", C.c_uint32), ("dst_rect", C.c_int32 * 4),
                ("win_w", C.c_int32), ("win_h", C.c_int32), ("backbuffer_fmt", C.c_int32),
                ("iTexFormat", C.c_int32), ("iChromaScaling", C.c_int32), ("iUpscaling", C.c_int32),
                ("iDownscaling", C.c_int32), ("bInterpolateAt50pct", C.c_int32), ("bUseDither", C.c_int32),
                ("bConvertToSdr", C.c_int32), ("iSDRDisplayNits", C.c_int32), ("bDeintBlend", C.c_int32),
                ("brightness", C.c_float), ("contrast", C.c_float), ("hue", C.c_float), ("saturation", C.c_float),
                ("variant", C.c_int32),
                ("hdr_out", C.c_int32), ("tonemap", C.c_int32),
                ("tm_min_nits", C.c_float), ("tm_max_nits", C.c_float), ("tm_max_cll", C.c_float), ("tm_max_fall", C.c_float),
                ("tm_display_nits", C.c_float), ("tm_type", C.c_int32), ("rotation", C.c_int32), ("flip", C.c_int32),
                ("correction", C.c_int32)]


class OrcDebug(C.Structure):
    _fields_ = [("convert_f32", C.c_void_p), ("convert_surf", C.c_void_p), ("convert_surf_pitch", C.c_int32),
                ("resize_x_f16", C.c_void_p), ("postscale_surf", C.c_void_p), ("postscale_pitch", C.c_int32),
                ("out_internal_fmt", C.c_int32), ("out_final_pass", C.c_int32),
                ("out_resize_x", C.c_int32), ("out_resize_y", C.c_int32)]


class OrcDoviCurve(C.Structure):      # orc_dovi_curve
    _fields_ = [("num_pivots", C.c_uint8), ("pivots", C.c_uint16 * 9), ("mapping_idc", C.c_uint8 * 8), ("poly_order", C.c_uint8 * 8),
                ("poly_coef", (C.c_int64 * 3) * 8), ("mmr_order", C.c_uint8 * 8), ("mmr_constant", C.c_int64 * 8),
                ("mmr_coef", ((C.c_int64 * 7) * 3) * 8)]


class OrcDoviExt(C.Structure):
    _fields_ = [("level", C.c_uint8), ("v", C.c_uint16 * 7)]


class OrcDovi(C.Structure):           # orc_dovi
    _fields_ = [("coef_log2_denom", C.c_uint8), ("bl_bit_depth", C.c_uint8), ("curves", OrcDoviCurve * 3),
                ("ycc_to_rgb_matrix", C.c_double * 9), ("ycc_to_rgb_offset", C.c_double * 3), ("rgb_to_lms_matrix", C.c_double * 9),
                ("source_min_pq", C.c_uint16), ("source_max_pq", C.c_uint16), ("ext", OrcDoviExt * 32)]


class OrcFmt(C.Structure):
    _fields_ = [("cformat", C.c_int32), ("layout", C.c_int32), ("bytes", C.c_int32), ("cdepth", C.c_int32),
                ("cstype", C.c_int32), ("subsampling", C.c_int32), ("div_w", C.c_int32), ("div_h", C.c_int32),
                ("lsb_shift", C.c_int32), ("packsize", C.c_int32), ("pitchcoeff", C.c_int32),
                ("vu_swapped", C.c_int32), ("packed_kind", C.c_int32), ("name", C.c_char_p)]


_lib = None
_ref = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORC_DIR], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORC_DIR, "liborc.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_process.restype = C.c_int
        L.orc_process.argtypes = [C.POINTER(OrcConfig), C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_void_p,
                                  C.c_int32, C.POINTER(OrcDebug), C.c_int, C.c_int]
        L.orc_process_dovi.restype = C.c_int
        L.orc_process_dovi.argtypes = [C.POINTER(OrcConfig), C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_void_p,
                                       C.c_int32, C.POINTER(OrcDebug), C.c_int, C.c_int, C.c_void_p]
        L.orc_dovi_l1_nits.restype = C.c_int
        L.orc_dovi_l1_nits.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_process_from_convert.restype = C.c_int
        L.orc_process_from_convert.argtypes = [C.POINTER(OrcConfig), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                               C.POINTER(OrcDebug), C.c_int, C.c_int]
        L.orc_csp_matrix.restype = C.c_int
        L.orc_csp_matrix.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_void_p, C.c_void_p]
        L.orc_gamut_matrix.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_gamut_matrix_2020_to_709.argtypes = [C.c_void_p]
        L.orc_set_colorspace.argtypes = [C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_specify_extended_format.restype = C.c_uint32
        L.orc_specify_extended_format.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_repack.restype = C.c_int
        L.orc_repack.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_get_fmt.restype = C.POINTER(OrcFmt)
        L.orc_get_fmt.argtypes = [C.c_int]
        L.orc_dither_thresholds.restype = C.POINTER(C.c_uint16)
        L.orc_test_lg2.restype = C.c_float; L.orc_test_lg2.argtypes = [C.c_float]
        L.orc_test_sin.restype = C.c_float; L.orc_test_sin.argtypes = [C.c_float]
        L.orc_test_ex2.restype = C.c_float; L.orc_test_ex2.argtypes = [C.c_float]
        L.orc_test_pow.restype = C.c_float; L.orc_test_pow.argtypes = [C.c_float, C.c_float, C.c_int]
        L.orc_axis_taps.restype = C.c_int
        L.orc_axis_taps.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_float)]
        L.orc_internal_fmt.argtypes = [C.c_int, C.c_int]
        L.orc_need_final_pass.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_scaler_kind.argtypes = [C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


def ref():
    """the reference's own csputils compiled in place (None when not built / not shipped)"""
    global _ref
    if _ref is None:
        path = os.path.join(ORC_DIR, "_ref", "libref_csputils.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_csp_matrix.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p]
        R.ref_gamut_matrix.argtypes = [C.c_int, C.c_int, C.c_void_p]
        _ref = R
    return _ref


_refcopy = None


def ref_copy():
    """the reference's own CopyFrame* functions compiled in place (None when not built / not shipped)"""
    global _refcopy
    if _refcopy is None:
        path = os.path.join(ORC_DIR, "_ref", "libref_copy.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_copy_frame.restype = C.c_int
        R.ref_copy_frame.argtypes = [C.c_int, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_int]
        _refcopy = R
    return _refcopy


def repack(cformat, width, lines, src, src_pitch, dst_pitch):
    """orc_repack: sample buffer -> source texture memory; src_pitch < 0 = bottom-up DIB (src = start of the buffer)"""
    dst = np.zeros((lines, dst_pitch), np.uint8)
    rc = lib().orc_repack(cformat, width, lines, src.ctypes.data, src_pitch, dst.ctypes.data, dst_pitch)
    assert rc == 0, rc
    return dst


def csp_matrix(ex, cdepth, gray=0, brightness=0.0, contrast=1.0, hue=0.0, saturation=1.0):
    m = np.zeros((3, 3), np.float32); c = np.zeros(3, np.float32)
    rc = lib().orc_csp_matrix(ex, cdepth, gray, brightness, contrast, hue, saturation, m.ctypes.data, c.ctypes.data)
    assert rc == 0
    return m, c


def surf_bpp(fmt):
    return 8 if fmt == SURF_RGBA16F else 4


def make_config(cformat, width, height, dst_w, dst_h, ex, backbuffer=SURF_BGRA8, src_rect=None, dst_rect=None,
                win=None, tex=TEXFMT_AUTOINT, chroma=CHROMA_Bilinear, up=UPSCALE["CatmullRom"],
                down=DOWNSCALE["Hamming"], interp50=1, dither=1, to_sdr=1, nits=125, deint_blend=0,
                brightness=0.0, contrast=1.0, hue=0.0, saturation=1.0, variant=VARIANT_D3D11, hdr=None,
                rotation=0, flip=0, correction=0):
    cfg = OrcConfig()
    cfg.correction = correction
    cfg.rotation = rotation; cfg.flip = flip
    if hdr:   # dict(out=1, tonemap=0/1, min, max, cll, fall, display, type): SetHDR10ShaderParams' raw arguments
        cfg.hdr_out = int(hdr.get("out", 1)); cfg.tonemap = int(hdr.get("tonemap", 0))
        cfg.tm_min_nits = hdr.get("min", 0.0); cfg.tm_max_nits = hdr.get("max", 0.0); cfg.tm_max_cll = hdr.get("cll", 0.0)
        cfg.tm_max_fall = hdr.get("fall", 0.0); cfg.tm_display_nits = hdr.get("display", 1000.0); cfg.tm_type = int(hdr.get("type", 1))
    cfg.cformat = cformat; cfg.width = width; cfg.height = height
    sr = src_rect or (0, 0, width, height)
    dr = dst_rect or (0, 0, dst_w, dst_h)
    for i in range(4):
        cfg.src_rect[i] = sr[i]; cfg.dst_rect[i] = dr[i]
    cfg.exfmt = ex
    cfg.win_w, cfg.win_h = win or (dst_w, dst_h)
    cfg.backbuffer_fmt = backbuffer
    cfg.iTexFormat = tex; cfg.iChromaScaling = chroma; cfg.iUpscaling = up; cfg.iDownscaling = down
    cfg.bInterpolateAt50pct = interp50; cfg.bUseDither = dither; cfg.bConvertToSdr = to_sdr
    cfg.iSDRDisplayNits = nits; cfg.bDeintBlend = deint_blend
    cfg.brightness = brightness; cfg.contrast = contrast; cfg.hue = hue; cfg.saturation = saturation
    cfg.variant = variant
    return cfg


def process(cfg, planes, pitches, mode=MODE_SPEC, threads=0, want_debug=False, from_convert=None, dovi=None):
    """planes: list of numpy arrays (kept alive by caller). Returns out bytes array [win_h, win_w*bpp] (+debug dict)"""
    L = lib()
    bpp = surf_bpp(cfg.backbuffer_fmt)
    out = np.zeros((cfg.win_h, cfg.win_w * bpp), np.uint8)
    dbg = OrcDebug()
    keep = {}
    sw = cfg.src_rect[2] - cfg.src_rect[0]; sh = cfg.src_rect[3] - cfg.src_rect[1]
    if want_debug:
        keep["convert_f32"] = np.zeros((sh, sw, 3), np.float32)
        keep["convert_surf"] = np.zeros((sh, sw * 8), np.uint8)
        keep["resize_x_f16"] = np.zeros((sh, cfg.win_w, 4), np.float16)
        keep["postscale_surf"] = np.zeros((cfg.win_h, cfg.win_w * 8), np.uint8)
        dbg.convert_f32 = keep["convert_f32"].ctypes.data
        dbg.convert_surf = keep["convert_surf"].ctypes.data; dbg.convert_surf_pitch = sw * 8
        dbg.resize_x_f16 = keep["resize_x_f16"].ctypes.data
        dbg.postscale_surf = keep["postscale_surf"].ctypes.data; dbg.postscale_pitch = cfg.win_w * 8
    if from_convert is not None:
        rc = L.orc_process_from_convert(C.byref(cfg), from_convert.ctypes.data, from_convert.strides[0],
                                        out.ctypes.data, out.strides[0], C.byref(dbg), mode, threads)
    else:
        pl = (C.c_void_p * 3)(*[p.ctypes.data if p is not None else None for p in (list(planes) + [None] * 3)[:3]])
        pi = (C.c_int32 * 3)(*((list(pitches) + [0] * 3)[:3]))
        rc = L.orc_process_dovi(C.byref(cfg), pl, pi, out.ctypes.data, out.strides[0], C.byref(dbg), mode, threads,
                                C.addressof(dovi) if dovi is not None else None)
    if rc != 0:
        raise RuntimeError("orc_process failed rc=%d" % rc)
    if want_debug:
        keep["internal_fmt"] = dbg.out_internal_fmt; keep["final_pass"] = dbg.out_final_pass
        keep["resize_x"] = dbg.out_resize_x; keep["resize_y"] = dbg.out_resize_y
        return out, keep
    return out


def axis_taps(kind, filt, variant, origin, src_len, dst_len, i):
    base = C.c_int(); wsum = C.c_float(); w = np.zeros(1024, np.float32)
    n = lib().orc_axis_taps(kind, filt, variant, origin, src_len, dst_len, i, C.byref(base), w.ctypes.data, C.byref(wsum))
    return base.value, w[:max(n, 0)].copy(), wsum.value, n
