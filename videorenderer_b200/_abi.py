"""ctypes mirror of include/mpcvr_b200.h and include/mpcvr_b200_debug.h."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    # MPCVR_B200_LIB: developer override used by tools/sweep_threads.sh to compare kernel variants; unset in normal use
    return os.environ.get("MPCVR_B200_LIB") or os.path.join(_HERE, "libmpcvr_b200.so")


class MpcvrError(RuntimeError):
    def __init__(self, hr, msg):
        super().__init__("HRESULT 0x%08X: %s" % (hr & 0xFFFFFFFF, msg))
        self.hr = hr


class HRESULT:
    S_OK = 0
    S_FALSE = 1
    E_NOTIMPL = C.c_int32(0x80004001).value
    E_POINTER = C.c_int32(0x80004003).value
    E_ABORT = C.c_int32(0x80004004).value
    E_FAIL = C.c_int32(0x80004005).value
    E_UNEXPECTED = C.c_int32(0x8000FFFF).value
    E_OUTOFMEMORY = C.c_int32(0x8007000E).value
    E_INVALIDARG = C.c_int32(0x80070057).value


class CF:  # ColorFormat_t, Source/Helper.h:86-127
    NONE, NV12, P010, P016, YUY2, UYVY, P210, P216, Y210, Y216, V210, AYUV, Y410, Y416, YV12, YV16, YV24, \
        YUV420P8, YUV422P8, YUV444P8, YUV420P10, YUV420P16, YUV422P10, YUV422P16, YUV444P10, YUV444P16, \
        GBRP8, GBRP10, GBRP16, RGB24, XRGB32, ARGB32, r210, RGB48, BGR48, BGRA64, B64A, Y8, Y10, Y16 = range(40)


class SURF:
    BGRA8, RGB10A2, RGBA16F = 0, 1, 2
    RGB8 = 4              # extension: packed R,G,B bytes


class TEXFMT:
    AUTOINT, INT8, INT10, FLOAT16 = 0, 8, 10, 16


class CHROMA:
    Nearest, Bilinear, CatmullRom = 0, 1, 2


class UPSCALE:
    Nearest, Mitchell, CatmullRom, Lanczos2, Lanczos3, Jinc2 = range(6)


class DOWNSCALE:
    Box, Bilinear, Hamming, Bicubic, BicubicSharp, Lanczos = range(6)


class PATH:
    AUTO, MULTIPASS, FUSED = 0, 1, 2


def exfmt(chroma=0, range_=0, matrix=0, light=0, prim=0, trc=0, sample=0):
    """DXVA2_ExtendedFormat.value from its bit fields."""
    return (sample | (chroma << 8) | (range_ << 12) | (matrix << 15) | (light << 18) | (prim << 22) | (trc << 27)) & 0xFFFFFFFF


class Settings(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "bUseD3D11", "bShowStats", "iResizeStats", "iTexFormat", "VPFmts_bNV12", "VPFmts_bP01x", "VPFmts_bYUY2",
        "VPFmts_bOther", "iVPDeinterlacing", "bDeintDouble", "bVPScaling", "iVPSuperRes", "bVPRTXVideoHDR",
        "iChromaScaling", "iUpscaling", "iDownscaling", "bInterpolateAt50pct", "bUseDither", "bDeintBlend",
        "iSwapEffect", "bExclusiveFS", "bVBlankBeforePresent", "bAdjustPresentTime", "bReinitByDisplay",
        "bHdrPreferDoVi", "bHdrPassthrough", "iHdrToggleDisplay", "iHdrOsdBrightness", "bConvertToSdr",
        "iSDRDisplayNits", "bHdrLocalToneMapping", "iHdrLocalToneMappingType", "iHdrDisplayMaxNits")]

    @classmethod
    def default(cls):
        s = cls()
        lib().mpcvr_settings_default(C.byref(s))
        return s


class Rect(C.Structure):
    _fields_ = [("left", C.c_int32), ("top", C.c_int32), ("right", C.c_int32), ("bottom", C.c_int32)]

    def tuple(self):
        return (self.left, self.top, self.right, self.bottom)


class MediaType(C.Structure):
    _fields_ = [("cformat", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("frame_width", C.c_int32),
                ("frame_height", C.c_int32), ("src_rect", Rect), ("aspect_x", C.c_uint32), ("aspect_y", C.c_uint32),
                ("ext_format", C.c_uint32), ("interlaced", C.c_int32), ("avg_time_per_frame", C.c_int64),
                ("bottom_up", C.c_int32), ("reserved", C.c_int32)]


class SideDataHDR(C.Structure):
    _pack_ = 1
    _fields_ = [("display_primaries_x", C.c_double * 3), ("display_primaries_y", C.c_double * 3),
                ("white_point_x", C.c_double), ("white_point_y", C.c_double),
                ("max_display_mastering_luminance", C.c_double), ("min_display_mastering_luminance", C.c_double)]


class SideDataHDRCLL(C.Structure):
    _pack_ = 1
    _fields_ = [("MaxCLL", C.c_uint32), ("MaxFALL", C.c_uint32)]


class DoviCurve(C.Structure):         # mpcvr_dovi_curve (natural C layout)
    _fields_ = [("num_pivots", C.c_uint8), ("pivots", C.c_uint16 * 9), ("mapping_idc", C.c_uint8 * 8), ("poly_order", C.c_uint8 * 8),
                ("poly_coef", (C.c_int64 * 3) * 8), ("mmr_order", C.c_uint8 * 8), ("mmr_constant", C.c_int64 * 8),
                ("mmr_coef", ((C.c_int64 * 7) * 3) * 8)]


class DoviExt(C.Structure):
    _fields_ = [("level", C.c_uint8), ("v", C.c_uint16 * 7)]


class DoviNlq(C.Structure):
    _fields_ = [("nlq_offset", C.c_uint16), ("vdr_in_max", C.c_uint64), ("linear_deadzone_slope", C.c_uint64), ("linear_deadzone_threshold", C.c_uint64)]


class DoviMetadata(C.Structure):      # mpcvr_dovi_metadata
    _fields_ = [("coef_log2_denom", C.c_uint8), ("bl_bit_depth", C.c_uint8), ("curves", DoviCurve * 3),
                ("ycc_to_rgb_matrix", C.c_double * 9), ("ycc_to_rgb_offset", C.c_double * 3), ("rgb_to_lms_matrix", C.c_double * 9),
                ("source_min_pq", C.c_uint16), ("source_max_pq", C.c_uint16), ("ext", DoviExt * 32),
                ("el_spatial_resampling_filter_flag", C.c_uint8), ("disable_residual_flag", C.c_uint8), ("vdr_bit_depth", C.c_uint8),
                ("nlq_method_idc", C.c_uint8), ("nlq", DoviNlq * 3)]


SAMPLE_DEVICE_MEMORY = 0x1


class Sample(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("data", C.c_void_p), ("size", C.c_int64), ("pitch", C.c_int32),
                ("planes", C.c_void_p * 3), ("pitches", C.c_int32 * 3), ("ready_event", C.c_void_p),
                ("rt_start", C.c_int64), ("rt_end", C.c_int64), ("hdr", C.POINTER(SideDataHDR)),
                ("hdr_cll", C.POINTER(SideDataHDRCLL)), ("dovi", C.POINTER(DoviMetadata))]


class ProcAmp(C.Structure):
    _fields_ = [("brightness", C.c_float), ("contrast", C.c_float), ("hue", C.c_float), ("saturation", C.c_float)]


class PlanInfo(C.Structure):
    _fields_ = [("internal_fmt", C.c_int32), ("final_pass", C.c_int32), ("scaler_x", C.c_int32), ("scaler_y", C.c_int32),
                ("filter_x", C.c_int32), ("filter_y", C.c_int32), ("taps_x", C.c_int32), ("taps_y", C.c_int32),
                ("path", C.c_int32), ("transfer", C.c_int32), ("src_exfmt", C.c_uint32), ("src_pitch", C.c_int32),
                ("src_lines", C.c_int32), ("variant", C.c_int32), ("correction", C.c_int32)]


SUB_RENDER_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, Rect, Rect, Rect, C.c_double,
                            C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32)
SUB_SET_DEVICE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32)

# name -> (restype, argtypes); every symbol include/mpcvr_b200.h and include/mpcvr_b200_debug.h declare
P = C.POINTER
vp = C.c_void_p
SIGNATURES = {
    "mpcvr_settings_default": (None, [P(Settings)]),
    "mpcvr_create": (C.c_int32, [P(Settings), C.c_int32, vp, P(vp)]),
    "mpcvr_destroy": (None, [vp]),
    "mpcvr_abi_version": (C.c_int32, []),
    "mpcvr_last_error": (C.c_char_p, [vp]),
    "mpcvr_verify_media_type": (C.c_int32, [vp, P(MediaType)]),
    "mpcvr_init_media_type": (C.c_int32, [vp, P(MediaType)]),
    "mpcvr_get_alignment_size": (C.c_int32, [vp, P(MediaType), P(C.c_int32), P(C.c_int32)]),
    "mpcvr_process_sample": (C.c_int32, [vp, P(Sample)]),
    "mpcvr_render": (C.c_int32, [vp, C.c_int32, C.c_int64]),
    "mpcvr_fill_black": (C.c_int32, [vp]),
    "mpcvr_flush": (None, [vp]),
    "mpcvr_reset": (C.c_int32, [vp, C.c_int32]),
    "mpcvr_wait_idle": (C.c_int32, [vp]),
    "mpcvr_get_source_rect": (None, [vp, P(Rect)]),
    "mpcvr_get_video_rect": (None, [vp, P(Rect)]),
    "mpcvr_set_video_rect": (None, [vp, P(Rect)]),
    "mpcvr_set_window_rect": (C.c_int32, [vp, P(Rect)]),
    "mpcvr_set_rotation": (C.c_int32, [vp, C.c_int32]),
    "mpcvr_set_flip": (C.c_int32, [vp, C.c_int32]),
    "mpcvr_set_display_hdr": (C.c_int32, [vp, C.c_int32]),
    "mpcvr_set_upstream_converter": (C.c_int32, [vp, C.c_int32, C.c_int32]),
    "mpcvr_get_video_size": (None, [vp, P(C.c_int32), P(C.c_int32)]),
    "mpcvr_get_aspect_ratio": (None, [vp, P(C.c_int32), P(C.c_int32)]),
    "mpcvr_get_settings": (None, [vp, P(Settings)]),
    "mpcvr_configure": (None, [vp, P(Settings)]),
    "mpcvr_get_active": (C.c_int32, [vp]),
    "mpcvr_set_procamp": (C.c_int32, [vp, C.c_uint32, P(ProcAmp)]),
    "mpcvr_get_procamp": (C.c_int32, [vp, C.c_uint32, P(ProcAmp)]),
    "mpcvr_set_backbuffer_format": (C.c_int32, [vp, C.c_int32]),
    "mpcvr_bind_backbuffer": (C.c_int32, [vp, vp, C.c_size_t]),
    "mpcvr_get_backbuffer": (C.c_int32, [vp, P(vp), P(C.c_size_t), P(C.c_int32), P(C.c_int32), P(C.c_int32)]),
    "mpcvr_read_backbuffer": (C.c_int32, [vp, vp, C.c_size_t]),
    "mpcvr_read_backbuffer_async": (C.c_int32, [vp, vp, C.c_size_t]),
    "mpcvr_get_current_image": (C.c_int32, [vp, vp, C.c_size_t, P(C.c_int32), P(C.c_int32)]),
    "mpcvr_get_displayed_image": (C.c_int32, [vp, vp, C.c_size_t, P(C.c_int32), P(C.c_int32)]),
    "mpcvr_get_vp_info": (C.c_int32, [vp, C.c_char_p, C.c_size_t]),
    "mpcvr_set_sub_callback": (C.c_int32, [vp, SUB_RENDER_FN, SUB_SET_DEVICE_FN, vp]),
    "mpcvr_activate_device_decoding": (C.c_int32, [vp, C.c_int32, vp, C.c_uint32]),
    "mpcvr_get_device_index": (C.c_int32, [vp]),
    "mpcvr_get_stream": (C.c_int32, [vp, P(vp)]),
    "mpcvr_get_last_frame_kernels": (C.c_int32, [vp, P(C.c_int32)]),
    # debug header
    "mpcvr_debug_set_path": (C.c_int32, [vp, C.c_int32]),
    "mpcvr_debug_set_gamma_table": (C.c_int32, [C.c_int32]),
    "mpcvr_debug_render_n": (C.c_int32, [C.c_void_p, C.c_int32]),
    "mpcvr_debug_set_fused1x_vector": (C.c_int32, [C.c_int32]),
    "mpcvr_debug_fmt_params": (C.c_int32, [C.c_int32, C.POINTER(C.c_int32)]),
    "mpcvr_debug_specify_extended_format": (C.c_uint32, [C.c_uint32, C.c_int32, C.c_uint32, C.c_uint32]),
    "mpcvr_debug_set_strip_kernel": (C.c_int32, [C.c_int32, C.c_int32]),
    "mpcvr_debug_strip_status": (C.c_int32, [C.POINTER(C.c_uint32), C.c_int32]),
    "mpcvr_debug_set_convert_two_lane": (C.c_int32, [C.c_int32]),
    "mpcvr_debug_set_resize_four": (C.c_int32, [C.c_int32]),
    "mpcvr_debug_verify_gamma_table": (C.c_int64, [vp]),
    "mpcvr_debug_get_plan": (C.c_int32, [vp, P(PlanInfo)]),
    "mpcvr_debug_get_color_params": (C.c_int32, [vp, vp, vp, vp, P(C.c_float)]),
    "mpcvr_debug_get_device_params": (C.c_int32, [vp, C.c_int32, vp, C.c_size_t]),
    "mpcvr_debug_set_dovi": (C.c_int32, [vp, vp]),
    "mpcvr_debug_build_src_desc": (C.c_int32, [vp, vp, vp, vp, C.c_size_t]),
    "mpcvr_debug_axis_taps": (C.c_int32, [vp, C.c_int32, C.c_int32, P(C.c_int32), vp, C.c_int32, P(C.c_float)]),
    "mpcvr_debug_capture_f32": (C.c_int32, [vp, C.c_int32]),
    "mpcvr_debug_read_surface": (C.c_int32, [vp, C.c_int32, vp, C.c_size_t, P(C.c_int32), P(C.c_int32), P(C.c_int32)]),
    "mpcvr_debug_render_from_convert": (C.c_int32, [vp, vp, C.c_size_t]),
    "mpcvr_debug_math": (C.c_int32, [C.c_int32, C.c_int32, vp, vp, vp, C.c_int64]),
    "mpcvr_debug_probe_pq": (C.c_int32, [vp, vp, vp, C.c_int32, C.c_int32]),
}

_lib = None


def lib():
    """Load libmpcvr_b200.so.  Raises if it has not been built: there is no Python/CPU implementation."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the frame path has no CPU fallback)" % path)
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
