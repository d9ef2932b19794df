"""VideoProcessor -- Python mirror of the reference's CVideoProcessor verbs
(Source/VideoProcessor.h:168-236) on top of the C ABI.  Used by tests/ and bench.py the way
CMpcVideoRenderer uses its m_VideoProcessor: create -> InitMediaType -> Set*Rect -> ProcessSample*."""
import ctypes as C
import numpy as np

from . import _abi as A


def _surf_bpp(fmt):
    return 8 if fmt in (A.SURF.RGBA16F, 3) else (3 if fmt == A.SURF.RGB8 else 4)


class VideoProcessor:
    def __init__(self, settings=None, device=0, stream=None):
        self._L = A.lib()
        self._h = C.c_void_p()
        s = settings if settings is not None else A.Settings.default()
        hr = self._L.mpcvr_create(C.byref(s), device, stream, C.byref(self._h))
        if hr != 0:
            raise A.MpcvrError(hr, "mpcvr_create(device=%d) failed" % device)
        self._keep = None

    # -- lifetime
    def close(self):
        if self._h:
            self._L.mpcvr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, hr, what):
        if hr < 0:
            raise A.MpcvrError(hr, "%s: %s" % (what, self.last_error()))
        return hr

    def last_error(self):
        return (self._L.mpcvr_last_error(self._h) or b"").decode()

    # -- media type
    @staticmethod
    def media_type(cformat, width, height, ext_format=0, src_rect=None, frame_size=None, aspect=(0, 0), interlaced=0,
                   avg_time_per_frame=400000, bottom_up=0):
        mt = A.MediaType()
        mt.bottom_up = bottom_up
        mt.cformat = cformat; mt.width = width; mt.height = height
        if frame_size:
            mt.frame_width, mt.frame_height = frame_size
        if src_rect:
            mt.src_rect = A.Rect(*src_rect)
        mt.aspect_x, mt.aspect_y = aspect
        mt.ext_format = ext_format; mt.interlaced = interlaced; mt.avg_time_per_frame = avg_time_per_frame
        return mt

    def VerifyMediaType(self, mt):
        return bool(self._L.mpcvr_verify_media_type(self._h, C.byref(mt)))

    def InitMediaType(self, mt):
        return bool(self._L.mpcvr_init_media_type(self._h, C.byref(mt)))

    def GetAlignmentSize(self, mt):
        cx, cy = C.c_int32(), C.c_int32()
        ok = self._L.mpcvr_get_alignment_size(self._h, C.byref(mt), C.byref(cx), C.byref(cy))
        return bool(ok), cx.value, cy.value

    # -- geometry / settings
    def SetVideoRect(self, rect):
        self._L.mpcvr_set_video_rect(self._h, C.byref(A.Rect(*rect)))

    def SetWindowRect(self, rect):
        return self._check(self._L.mpcvr_set_window_rect(self._h, C.byref(A.Rect(*rect))), "SetWindowRect")

    def GetVideoRect(self):
        r = A.Rect(); self._L.mpcvr_get_video_rect(self._h, C.byref(r)); return r.tuple()

    def GetSourceRect(self):
        r = A.Rect(); self._L.mpcvr_get_source_rect(self._h, C.byref(r)); return r.tuple()

    def GetVideoSize(self):
        w, h = C.c_int32(), C.c_int32(); self._L.mpcvr_get_video_size(self._h, C.byref(w), C.byref(h)); return w.value, h.value

    def GetAspectRatio(self):
        w, h = C.c_int32(), C.c_int32(); self._L.mpcvr_get_aspect_ratio(self._h, C.byref(w), C.byref(h)); return w.value, h.value

    def SetRotation(self, deg):
        return self._L.mpcvr_set_rotation(self._h, deg)

    def SetFlip(self, flip):
        return self._L.mpcvr_set_flip(self._h, int(flip))

    def SetUpstreamConverter(self, present, pq_supported=False):
        return self._L.mpcvr_set_upstream_converter(self._h, int(bool(present)), int(bool(pq_supported)))

    def SetDisplayHdr(self, supported):
        return self._L.mpcvr_set_display_hdr(self._h, int(bool(supported)))

    def GetSettings(self):
        s = A.Settings(); self._L.mpcvr_get_settings(self._h, C.byref(s)); return s

    def Configure(self, settings):
        self._L.mpcvr_configure(self._h, C.byref(settings))

    def GetActive(self):
        return bool(self._L.mpcvr_get_active(self._h))

    def SetProcAmpValues(self, brightness=None, contrast=None, hue=None, saturation=None):
        v = A.ProcAmp(); flags = 0
        for bit, name, val in ((1, "brightness", brightness), (2, "contrast", contrast), (4, "hue", hue), (8, "saturation", saturation)):
            if val is not None:
                setattr(v, name, val); flags |= bit
        return self._check(self._L.mpcvr_set_procamp(self._h, flags, C.byref(v)), "SetProcAmpValues")

    def SetBackbufferFormat(self, fmt):
        return self._check(self._L.mpcvr_set_backbuffer_format(self._h, fmt), "SetBackbufferFormat")

    def BindBackbuffer(self, dev_ptr, pitch):
        return self._check(self._L.mpcvr_bind_backbuffer(self._h, dev_ptr, pitch), "BindBackbuffer")

    # -- frames
    def ProcessSample(self, data=None, pitch=0, device_planes=None, device_pitches=None, rt_start=0, rt_end=0, check=True,
                      hdr=None, hdr_cll=None, dovi=None):
        """hdr: dict(max_lum, min_lum[, primaries_x, primaries_y, white]) -> MediaSideDataHDR; hdr_cll: (MaxCLL, MaxFALL)"""
        s = A.Sample()
        s.rt_start, s.rt_end = rt_start, rt_end
        if hdr is not None:
            h = A.SideDataHDR()
            px = hdr.get("primaries_x", (0.265, 0.150, 0.680)); py = hdr.get("primaries_y", (0.690, 0.060, 0.320))
            for i in range(3):
                h.display_primaries_x[i] = px[i]; h.display_primaries_y[i] = py[i]
            h.white_point_x, h.white_point_y = hdr.get("white", (0.3127, 0.3290))
            h.max_display_mastering_luminance = hdr["max_lum"]; h.min_display_mastering_luminance = hdr["min_lum"]
            self._keep_hdr = h
            s.hdr = C.pointer(h)
        if dovi is not None:                       # an _abi.DoviMetadata, or a ctypes struct with the layout of its leading part
            if not isinstance(dovi, A.DoviMetadata):
                # e.g. the oracle's OrcDovi (curves, matrices, extension blocks): completed as a base-layer-only stream (profile 5 / 8);
                # a `header` attribute (dict) on the object overrides the acceptance fields
                d = A.DoviMetadata()
                C.memmove(C.byref(d), C.byref(dovi), min(C.sizeof(dovi), C.sizeof(d)))
                d.disable_residual_flag = 1
                for k, v in (getattr(dovi, "header", None) or {}).items():
                    if k == "nlq":
                        for i, q in enumerate(v):
                            for kk, vv in q.items():
                                setattr(d.nlq[i], kk, vv)
                    else:
                        setattr(d, k, v)
                dovi = d
            self._keep_dovi = dovi
            s.dovi = C.cast(C.pointer(dovi), C.POINTER(A.DoviMetadata))
        if hdr_cll is not None:
            c = A.SideDataHDRCLL(); c.MaxCLL, c.MaxFALL = hdr_cll
            self._keep_cll = c
            s.hdr_cll = C.pointer(c)
        if device_planes is not None:
            s.flags = A.SAMPLE_DEVICE_MEMORY
            for i, p in enumerate(device_planes):
                s.planes[i] = p
                s.pitches[i] = device_pitches[i]
        else:
            if isinstance(data, np.ndarray):
                self._keep = data
                s.data = data.ctypes.data; s.size = data.nbytes
            else:
                s.data = data[0]; s.size = data[1]
            s.pitch = pitch
        hr = self._L.mpcvr_process_sample(self._h, C.byref(s))
        return self._check(hr, "ProcessSample") if check else hr

    def Render(self, field=1, t=0):
        return self._check(self._L.mpcvr_render(self._h, field, t), "Render")

    def WaitIdle(self):
        return self._check(self._L.mpcvr_wait_idle(self._h), "WaitIdle")

    def FillBlack(self):
        return self._check(self._L.mpcvr_fill_black(self._h), "FillBlack")

    def Flush(self):
        self._L.mpcvr_flush(self._h)

    def GetBackbuffer(self):
        p = C.c_void_p(); pitch = C.c_size_t(); w, h, f = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self._L.mpcvr_get_backbuffer(self._h, C.byref(p), C.byref(pitch), C.byref(w), C.byref(h), C.byref(f)), "GetBackbuffer")
        return p.value, pitch.value, w.value, h.value, f.value

    def ReadBackbuffer(self, out=None):
        _, _, w, h, f = self.GetBackbuffer()
        if out is None:
            out = np.empty((h, w * _surf_bpp(f)), np.uint8)
        self._check(self._L.mpcvr_read_backbuffer(self._h, out.ctypes.data, out.strides[0]), "ReadBackbuffer")
        return out

    def GetCurentImage(self):
        w, h = C.c_int32(), C.c_int32()
        self._check(self._L.mpcvr_get_current_image(self._h, None, 0, C.byref(w), C.byref(h)), "GetCurentImage(size)")
        out = np.empty((h.value, w.value * 4), np.uint8)
        self._check(self._L.mpcvr_get_current_image(self._h, out.ctypes.data, out.strides[0], C.byref(w), C.byref(h)), "GetCurentImage")
        return out

    def GetDisplayedImage(self):
        w, h = C.c_int32(), C.c_int32()
        self._check(self._L.mpcvr_get_displayed_image(self._h, None, 0, C.byref(w), C.byref(h)), "GetDisplayedImage(size)")
        out = np.empty((h.value, w.value * 4), np.uint8)
        self._check(self._L.mpcvr_get_displayed_image(self._h, out.ctypes.data, out.strides[0], C.byref(w), C.byref(h)), "GetDisplayedImage")
        return out

    def GetVPInfo(self):
        buf = C.create_string_buffer(2048)
        self._check(self._L.mpcvr_get_vp_info(self._h, buf, 2048), "GetVPInfo")
        return buf.value.decode()

    def SetSubCallback(self, render_fn, set_device_fn=None, user=None):
        self._sub = (A.SUB_RENDER_FN(render_fn) if render_fn else A.SUB_RENDER_FN(),
                     A.SUB_SET_DEVICE_FN(set_device_fn) if set_device_fn else A.SUB_SET_DEVICE_FN())
        return self._check(self._L.mpcvr_set_sub_callback(self._h, self._sub[0], self._sub[1], user), "SetSubCallback")

    def ActivateDeviceDecoding(self, device, stream=None, flags=0):
        return self._L.mpcvr_activate_device_decoding(self._h, device, stream, flags)

    def GetDeviceIndex(self):
        return self._L.mpcvr_get_device_index(self._h)

    def GetStream(self):
        st = C.c_void_p(); self._check(self._L.mpcvr_get_stream(self._h, C.byref(st)), "GetStream"); return st.value

    def LastFrameKernels(self):
        n = C.c_int32(); self._L.mpcvr_get_last_frame_kernels(self._h, C.byref(n)); return n.value

    # -- debug hooks
    def debug_verify_gamma_table(self):
        return int(self._L.mpcvr_debug_verify_gamma_table(self._h))

    def debug_set_resize_four(self, enable):
        return self._L.mpcvr_debug_set_resize_four(int(enable))

    def debug_set_convert_two_lane(self, enable):
        return self._L.mpcvr_debug_set_convert_two_lane(int(enable))

    def debug_set_gamma_table(self, enable):
        return self._L.mpcvr_debug_set_gamma_table(int(bool(enable)))

    def debug_render_n(self, n):
        self._check(self._L.mpcvr_debug_render_n(self._h, int(n)), "debug_render_n")

    def debug_set_fused1x_vector(self, enable):
        return self._L.mpcvr_debug_set_fused1x_vector(int(enable))

    def debug_set_strip_kernel(self, enable, cfg=0):
        return self._L.mpcvr_debug_set_strip_kernel(int(bool(enable)), int(cfg))

    def debug_strip_status(self, clear=True):
        """watchdog record of the persistent 2x kernel: list of 8 ints, [0] = timed-out waits (0 expected)"""
        import ctypes as C
        out = (C.c_uint32 * 8)()
        hr = self._L.mpcvr_debug_strip_status(out, 1 if clear else 0)
        if hr != 0:
            raise RuntimeError("mpcvr_debug_strip_status failed")
        return list(out)

    def debug_set_path(self, path):
        return self._check(self._L.mpcvr_debug_set_path(self._h, path), "debug_set_path")

    def debug_plan(self):
        p = A.PlanInfo(); self._check(self._L.mpcvr_debug_get_plan(self._h, C.byref(p)), "debug_get_plan"); return p

    def debug_color_params(self):
        m = np.zeros((3, 3), np.float32); c = np.zeros(3, np.float32); g = np.zeros((3, 3), np.float32); lum = C.c_float()
        self._check(self._L.mpcvr_debug_get_color_params(self._h, m.ctypes.data, c.ctypes.data, g.ctypes.data, C.byref(lum)), "debug_color")
        return m, c, g, lum.value

    def debug_device_params(self, which):
        """bytes of the ColorParams (0) / DoviParams (1) / TonemapParams (2) block the kernels receive (csrc/frame_plan.h)"""
        n = self._check(self._L.mpcvr_debug_get_device_params(self._h, int(which), None, 0), "debug_device_params")
        buf = C.create_string_buffer(n)
        self._check(self._L.mpcvr_debug_get_device_params(self._h, int(which), buf, n), "debug_device_params")
        return buf.raw

    def debug_build_src_desc(self, planes, pitches):
        """bytes of the SrcDesc block (csrc/frame_plan.h) for planes at these addresses, built without a GPU"""
        pl = (C.c_void_p * 3)(*(list(planes) + [None] * 3)[:3]); pi = (C.c_int32 * 3)(*(list(pitches) + [0] * 3)[:3])
        n = self._check(self._L.mpcvr_debug_build_src_desc(self._h, pl, pi, None, 0), "debug_build_src_desc")
        buf = C.create_string_buffer(n)
        self._check(self._L.mpcvr_debug_build_src_desc(self._h, pl, pi, buf, n), "debug_build_src_desc")
        return buf.raw

    def debug_set_dovi(self, md):
        self._check(self._L.mpcvr_debug_set_dovi(self._h, C.addressof(md)), "debug_set_dovi")

    def debug_axis_taps(self, axis, i):
        base = C.c_int32(); wsum = C.c_float(); w = np.zeros(1024, np.float32)
        n = self._check(self._L.mpcvr_debug_axis_taps(self._h, axis, i, C.byref(base), w.ctypes.data, 1024, C.byref(wsum)), "debug_axis_taps")
        return base.value, w[:n].copy(), wsum.value

    def debug_capture_f32(self, enable=True):
        self._L.mpcvr_debug_capture_f32(self._h, int(enable))

    def debug_read_surface(self, which):
        w, h, f = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self._L.mpcvr_debug_read_surface(self._h, which, None, 0, C.byref(w), C.byref(h), C.byref(f)), "debug_read_surface(size)")
        bpp = 16 if which == 3 else _surf_bpp(f.value)
        out = np.empty((h.value, w.value * bpp), np.uint8)
        self._check(self._L.mpcvr_debug_read_surface(self._h, which, out.ctypes.data, out.strides[0], C.byref(w), C.byref(h), C.byref(f)), "debug_read_surface")
        return out, f.value

    def debug_render_from_convert(self, surf):
        return self._check(self._L.mpcvr_debug_render_from_convert(self._h, surf.ctypes.data, surf.strides[0]), "debug_render_from_convert")
