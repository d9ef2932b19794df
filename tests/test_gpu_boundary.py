"""GPU tests of the boundary verbs around the frame path (Source/VideoProcessor.h:178-236): device-memory samples
(IMediaSampleD3D11 analogue), caller-bound back buffers, Render / Flush / FillBlack, GetCurentImage / GetDisplayedImage and the
ISubRender11 callback hook -- each compared with the oracle where pixels are involved."""
import ctypes as C

import numpy as np
import pytest

import harness as H
import orclib as O

pytestmark = pytest.mark.gpu


def _frame(case, kind="natural", seed=1):
    codes = H.make_codes(case.cf, case.w, case.h, kind, seed)
    return H.pack(case.cf, case.w, case.h, codes)


def test_device_memory_sample_and_bound_backbuffer():
    """planes already in device memory are read in place; the output goes to a caller-owned device buffer"""
    torch = pytest.importorskip("torch")
    case = H.Case(O.CF["P010"], 96, 54, 192, 108, H.EX_2020_PQ, up=O.UPSCALE["Lanczos3"], name="devsample")
    buf, pitch, planes, pitches = _frame(case)
    ref = H.run_oracle(case, planes, pitches)
    vp = case.make_processor(device=0)
    dev_pitch = 256                                           # padded device rows (decoder surfaces are pitched)
    y = torch.zeros((case.h, dev_pitch), dtype=torch.uint8, device="cuda:0"); uv = torch.zeros((case.h // 2, dev_pitch), dtype=torch.uint8, device="cuda:0")
    host = torch.from_numpy(buf.copy())
    y[:, :pitch] = host[:pitch * case.h].view(case.h, pitch).cuda(); uv[:, :pitch] = host[pitch * case.h:].view(case.h // 2, pitch).cuda()
    out_pitch = 192 * 4 + 256
    target = torch.zeros((108, out_pitch), dtype=torch.uint8, device="cuda:0")
    vp.BindBackbuffer(target.data_ptr(), out_pitch)
    torch.cuda.synchronize()
    vp.ProcessSample(device_planes=[y.data_ptr(), uv.data_ptr()], device_pitches=[dev_pitch, dev_pitch])
    vp.WaitIdle()
    got = target.cpu().numpy()[:, :192 * 4]
    assert H.compare(np.ascontiguousarray(got), ref, case.backbuffer) == (0, 0)
    assert not target.cpu().numpy()[:, 192 * 4:].any()      # nothing written past the row


def test_render_again_flush_and_fill_black():
    case = H.Case(O.CF["NV12"], 64, 36, 96, 54, H.EX_709, up=O.UPSCALE["CatmullRom"], name="render")
    buf, pitch, planes, pitches = _frame(case)
    ref = H.run_oracle(case, planes, pitches)
    vp = case.make_processor(device=0)
    vp.ProcessSample(data=buf, pitch=pitch)
    a = vp.ReadBackbuffer()
    vp.FillBlack()
    assert not vp.ReadBackbuffer().any()
    vp.Render(1, 0)                                           # Render(field, t) re-runs Process on the current source
    b = vp.ReadBackbuffer()
    assert np.array_equal(a, b) and H.compare(a, ref, case.backbuffer) == (0, 0)
    vp.Flush()
    import videorenderer_b200 as V
    assert vp._L.mpcvr_render(vp._h, 1, 0) == V.HRESULT.E_UNEXPECTED


@pytest.mark.parametrize("rot", [0, 90])
def test_get_current_image_is_the_source_sized_frame(rot):
    """GetCurentImage renders at video size into BGRX with m_videoRect = m_windowRect = image rect (:3495-3557)"""
    case = H.Case(O.CF["P010"], 64, 36, 160, 90, H.EX_2020_PQ, backbuffer=O.SURF_RGB10A2, up=O.UPSCALE["Lanczos2"], rotation=rot, name="curimg")
    buf, pitch, planes, pitches = _frame(case)
    vp = case.make_processor(device=0)
    vp.ProcessSample(data=buf, pitch=pitch)
    before = vp.ReadBackbuffer().copy()
    img = vp.GetCurentImage()
    w, h = (36, 64) if rot else (64, 36)
    assert img.shape == (h, w * 4)
    same = H.Case(O.CF["P010"], 64, 36, w, h, H.EX_2020_PQ, backbuffer=O.SURF_BGRA8, up=O.UPSCALE["Lanczos2"], rotation=rot)
    assert H.compare(img, H.run_oracle(same, planes, pitches), O.SURF_BGRA8) == (0, 0)
    vp.Render(1, 0)                                           # rects / formats were restored
    assert np.array_equal(vp.ReadBackbuffer(), before)


@pytest.mark.parametrize("bb", [O.SURF_BGRA8, O.SURF_RGB10A2])
def test_get_displayed_image_is_the_back_buffer_as_bgrx(bb):
    case = H.Case(O.CF["NV12"], 64, 36, 96, 54, H.EX_709, backbuffer=bb, up=O.UPSCALE["Mitchell"], name="dispimg")
    buf, pitch, planes, pitches = _frame(case)
    vp = case.make_processor(device=0)
    vp.ProcessSample(data=buf, pitch=pitch)
    img = vp.GetDisplayedImage()
    back = H.rgb_view(vp.ReadBackbuffer(), bb)
    got = H.rgb_view(img, O.SURF_BGRA8)
    if bb == O.SURF_BGRA8:
        assert np.array_equal(got, back)
    else:                                                     # 10 -> 8 bit: UNORM read then UNORM write (exact for the conversion blit)
        exp = np.floor((back.astype(np.float32) / np.float32(1023)) * np.float32(255) + np.float32(0.5)).astype(np.int32)
        assert np.array_equal(got, exp)


def test_subtitle_callback_sees_the_frame_on_its_stream():
    """ISubRender11Callback::Render11 analogue: called after the pixel kernels with the rects of :3262-3282"""
    import videorenderer_b200._abi as A
    case = H.Case(O.CF["NV12"], 64, 36, 96, 54, H.EX_709, dst_rect=(10, 6, 106, 60), win=(120, 70), up=O.UPSCALE["CatmullRom"], name="sub")
    buf, pitch, planes, pitches = _frame(case)
    vp = case.make_processor(device=0)
    seen = []

    def render(user, rt_start, rt_stop, avg, cropped, original, viewport, stretch, xoff, flags, stream, target, tpitch, fmt):
        seen.append((rt_start, rt_stop, cropped.tuple(), viewport.tuple(), stream, target, tpitch, fmt))
        return 0

    devs = []
    vp.SetSubCallback(render, lambda user, dev: devs.append(dev) or 0)
    vp.ProcessSample(data=buf, pitch=pitch, rt_start=400000, rt_end=800000)
    vp.WaitIdle()
    assert len(seen) == 1
    rt_start, rt_stop, cropped, viewport, stream, target, tpitch, fmt = seen[0]
    assert (rt_start, rt_stop) == (400000, 800000) and cropped == (10, 6, 106, 60) and viewport == (0, 0, 120, 70)
    ptr, bpitch, w, h, f = vp.GetBackbuffer()
    assert target == ptr and tpitch == bpitch and fmt == f and stream == vp.GetStream()
    assert devs == [0]


def test_midstream_changes_rebuild_the_plan():
    """one instance, one frame, a sequence of Configure / rect / ProcAmp / rotation / format changes (Source/DX11VideoProcessor.cpp:3800-4050,
    SetVideoRect, SetProcAmpValues, SetRotation): after each change Render() must equal the oracle configured from scratch"""
    base = dict(cf=O.CF["P010"], w=64, h=36, ex=H.EX_2020_PQ)
    frame_case = H.Case(base["cf"], 64, 36, 128, 72, base["ex"], up=O.UPSCALE["Lanczos3"])
    buf, pitch, planes, pitches = _frame(frame_case, "natural", 7)
    vp = frame_case.make_processor(device=0)
    vp.ProcessSample(data=buf, pitch=pitch)

    def expect(**kw):
        c = H.Case(base["cf"], 64, 36, kw.pop("dw", 128), kw.pop("dh", 72), base["ex"], **kw)
        return H.run_oracle(c, planes, pitches), c.backbuffer

    def check(tag, **kw):
        vp.Render(1, 0)
        ref, fmt = expect(**kw)
        assert H.compare(vp.ReadBackbuffer(), ref, fmt) == (0, 0), tag

    check("initial", up=O.UPSCALE["Lanczos3"])
    s = vp.GetSettings(); s.iUpscaling = O.UPSCALE["CatmullRom"]; vp.Configure(s)
    check("upscaler", up=O.UPSCALE["CatmullRom"])
    s.iSDRDisplayNits = 200; s.bUseDither = 0; vp.Configure(s)
    check("nits+dither", up=O.UPSCALE["CatmullRom"], nits=200, dither=0)
    s.iTexFormat = O.TEXFMT_16FLOAT; s.bUseDither = 1; vp.Configure(s)
    check("texfmt", up=O.UPSCALE["CatmullRom"], nits=200, tex=O.TEXFMT_16FLOAT)
    vp.SetProcAmpValues(brightness=8.0, contrast=1.15, hue=10.0, saturation=0.9)
    pa = dict(brightness=8.0, contrast=1.15, hue=10.0, saturation=0.9)
    check("procamp", up=O.UPSCALE["CatmullRom"], nits=200, tex=O.TEXFMT_16FLOAT, procamp=pa)
    vp.SetWindowRect((0, 0, 100, 80)); vp.SetVideoRect((5, 10, 95, 61))
    check("rects", dw=100, dh=80, win=(100, 80), dst_rect=(5, 10, 95, 61), up=O.UPSCALE["CatmullRom"], nits=200, tex=O.TEXFMT_16FLOAT, procamp=pa)
    s.bConvertToSdr = 0; vp.Configure(s)
    check("no sdr", dw=100, dh=80, win=(100, 80), dst_rect=(5, 10, 95, 61), up=O.UPSCALE["CatmullRom"], nits=200, tex=O.TEXFMT_16FLOAT, procamp=pa, to_sdr=0)
    vp.SetRotation(180); vp.SetFlip(True)
    check("turned", dw=100, dh=80, win=(100, 80), dst_rect=(5, 10, 95, 61), up=O.UPSCALE["CatmullRom"], nits=200, tex=O.TEXFMT_16FLOAT, procamp=pa, to_sdr=0,
          rotation=180, flip=1)
    vp.SetRotation(0); vp.SetFlip(False); vp.SetBackbufferFormat(O.SURF_RGB10A2)
    s.iDownscaling = O.DOWNSCALE["Lanczos"]; vp.Configure(s); vp.SetWindowRect((0, 0, 40, 20)); vp.SetVideoRect((0, 0, 40, 20))
    check("downscale to 10 bit", dw=40, dh=20, backbuffer=O.SURF_RGB10A2, down=O.DOWNSCALE["Lanczos"], up=O.UPSCALE["CatmullRom"], nits=200,
          tex=O.TEXFMT_16FLOAT, procamp=pa, to_sdr=0)
