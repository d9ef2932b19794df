"""GPU tests added after the round's GPU budget was spent (they run for the first time at round end).  Kept in a file of their own, collected last,
so that the record of the 458 tests that did run on the B200 this round (profiles/r02c_gputests.txt) stays comparable test for test."""
import numpy as np
import pytest

import harness as H
import orclib as O
from test_gpu_parity import _check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tm_type", [1, 3, 5, 6])
@pytest.mark.parametrize("trims", ["negative_offset", "zero_satgain"])
def test_dolby_vision_tonemap_trims_nan_pixels(tm_type, trims):
    """The same unguarded trims inside ps_hdr10_tonemap.hlsl (HDR display + local tone mapping): curves 1-4 saturate the NaN to 0 and carry on,
    BT.2390 spreads it over the pixel when its curve is active, ST 2094-10 passes it through per channel.  tests/test_reference_tonemap_exec_pin.py
    holds the oracle to the executed shader; tests/test_device_code_on_host.py holds tm::pixel, compiled for the host, to the oracle; here the
    kernel on the GPU."""
    import ctypes as C
    l2 = dict(negative_offset=[(3079, 2100, 1900, 2010, 2100, 2200)], zero_satgain=[(3079, 2100, 2000, 2048, 2200, 2048)])[trims]
    dv = H.make_dovi(83, mmr=1, l2=l2, source_max_pq=3696, l1=(40, 3300, 1500), l3=(2040, 2100, 2000))
    nits = (C.c_float * 3)()
    assert O.lib().orc_dovi_l1_nits(C.addressof(dv), nits) == 1
    hdr = dict(out=1, tonemap=1, min=nits[0], max=nits[1], cll=nits[1], fall=nits[2], display=1000.0, type=6 if tm_type == 5 else tm_type, settings_type=tm_type)
    ex = O.exfmt(chroma=5, range_=1, matrix=0, prim=0, trc=0)
    case = H.Case(O.CF["P010"], 64, 36, 64, 36, ex, backbuffer=O.SURF_RGB10A2, up=O.UPSCALE["Nearest"], dovi=dv, hdr=hdr, name="dovi_tm_nan_%s_%d" % (trims, tm_type))
    _check(case, "fullnoise", 83)
