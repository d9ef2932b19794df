import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, harness as H
case = H.baseline_cases(0.1)[1]
vp = case.make_processor(device=0)
yuv = np.array([[0.5, 0.5, 0.5], [0.3, 0.5, 0.5], [0.9, 0.4, 0.6], [0.1, 0.5, 0.5]], np.float32)
for mode in (6, 8, 9, 10):
    out = np.empty_like(yuv)
    assert vp._L.mpcvr_debug_probe_pq(vp._h, yuv.ctypes.data, out.ctypes.data, len(yuv), mode) == 0
    print(mode, out.tolist())
