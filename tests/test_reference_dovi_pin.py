"""Fourth pin against reference code EXECUTED here: the Dolby Vision reshaping constants.

The reference turns the RPU's mapping curves into the `PS_DOVI_CURVE` / `PS_DOVI_POLY_CURVE` constant buffers in
`CDX11VideoProcessor::SetShaderDoviCurves` / `SetShaderDoviCurvesPoly` (Source/DX11VideoProcessor.cpp:990-1122).  Their loops are
compiled where they lie, together with the reference's own struct definitions (Include/IMediaSideData.h:154-349,
Source/Shaders.h:32-50), into oracle/_ref/libref_dovi.so (oracle/Makefile, oracle/ref_stub/dovi_*.{h,inc}; nothing is copied) and
compared bit for bit with what the oracle's `build_dovi` (oracle/orc_core.inc, SPEC mode) derives from the same metadata: pivots,
polynomial and MMR coefficients (packing of the MMR weights into float4 pairs included), methods / mmr_single / min_order /
max_order.  The product's kernels are compared with the oracle pixel by pixel in tests/test_gpu_parity.py (`test_dolby_vision*`);
this file ties the oracle's parameter derivation to the reference's."""
import ctypes as C
import os

import numpy as np
import pytest

import orclib as O

REF = os.path.join(O.ORC_DIR, "_ref", "libref_dovi.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_dovi.so not built (reference tree absent at build time)")

PIV = (C.c_float * 7) * 3
COEF = ((C.c_float * 4) * 8) * 3
MMR = ((C.c_float * 4) * 48) * 3
PAR = (C.c_int32 * 4) * 3


def _libs():
    R = C.CDLL(REF)
    R.ref_dovi_curves.restype = C.c_int
    R.ref_dovi_curves.argtypes = [C.POINTER(O.OrcDovi), PIV, COEF, MMR, PAR]
    R.ref_dovi_curves_poly.restype = C.c_int
    R.ref_dovi_curves_poly.argtypes = [C.POINTER(O.OrcDovi), PIV, COEF]
    L = O.lib()
    L.orc_test_dovi_curves.restype = C.c_int
    L.orc_test_dovi_curves.argtypes = [C.POINTER(O.OrcDovi), PIV, COEF, MMR, PAR]
    return R, L


def _random_metadata(rng, allow_mmr):
    d = O.OrcDovi()
    d.coef_log2_denom = int(rng.integers(13, 24))
    d.bl_bit_depth = int(rng.choice([8, 10, 12, 16]))
    top = (1 << d.bl_bit_depth) - 1
    for c in range(3):
        cv = d.curves[c]
        n = int(rng.integers(2, 10))                                   # [2, 9] pivots -> n - 1 pieces
        cv.num_pivots = n
        piv = np.sort(rng.integers(0, top + 1, n))
        for i in range(n):
            cv.pivots[i] = int(piv[i])
        for i in range(8):
            cv.mapping_idc[i] = int(rng.integers(0, 2)) if (allow_mmr and c > 0) or (allow_mmr and rng.random() < 0.3) else 0
            cv.poly_order[i] = int(rng.integers(0, 3))                 # 0 is not legal in a stream; the code's ">= 1" / ">= 2" tests are exercised
            for k in range(3):
                cv.poly_coef[i][k] = int(rng.integers(-(1 << (d.coef_log2_denom + 4)), 1 << (d.coef_log2_denom + 4)))
            cv.mmr_order[i] = int(rng.integers(1, 4))
            cv.mmr_constant[i] = int(rng.integers(-(1 << (d.coef_log2_denom + 2)), 1 << (d.coef_log2_denom + 2)))
            for j in range(3):
                for k in range(7):
                    cv.mmr_coef[i][j][k] = int(rng.integers(-(1 << (d.coef_log2_denom + 3)), 1 << (d.coef_log2_denom + 3)))
    for i in range(9):
        d.rgb_to_lms_matrix[i] = float(rng.uniform(0, 1)); d.ycc_to_rgb_matrix[i] = float(rng.uniform(-2, 2))
    return d


def _bits(a):
    return np.frombuffer(bytes(a), np.uint32)


@pytest.mark.parametrize("allow_mmr", [False, True])
def test_reshaping_constants_equal_the_references(allow_mmr):
    R, L = _libs()
    rng = np.random.Generator(np.random.PCG64(1234 + int(allow_mmr)))
    seen_mmr = seen_mixed = 0
    for _ in range(400):
        d = _random_metadata(rng, allow_mmr)
        rp, rc, rm, rq = PIV(), COEF(), MMR(), PAR()
        op, oc, om, oq = PIV(), COEF(), MMR(), PAR()
        assert R.ref_dovi_curves(C.byref(d), rp, rc, rm, rq) == 0
        has_mmr = L.orc_test_dovi_curves(C.byref(d), op, oc, om, oq)
        assert np.array_equal(_bits(rp), _bits(op)), "pivots"
        assert np.array_equal(_bits(rc), _bits(oc)), "coefficients"
        assert np.array_equal(_bits(rm), _bits(om)), "mmr weights"
        assert np.array_equal(np.frombuffer(bytes(rq), np.int32), np.frombuffer(bytes(oq), np.int32)), "methods / mmr_single / min_order / max_order"
        seen_mmr += int(has_mmr != 0)
        methods = np.frombuffer(bytes(oq), np.int32).reshape(3, 4)[:, 0]
        seen_mixed += int(np.any(methods == 3))
        if not has_mmr:
            # streams without MMR pieces take SetShaderDoviCurvesPoly in the reference (:2510-2514): same pivots and coefficients
            pp, pc = PIV(), COEF()
            assert R.ref_dovi_curves_poly(C.byref(d), pp, pc) == 0
            assert np.array_equal(_bits(pp), _bits(op)) and np.array_equal(_bits(pc), _bits(oc))
    if allow_mmr:
        assert seen_mmr > 100 and seen_mixed > 50          # MMR pieces, and curves that mix polynomial and MMR pieces, were covered
    else:
        assert seen_mmr == 0


def _with_levels(rng, d, kind):
    """Level 1 / 2 / 3 extension blocks in the metadata's slots (orc_dovi_ext: level + seven 16-bit fields)"""
    slots = list(rng.permutation(32)[:8])
    def put(level, vals):
        e = d.ext[int(slots.pop())]; e.level = level
        for i, v in enumerate(vals): e.v[i] = int(v)
    if kind != "no_l1":
        put(1, [rng.integers(0, 300), rng.integers(1500, 4096), rng.integers(300, 1500)])
        if rng.random() < 0.5: put(3, [rng.integers(2000, 2100), rng.integers(2000, 2100), rng.integers(2000, 2100)])
    n2 = {"none": 0, "one": 1, "many": int(rng.integers(2, 5)), "no_l1": 2}[kind]
    for _ in range(n2):
        put(2, [rng.integers(1500, 3900)] + [rng.integers(1500, 2600) for _ in range(5)] + [rng.integers(0, 4096)])
    d.source_min_pq = int(rng.integers(0, 100)); d.source_max_pq = int(rng.integers(2500, 4096))
    return d


def test_level_metadata_handling_equals_the_references():
    """CopySample's Dolby Vision level block (Source/DX11VideoProcessor.cpp:2324-2469) compiled where it lies: Level 1 (+ Level 3 offsets)
    -> nits (integer, through powf) and the Level 2 trim set chosen / interpolated for the display (between two targets, brighter than all
    -> towards neutral by source_max_pq, dimmer than all), then the constants of SetDolbyVisionDynamicParams (:956-960) -- against the
    oracle's orc_dovi_l1_nits and build_dovi, bit for bit, on displays from 100 to 4000 nits.  (std::lerp is this toolchain's libstdc++; the
    oracle and the product spell its case analysis out, so the test also holds them to it.)"""
    R, L = _libs()
    R.ref_dovi_dynamic.restype = C.c_int
    R.ref_dovi_dynamic.argtypes = [C.POINTER(O.OrcDovi), C.c_int, C.c_int, C.c_int, C.c_float * 3, C.POINTER(C.c_int32), C.c_float * 5, C.POINTER(C.c_int32)]
    L.orc_test_dovi_trims.restype = C.c_int
    L.orc_test_dovi_trims.argtypes = [C.POINTER(O.OrcDovi), C.c_int, C.c_float * 5]
    L.orc_dovi_l1_nits.restype = C.c_int
    L.orc_dovi_l1_nits.argtypes = [C.POINTER(O.OrcDovi), C.c_float * 3]
    rng = np.random.Generator(np.random.PCG64(77))
    scen = {"between": 0, "brighter": 0, "dimmer": 0, "none": 0, "l1": 0}
    for trial in range(600):
        kind = ("none", "one", "many", "no_l1")[trial % 4]
        d = _with_levels(rng, _random_metadata(rng, False), kind)
        for nits in (100, 300, 600, 1000, 2000, 4000):
            rl1, rp, rt, re = (C.c_float * 3)(), C.c_int32(), (C.c_float * 5)(), C.c_int32()
            assert R.ref_dovi_dynamic(C.byref(d), nits, 1, 1, rl1, C.byref(rp), rt, C.byref(re)) == 0
            ot = (C.c_float * 5)()
            oe = L.orc_test_dovi_trims(C.byref(d), nits, ot)
            assert oe == re.value, (trial, kind, nits)
            if oe:
                assert bytes(ot) == bytes(rt), (trial, kind, nits, list(ot), list(rt))
            ol1 = (C.c_float * 3)()
            op = L.orc_dovi_l1_nits(C.byref(d), ol1)
            assert op == rp.value
            if op:
                assert list(ol1) == list(rl1), (trial, list(ol1), list(rl1)); scen["l1"] += 1
            # which scenario was it (for the coverage assertion below)
            t2 = sorted(d.ext[i].v[0] for i in range(32) if d.ext[i].level == 2)
            if not t2: scen["none"] += 1
            else:
                pq = (nits / 10000.0) ** 0.1593017578125; pq = ((0.8359375 + 18.8515625 * pq) / (1 + 18.6875 * pq)) ** 78.84375
                lo = [t for t in t2 if t / 4095.0 <= pq]; hi = [t for t in t2 if t / 4095.0 > pq]
                scen["between" if lo and hi else ("brighter" if lo else "dimmer")] += 1
    assert all(v > 50 for v in scen.values()), scen
