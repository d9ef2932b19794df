"""Seventh pin: the reference's own shader text, EXECUTED.

The transfer functions of the convert shader live in three HLSL include files -- Shaders/convert/st2084.hlsl (ST2084ToLinear,
LinearToST2084), hlg.hlsl (inverse_HLG, HLGtoLinear) and hdr_tone_mapping.hlsl (hable, ToneMappingHable).  There is no HLSL compiler here,
but those files are plain expression code: with a small header that gives C++ HLSL's vector types and intrinsics
(oracle/ref_stub/hlsl_compat.h: componentwise float3 / float4, float literals, pow as fxc lowers it = exp2(y * log2 x), libm's
single-precision exp / exp2 / log2) they compile AS THEY ARE (oracle/Makefile cats them onto g++'s stdin, -ffp-contract=off; nothing is
copied) into oracle/_ref/libref_hlsl.so.  That is exactly what the oracle's LITERAL mode claims to be -- the shader text in its written
order, one IEEE operation per operation, libm -- so the two must agree BIT FOR BIT, and this file checks that they do.  SPEC mode
(the contract the CUDA kernels match) differs from LITERAL only by fused multiply-adds, reciprocal-multiplication for literal divisors
and the fixed exp2 / log2 polynomials; how close it stays is asserted here in ULPs of the result as well."""
import ctypes as C
import os

import numpy as np
import pytest

import orclib as O

REF = os.path.join(O.ORC_DIR, "_ref", "libref_hlsl.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_hlsl.so not built (reference tree absent at build time)")

FP = C.POINTER(C.c_float)


def _libs():
    R = C.CDLL(REF); L = O.lib()
    R.ref_hlsl_scalar.argtypes = [C.c_int, FP, FP, C.c_int, C.c_float]
    R.ref_hlsl_hlg_to_linear.argtypes = [FP, FP, C.c_int]
    L.orc_test_transfer.argtypes = [C.c_int, C.c_int, FP, FP, C.c_int, C.c_float]
    L.orc_test_hlg_to_linear_literal.argtypes = [FP, FP, C.c_int]
    return R, L


def _p(a):
    return a.ctypes.data_as(FP)


def _inputs(rng, lo, hi, n=200000):
    codes10 = np.arange(1024, dtype=np.float32) / np.float32(1023)
    x = np.concatenate([rng.uniform(lo, hi, n).astype(np.float32), codes10[(codes10 >= lo) & (codes10 <= hi)],
                        np.array([lo, hi, 0.5, np.nextafter(np.float32(0.5), np.float32(1)), np.nextafter(np.float32(0.5), np.float32(0))], np.float32)])
    return np.ascontiguousarray(x[(x >= lo) & (x <= hi)])


CASES = [(0, "ST2084ToLinear", 0.0, 1.0, 80.0), (0, "ST2084ToLinear", 0.0, 1.0, 1.0), (1, "LinearToST2084", 0.0, 1000.0, 1000.0), (1, "LinearToST2084", 0.0, 1.0, 1.0),
         (2, "inverse_HLG", 0.0, 1.0, 0.0), (3, "ToneMappingHable", 0.0, 80.0, 0.0)]


@pytest.mark.parametrize("fn,name,lo,hi,param", CASES, ids=[c[1] + "_%g" % c[4] for c in CASES])
def test_transfer_functions_equal_the_reference_hlsl(fn, name, lo, hi, param):
    R, L = _libs()
    rng = np.random.Generator(np.random.PCG64(100 + fn))
    x = _inputs(rng, lo, hi)
    ref = np.empty_like(x); lit = np.empty_like(x); spec = np.empty_like(x)
    assert R.ref_hlsl_scalar(fn, _p(x), _p(ref), x.size, param) == 0
    assert L.orc_test_transfer(fn, 1, _p(x), _p(lit), x.size, param) == 0
    bad = np.nonzero(ref.view(np.uint32) != lit.view(np.uint32))[0]
    assert bad.size == 0, (name, bad.size, x[bad[:5]], ref[bad[:5]], lit[bad[:5]])
    # SPEC against the reference text: a few ULP of the result (fused mads, reciprocal of literal divisors, polynomial exp2 / log2)
    assert L.orc_test_transfer(fn, 0, _p(x), _p(spec), x.size, param) == 0
    ok = np.isfinite(ref) & (np.abs(ref) > 1e-30)
    ulp = np.abs(spec[ok].astype(np.float64) - ref[ok]) / np.spacing(np.abs(ref[ok])).astype(np.float64)
    # measured on this toolchain (max / 99.9th percentile / median ULP of the result): ST2084ToLinear 950 / 540 / 93 -- `a - c1` cancels near
    # black and the 1/m1 = 6.3 exponent amplifies the last bits of the first pow; LinearToST2084 281 / 276 / 0 -- the m2 = 78.8 exponent;
    # inverse_HLG 4 / 3 / 0; ToneMappingHable 57 / 4 / 1 -- hable(x) - E/F cancels near 0.  Any fp32 evaluation of these expressions moves by
    # this much with the implementation of exp2 / log2 (a GPU's included); after the 10-bit quantiser SPEC and LITERAL frames differ in
    # < 1 % of the values, by one code (test_spec_arithmetic_stays_close_to_literal_and_f64).
    tol = {0: 2048, 1: 1024, 2: 16, 3: 128}[fn]
    assert ulp.max() <= tol, (name, float(ulp.max()), float(np.percentile(ulp, 99.9)))


def test_hlg_ootf_equals_the_reference_hlsl():
    """HLGtoLinear: inverse_HLG on the three channels, then the gain pow(2000 * dot(ootf_2020, rgb), 0.2) (hlg.hlsl:13-20).  The vector
    conditional of inverse_HLG only converts when the three lanes take the same branch, so the triplets lie on one side of 0.5."""
    R, L = _libs()
    rng = np.random.Generator(np.random.PCG64(5))
    lo = rng.uniform(0.0, 0.5, (100000, 3)).astype(np.float32); hi = rng.uniform(0.5, 1.0, (100000, 3)).astype(np.float32)
    hi[hi <= 0.5] = np.nextafter(np.float32(0.5), np.float32(1))
    x = np.ascontiguousarray(np.concatenate([lo, hi]))
    ref = np.empty_like(x); lit = np.empty_like(x)
    assert R.ref_hlsl_hlg_to_linear(_p(x), _p(ref), x.shape[0]) == 0
    assert L.orc_test_hlg_to_linear_literal(_p(x), _p(lit), x.shape[0]) == 0
    bad = np.nonzero(ref.view(np.uint32) != lit.view(np.uint32))
    assert bad[0].size == 0, (bad[0].size, x[bad[0][:3]], ref[bad[0][:3]], lit[bad[0][:3]])
