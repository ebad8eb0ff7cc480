"""Operand vectors for the per-operation checks of the device numeric types: HDRFloat<float>,
HDRFloatComplex<float>, dblflt, dbldbl.  Used by the GPU suite (fs_selftest_numeric_op vs the oracle) and by the CPU suite
(the host build of the same product headers vs the oracle)."""
import numpy as np

_HF = np.dtype([("m", "<f4"), ("e", "<i4")])
_HC = np.dtype([("re", "<f4"), ("im", "<f4"), ("e", "<i4")])
_DF = np.dtype([("head", "<f4"), ("tail", "<f4")])
_DD = np.dtype([("head", "<f8"), ("tail", "<f8")])
_MIN_BIG = -(1 << 28)  # MIN_BIG_EXPONENT = INT32_MIN >> 3 (HDRFloat.h:50-58)


def _hdr_mantissas(rng, n, reduced):
    """Signed mantissas: reduced ones in [1, 2); otherwise also what the kernels hold between Reduce calls (products, sums
    after cancellation: anything from 2^-40 to 2^40), zeros and binary32 subnormals."""
    m = rng.uniform(1.0, 2.0, n).astype(np.float32)
    if not reduced:
        m = (m * np.exp2(rng.integers(-40, 41, n)).astype(np.float32)).astype(np.float32)
        k = rng.integers(0, 16, n)
        m[k == 0] = 0.0
        m[k == 1] = np.float32(1e-41)                      # subnormal
        m[k == 2] = np.float32(1.0)
        m[k == 3] = np.nextafter(np.float32(2.0), np.float32(0.0))
    return (m * rng.choice(np.array([-1.0, 1.0], np.float32), n)).astype(np.float32)


def _hdr_exponents(rng, n, base):
    """Exponents around `base` with every gap the alignment code distinguishes (0, +-1, +-119..+-121, +-126..+-129, far)."""
    gaps = np.array([0, 1, -1, 2, -2, 23, -24, 30, -30, 119, -119, 120, -120, 121, -121, 126, -126, 127, -127, 128, -128, 129, -129,
                     1000, -1000], np.int64)
    e = base + rng.choice(gaps, n) + rng.integers(-3, 4, n)
    k = rng.integers(0, 24, n)
    e[k == 0] = _MIN_BIG
    e[k == 1] = _MIN_BIG + rng.integers(0, 200, n)[k == 1]
    return np.clip(e, _MIN_BIG, 20_000_000).astype(np.int32)


def _hf_operands(rng, n, reduced):
    a = np.zeros(n, _HF)
    a["m"] = _hdr_mantissas(rng, n, reduced)
    a["e"] = _hdr_exponents(rng, n, int(rng.integers(-5000, 5000)))
    z = a["m"] == 0
    a["e"][z & (rng.integers(0, 2, n) == 0)] = _MIN_BIG   # the canonical zero, and zeros that kept an exponent
    return a


def _hc_operands(rng, n, reduced):
    a = np.zeros(n, _HC)
    a["re"] = _hdr_mantissas(rng, n, reduced)
    a["im"] = _hdr_mantissas(rng, n, False) * np.exp2(-rng.integers(0, 30, n)).astype(np.float32)
    a["e"] = _hdr_exponents(rng, n, int(rng.integers(-5000, 5000)))
    return a


def _dd_operands(rng, n, dtype, eps_bits):
    a = np.zeros(n, dtype)
    ft = dtype["head"].type
    head = (rng.uniform(1.0, 2.0, n) * np.exp2(rng.integers(-30, 31, n)) * rng.choice([-1.0, 1.0], n)).astype(ft)
    tail = (head.astype(np.float64) * np.exp2(-eps_bits - rng.integers(0, 8, n)) * rng.uniform(-1.0, 1.0, n)).astype(ft)
    k = rng.integers(0, 12, n)
    head[k == 0] = 0
    tail[k <= 1] = 0
    a["head"], a["tail"] = head, tail
    return a



OPS = [
    (0, "HDRFloat add"), (1, "HDRFloat subtract"), (2, "HDRFloat multiply"), (3, "HDRFloat square"), (4, "HDRFloat Reduce"),
    (5, "HDRFloat divide"), (6, "HDRFloat compareToBothPositiveReduced"),
    (10, "HDRFloatComplex plus"), (11, "HDRFloatComplex times"), (12, "HDRFloatComplex Reduce"),
    (13, "HDRFloatComplex chebychevNorm"), (14, "HDRFloatComplex times HDRFloat"),
    (20, "dblflt add"), (21, "dblflt sub"), (22, "dblflt mul"), (23, "dblflt sqr"),
    (30, "dbldbl add"), (31, "dbldbl sub"), (32, "dbldbl mul"),
    (40, "HDRx32 perturbation step (custom_perturb2)"),
    (50, "HDRFloat<double> add"), (51, "HDRFloat<double> subtract"), (52, "HDRFloat<double> multiply"),
    (53, "HDRFloat<double> square"), (54, "HDRFloat<double> Reduce"), (55, "HDRFloat<double> divide"),
    (56, "HDRFloat<double> compareToBothPositiveReduced"),
]
_HD = np.dtype([("m", "<f8"), ("e", "<i4"), ("pad", "<i4")])


def _hd_operands(rng, n, reduced):
    """HDRFloat<double> operands: as _hf_operands, with 53-bit mantissas (unreduced ones from 2^-300 to 2^300, subnormals)."""
    a = np.zeros(n, _HD)
    m = rng.uniform(1.0, 2.0, n)
    if not reduced:
        m = m * np.exp2(rng.integers(-300, 301, n).astype(np.float64))
        k = rng.integers(0, 16, n)
        m[k == 0] = 0.0
        m[k == 1] = 5e-320                                   # subnormal
        m[k == 2] = 1.0
        m[k == 3] = np.nextafter(2.0, 0.0)
    a["m"] = m * rng.choice(np.array([-1.0, 1.0]), n)
    a["e"] = _hdr_exponents(rng, n, int(rng.integers(-5000, 5000)))
    z = a["m"] == 0
    a["e"][z & (rng.integers(0, 2, n) == 0)] = _MIN_BIG
    return a
_STEP = np.dtype([("m0", "<f4"), ("e0", "<i4"), ("m1", "<f4"), ("e1", "<i4"), ("m2", "<f4"), ("e2", "<i4")])


def operands(op, n):
    """Operand pairs for `op`, aimed at the code's case distinctions: exponent gaps 0, +-1, +-119..121, +-126..129, far apart,
    the MIN_BIG exponent, zeros with and without it, unreduced and subnormal mantissas."""
    rng = np.random.default_rng(1000 + op)
    if 50 <= op <= 56:
        reduced = op == 56
        a, b = _hd_operands(rng, n, reduced), _hd_operands(rng, n, reduced)
        if op == 55:
            b["m"][b["m"] == 0] = 1.5
        if op == 56:
            a["m"], b["m"] = np.abs(a["m"]), np.abs(b["m"])
            same = rng.integers(0, 3, n) == 0
            b["e"][same] = a["e"][same]
        return a, b
    if op == 40:
        # reduced dX, dY, Zx, Zy, cX, cY (what the kernels hold at a step): exponents so that the two product sums and the
        # added c meet at every gap the alignment distinguishes, including the 127 rule of custom_perturb2
        a, b = np.zeros(n, _STEP), np.zeros(n, _STEP)
        for arr in (a, b):
            for k in range(3):
                h = _hf_operands(rng, n, True)
                arr["m%d" % k] = h["m"]
                arr["e%d" % k] = h["e"]
        # orbit elements Zx, Zy live near exponent 0..-60; deltas and c far below or near them
        a["e2"] = rng.integers(-60, 2, n).astype(np.int32)
        b["e0"] = rng.integers(-60, 2, n).astype(np.int32)
        near = rng.integers(0, 2, n) == 0
        for arr, f in ((a, "e0"), (a, "e1"), (b, "e1"), (b, "e2")):
            arr[f] = np.where(near, rng.integers(-140, 2, n), arr[f]).astype(np.int32)
        return a, b
    if op <= 6:
        reduced = op == 6  # the comparison is specified for reduced operands only
        a, b = _hf_operands(rng, n, reduced), _hf_operands(rng, n, reduced)
        if op == 5:
            b["m"][b["m"] == 0] = np.float32(1.5)          # x / 0 is not an operation the kernels perform
        if op == 6:
            a["m"], b["m"] = np.abs(a["m"]), np.abs(b["m"])
            same = rng.integers(0, 3, n) == 0              # equal exponents: the mantissas decide
            b["e"][same] = a["e"][same]
    elif op <= 14:
        a, b = _hc_operands(rng, n, False), _hc_operands(rng, n, False)
        if op == 14:
            b["re"] = _hdr_mantissas(rng, n, True)
    elif op <= 23:
        a, b = _dd_operands(rng, n, _DF, 24), _dd_operands(rng, n, _DF, 24)
    else:
        a, b = _dd_operands(rng, n, _DD, 53), _dd_operands(rng, n, _DD, 53)
    return a, b


def mismatches(got, want):
    """Indices where the results differ.  Not compared: the sign of an exactly-zero mantissa (where the reference returns an
    operand untouched -- exponent gap >= 120 -- it keeps that operand's -0, the library's forms give +0; no consumer can tell:
    zeros are tested with == 0, compared, multiplied or passed through |x|, and Reduce / the exponent extraction return
    before looking at a zero's bits) and NaN payloads (a NaN result is a NaN result)."""
    got, want = got.copy(), want.copy()
    n = got.shape[0]
    for f in got.dtype.names:
        if got.dtype[f].kind == "f":
            got[f][got[f] == 0] = 0
            want[f][want[f] == 0] = 0
            both = np.isnan(got[f]) & np.isnan(want[f])
            got[f][both] = 0
            want[f][both] = 0
    return np.flatnonzero((got.view(np.uint8).reshape(n, -1) != want.view(np.uint8).reshape(n, -1)).any(axis=1))
