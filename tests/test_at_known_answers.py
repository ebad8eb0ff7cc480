"""The reference's own known-answer tests for the AT shortcut (FractalSharkTest/TestATInfo.cpp:91-148:
PerformAT fixed point, immediate escape, period 2, known escape time, step length), replayed through this
repository's implementations of that loop:

* the CUDA path (``-m gpu``): a hand-built ATInfo is uploaded through the C-ABI with an empty LA table and the
  frame is rendered in LAO mode (AT, then LA stages, no perturbation), so every pixel's iteration count is exactly
  ``bla_iterations`` of PerformAT -- for T = double / IterType = uint64 (the instantiation the reference tests) and
  for T = HDRFloat<float>, where c = +-1 goes through the mantissa-recurrence form and c = 0 / c = 100 through the
  general float+exponent loop;
* the CPU oracle (HDRFloat<float>), in the CPU suite.
"""
import ctypes as C
import math
import struct

import numpy as np
import pytest

import oracle_cpu
from fractalshark_b200 import _native as N
from fractalshark_b200 import Numeric
from fractalshark_b200 import RenderAlgorithm as A
from fractalshark_b200.gpu_renderer import GPURenderer

# (name, SqrEscapeRadius, ThresholdC, StepLength, RefC, max iterations, expected bla_iterations)
KNOWN = [
    ("fixed_point", 256.0, 10.0, 1, (0.0, 0.0), 100, 100),          # c = 0: z stays 0, all 100 passes
    ("immediate_escape", 256.0, 1000.0, 1, (100.0, 0.0), 100, 1),   # z1 = 100, |z1|^2 > 256
    ("period_2", 256.0, 1000.0, 1, (-1.0, 0.0), 100, 100),          # 0 -> -1 -> 0 ...
    ("known_escape_time", 256.0, 1000.0, 1, (1.0, 0.0), 100, 4),    # 0, 1, 2, 5, 26: escapes at step 4
    ("step_length_5", 256.0, 1000.0, 5, (1.0, 0.0), 100, 20),       # bla_iterations = bla_steps * 5
]


def hdr(v):
    """HDRFloat<float> {mantissa, exp}, reduced."""
    if v == 0.0:
        return struct.pack("<fi", 0.0, -(1 << 28))
    m, e = math.frexp(v)
    return struct.pack("<fi", m * 2.0, e - 1)


def hdrc(re, im):
    """HDRFloatComplex<float> {re, im, exp}, reduced on the larger part."""
    if re == 0.0 and im == 0.0:
        return struct.pack("<ffi", 0.0, 0.0, -(1 << 28))
    _, e = math.frexp(max(abs(re), abs(im)))
    return struct.pack("<ffi", math.ldexp(re, 1 - e), math.ldexp(im, 1 - e), e - 1)


def at_blob(numeric, sqr_escape, threshold_c, step, ref_c):
    """ATInfo<IterType, T, SubType> in the reference layout (ATInfo.h:80-89); ZCoeff = CCoeff = 1."""
    one, zero = (1.0, 0.0), (0.0, 0.0)
    if numeric == Numeric.F64:   # IterType = uint64_t
        cx = lambda c: struct.pack("<dd", *c)
        return (struct.pack("<Qdd", step, threshold_c, sqr_escape) + cx(ref_c) + cx(one) + cx(one) + cx(one) + cx(one) +
                cx(one) + struct.pack("<ddd", 1.0, ref_c[0] ** 2 + ref_c[1] ** 2, 2.0 ** 32))
    return (struct.pack("<I", step) + hdr(threshold_c) + hdr(sqr_escape) + hdrc(*ref_c) + hdrc(*one) + hdrc(*one) +
            hdrc(*one) + hdrc(*one) + hdrc(*one) + hdr(1.0) + hdr(ref_c[0] ** 2 + ref_c[1] ** 2) + hdr(2.0 ** 32))


class _FlatOrbit:
    """Two zero orbit entries: the LAO-mode render never reads them, the upload needs an orbit."""

    def __init__(self, numeric):
        self.numeric, self.count, self.uncompressed_count, self.period, self.pextras = numeric, 2, 2, 0, 0
        self.elem_bytes = 16
        self._data = np.zeros(2 * 16, np.uint8)
        if numeric == Numeric.HDR32:   # zero = mantissa 0, exponent MIN_BIG
            self._data.view(np.int32)[[1, 2, 5, 6]] = -(1 << 28)
        self._low = np.zeros(16, np.uint8)

    def descriptor(self):
        return N.FsOrbit(self._data.ctypes.data, 2, 2, 0, self._low.ctypes.data, self._low.ctypes.data)


class _AtOnlyTable:
    def __init__(self, blob):
        self._at = np.frombuffer(blob, np.uint8).copy()
        self._pad = np.zeros(256, np.uint8)
        self.at_bytes = len(blob)

    def descriptor(self):
        # no LA records, no stages: LAStageCount = 0, UseAT, IsValid
        return N.FsLaReference(self._pad.ctypes.data, 0, self._pad.ctypes.data, 0, self._at.ctypes.data, 0, 1, 1)


def zero_coords(numeric):
    z = struct.pack("<d", 0.0) if numeric == Numeric.F64 else hdr(0.0)
    return {k: z for k in ("cx", "cy", "dx", "dy", "center_x", "center_y")}


@pytest.mark.gpu
@pytest.mark.parametrize("numeric,alg,iter_bytes", [(Numeric.F64, A.Gpu1x64PerturbedLAv2LAO, 8),
                                                    (Numeric.HDR32, A.GpuHDRx32PerturbedLAv2LAO, 4)],
                         ids=["double_u64", "hdr32_u32"])
@pytest.mark.parametrize("case", KNOWN, ids=[k[0] for k in KNOWN])
def test_perform_at_known_answers_on_the_cuda_path(case, numeric, alg, iter_bytes):
    _, sqr, thr, step, ref_c, n_iter, want = case
    w, h = 16, 8
    r = GPURenderer()
    assert r.InitializeMemory(w, h, 1, iter_bytes=iter_bytes) == 0
    assert r.InitializePerturb(1, _FlatOrbit(numeric), 0, None, _AtOnlyTable(at_blob(numeric, sqr, thr, step, ref_c))) == 0
    r.ClearMemory()
    assert r.RenderPerturbLAv2(alg, zero_coords(numeric), n_iter) == 0
    rc, iters, _, _ = r.RenderCurrent(n_iter)
    assert rc == 0
    r.close()
    assert (iters[:h, :w] == want).all(), iters[:h, :w]


@pytest.mark.parametrize("case", KNOWN, ids=[k[0] for k in KNOWN])
def test_perform_at_known_answers_on_the_oracle(case):
    _, sqr, thr, step, ref_c, n_iter, want = case
    w, h = 16, 8
    got, _ = oracle_cpu.render_lav2(A.GpuHDRx32PerturbedLAv2LAO, w, h, zero_coords(Numeric.HDR32), _FlatOrbit(Numeric.HDR32),
                                    _AtOnlyTable(at_blob(Numeric.HDR32, sqr, thr, step, ref_c)), n_iter)
    assert (got[:h, :w] == want).all(), got[:h, :w]


# ---- the plan of the kernel's AT shortcut (fs_at_fast.cuh), CPU suite ------------------------------------------------
def test_at_plan_guards(built):
    """ok / mono decisions of atfast::plan: the mantissa recurrence needs a reduced c with exponent <= 0; the lean
    chunk test (escape looked at on a chunk's last pass only) additionally needs R > 4 and |c| <= R/4."""
    # c = 1.5 x 2^-1 = 0.75, R = 256 = 1.0 x 2^8: both forms apply; the threshold is R / 2^(2E)
    ok, mono, E, thr = oracle_cpu.at_plan(1.5, 0.0, -1, 1.0, 8)
    assert (ok, mono, E, thr) == (True, True, -1, 1024.0)
    # |c| >= 2 (exponent 1): the reference's exponent doubles every pass, only the general loop reproduces that
    assert oracle_cpu.at_plan(1.0, 0.0, 1, 1.0, 8)[:2] == (False, False)
    # R = 4 is not "> 4" (ATInfo::Usable demands more, ATInfo.h:91-105): recurrence yes, lean test no
    assert oracle_cpu.at_plan(1.5, 0.0, -1, 1.0, 2)[:2] == (True, False)
    # R = 5: |c| = 1.5 > R/4 = 1.25 -> running maximum; |c| = 1.0 <= 1.25 -> lean
    assert oracle_cpu.at_plan(1.5, 0.0, 0, 1.25, 2)[:2] == (True, False)
    assert oracle_cpu.at_plan(1.0, 0.0, 0, 1.25, 2)[:2] == (True, True)
    # an unreduced radius mantissa, a radius beyond the range the comparison is prepared for, NaN: general loop / no lean
    assert oracle_cpu.at_plan(1.0, 0.0, 0, 2.5, 8)[0] is False
    assert oracle_cpu.at_plan(1.0, 0.0, 0, 1.0, 50)[1] is False
    assert oracle_cpu.at_plan(float("nan"), 0.0, 0, 1.0, 8)[1] is False
    # a tiny c (exponent -100) with R = 256: |R.e - 2E| = 208 > 126 cannot be pre-scaled in binary32 -> general loop
    assert oracle_cpu.at_plan(1.0, 1.0, -100, 1.0, 8)[0] is False


def test_at_escape_is_monotone_where_the_plan_says_so(built):
    """Random starts around the escape radius, random c with |c| <= R/4: once a pass reads as escaped every later pass
    does (growing, inf or NaN), which is what lets the kernel test only the last pass of a chunk."""
    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(4000):
        r_e = int(rng.integers(3, 33))                       # R = rm * 2^r_e in (4, 2^33)
        rm = float(np.float32(rng.uniform(1.0, 2.0)))
        R = rm * 2.0 ** r_e
        ce = int(rng.integers(-12, 1))
        cre, cim = (float(np.float32(v)) for v in rng.uniform(-2.0, 2.0, 2))
        if max(abs(cre), abs(cim)) < 1.0:                    # keep c reduced: larger mantissa in [1, 2)
            cre = math.copysign(1.0 + abs(cre) % 1.0, cre or 1.0)
        ok, mono, E, thr = oracle_cpu.at_plan(cre, cim, ce, rm, r_e)
        if not (ok and mono):
            continue
        # start near the radius: |z| = sqrt(R) * (0.5 .. 2), mantissa units of 2^E
        mag = math.sqrt(R) * float(rng.uniform(0.5, 2.0)) / 2.0 ** E
        ang = float(rng.uniform(0, 2 * math.pi))
        first, stays = oracle_cpu.at_growth(cre, cim, ce, rm, r_e, float(np.float32(mag * math.cos(ang))),
                                            float(np.float32(mag * math.sin(ang))), 48)
        assert stays, (cre, cim, ce, rm, r_e, mag, ang, first)
        checked += first >= 0
    assert checked > 1000
