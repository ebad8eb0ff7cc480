"""TEST INFRASTRUCTURE: ctypes wrapper of oracle/_ref/libref_host.so -- the reference's own LAReference / BLAS table
builders compiled from /root/reference (oracle/Makefile, ref_host_harness.cpp).  Fed with the in-tree generator's orbit
so that what is compared is table construction alone."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_HOST_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_host.so")


def available() -> bool:
    return os.path.exists(REF_HOST_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_HOST_LIB)
        V, I, U64 = C.c_void_p, C.c_int, C.c_uint64
        L.refhost_build_la.restype = V
        L.refhost_build_la.argtypes = [I, V, U64, V, U64, I]
        L.refhost_la_info.argtypes = [V, I, C.POINTER(U64)]
        L.refhost_la_copy.argtypes = [V, I, V, V, V]
        L.refhost_free.argtypes = [V, I]
        L.refhost_cpu_lav2.restype = U64
        L.refhost_cpu_lav2.argtypes = [V, I, I, I, V, V, V, V, U64, V, I, I, I]
        L.refhost_build_blas.restype = U64
        L.refhost_build_blas.argtypes = [V, I, C.POINTER(C.c_int32)]
        L.refhost_blas_level.restype = U64
        L.refhost_blas_level.argtypes = [V, I, U64, V]
        _lib = L
    return _lib


class RefLaTable:
    """LAReference<IterType, HDRFloat<float>, float, Disable>::GenerateApproximationData on `orbit` (HDRx32).
    threading: 0 = reference default, 1 = single-threaded builder, 2 = multi-threaded builder."""

    def __init__(self, orbit, iter_bytes: int, n_iterations: int, threading: int = 0, with_blas: bool = False):
        from fractalshark_b200 import _native as N
        L = lib()
        radius = N.host_lib().fsh_orbit_max_radius(orbit._h)
        h = L.refhost_build_la(iter_bytes, orbit.data_ptr, orbit.count, radius, n_iterations, threading)
        info = (C.c_uint64 * 8)()
        L.refhost_la_info(h, iter_bytes, info)
        (self.num_las, self.num_stages, self.stage_count, use_at, valid, self.las_elem_bytes, self.at_bytes,
         stage_bytes) = (int(v) for v in info)
        self.use_at, self.is_valid = bool(use_at), bool(valid)
        self.las = np.zeros((self.num_las, self.las_elem_bytes), np.uint8)
        self.stages = np.zeros((self.num_stages, 2), np.uint32 if iter_bytes == 4 else np.uint64)
        self.at = np.zeros(self.at_bytes, np.uint8)
        assert stage_bytes == 2 * iter_bytes
        L.refhost_la_copy(h, iter_bytes, self.las.ctypes.data, self.stages.ctypes.data, self.at.ctypes.data)
        self.blas_levels, self.blas_lm2 = None, None
        if with_blas:
            lm2 = C.c_int32(0)
            n_levels = int(L.refhost_build_blas(h, iter_bytes, C.byref(lm2)))
            self.blas_lm2 = int(lm2.value)
            self.blas_levels = []
            for lv in range(n_levels):
                n = int(L.refhost_blas_level(h, iter_bytes, lv, None))
                a = np.zeros((n, 44), np.uint8)
                if n:
                    L.refhost_blas_level(h, iter_bytes, lv, a.ctypes.data)
                self.blas_levels.append(a)
        self._h, self._iter_bytes, self._orbit = h, iter_bytes, orbit   # the orbit's memory is borrowed by the handle

    def cpu_lav2(self, w, h, coords, n_iterations, row_step=1, col_step=1, threads=0, want_iters=False):
        """The reference's CPU renderer for HDRx32 + LAv2 (Fractal::CalcCpuPerturbationFractalLAV2<IterType, float,
        Disable>, Fractal.cpp:2485-2691; loop restated in ref_host_harness.cpp on the reference's own types and compiled
        methods) over a regular sub-grid of the w x h frame.  Returns (sum of iteration counts, iters or None)."""
        out = None
        if want_iters:
            out = np.zeros((h, w), np.uint32 if self._iter_bytes == 4 else np.uint64)
        buf = lambda b: C.cast(C.create_string_buffer(b, len(b)), C.c_void_p)
        total = lib().refhost_cpu_lav2(self._h, self._iter_bytes, w, h, buf(coords["dx"]), buf(coords["dy"]),
                                       buf(coords["center_x"]), buf(coords["center_y"]), n_iterations,
                                       out.ctypes.data if out is not None else None, row_step, col_step, threads)
        return int(total), out

    def close(self):
        if getattr(self, "_h", None):
            lib().refhost_free(self._h, self._iter_bytes)
            self._h = None

    __del__ = close
