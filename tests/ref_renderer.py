"""TEST INFRASTRUCTURE: ctypes wrapper of oracle/_ref/libref_gpurender.so (the reference's own CUDA
kernels + GPURenderer built for sm_100a, see oracle/ref_harness.cu).  Same call shape as
fractalshark_b200.gpu_renderer.GPURenderer so parity tests feed both with identical inputs."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from fractalshark_b200.algorithms import RenderAlgorithm, traits
from fractalshark_b200.gpu_renderer import NB_THREADS_H, NB_THREADS_W, _round_up, default_palette

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_gpurender.so")


def available() -> bool:
    return os.path.exists(REF_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_LIB)
        V, U32, I32, U64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64
        L.refh_create.restype = V
        L.refh_destroy.argtypes = [V]
        L.refh_init_memory.argtypes = [V, U32, U32, U32, U32, V, U32, U32, U64, I32]
        L.refh_init_perturb.argtypes = [V, U32, I32, U64, V, U64, U64, V, V, V, U64, V, U64, V, U64, I32, I32]
        L.refh_clear.argtypes = [V, U32]
        L.refh_render_lav2.argtypes = [V, U32, U32, I32, I32, V, V, V, V, V, V, U64]
        L.refh_render_direct.argtypes = [V, U32, U32, I32, V, V, V, V, U64, I32]
        L.refh_render_bla.argtypes = [V, U32, U32, I32, V, U64, U64, V, V, U32, I32, V, V, V, V, V, V, U64]
        L.refh_render_bla.restype = U32
        L.refh_render_scaled.argtypes = [V, U32, U32, I32, V, V, U64, U64, V, V, V, V, V, V, U64]
        L.refh_render_scaled.restype = U32
        L.refh_init_perturb_rc.argtypes = [V, U32, I32, U64, V, U64, U64, U64, V, V, V, U64, V, U64, V, U64, I32, I32]
        L.refh_init_perturb_rc.restype = U32
        L.refh_render_lav2_rc.argtypes = [V, U32, U32, I32, I32, V, V, V, V, V, V, U64]
        L.refh_render_lav2_rc.restype = U32
        L.refh_render_current.argtypes = [V, U32, U64, V, V, V]
        L.refh_sync.argtypes = [V]
        L.refh_last_render_ms.argtypes = [V, C.POINTER(C.c_float)]
        for f in ("refh_init_memory", "refh_init_perturb", "refh_render_lav2", "refh_render_direct",
                  "refh_render_current", "refh_sync", "refh_last_render_ms", "refh_test_cuda"):
            getattr(L, f).restype = U32
        _lib = L
    return _lib


def _buf(b: bytes):
    return C.cast(C.create_string_buffer(b, len(b)), C.c_void_p)


class RefGPURenderer:
    def __init__(self):
        self._lib = lib()
        self._h = self._lib.refh_create()
        self._iter_bytes = 4

    def close(self):
        if getattr(self, "_h", None):
            self._lib.refh_destroy(self._h)
            self._h = None

    __del__ = close

    def InitializeMemory(self, w, h, antialiasing=1, palette=None, palette_aux_depth=0, palette_generation=1,
                         expected_reuse=False, iter_bytes=4):
        if palette is None:
            palette = default_palette()
        self._palette = np.ascontiguousarray(palette, dtype=np.uint16).reshape(-1, 4)
        self._iter_bytes, self._w, self._h_px, self._aa = iter_bytes, w, h, antialiasing
        return int(self._lib.refh_init_memory(self._h, iter_bytes, w, h, antialiasing, self._palette.ctypes.data,
                                              self._palette.shape[0], palette_aux_depth, palette_generation,
                                              int(expected_reuse)))

    def InitializePerturb(self, generation1, perturb1, generation2=0, perturb2=None, la=None, pextras=0):
        d = perturb1.descriptor()
        if la is not None:
            l = la.descriptor()
            args = (l.las, l.num_las, l.stages, l.num_stages, l.at, l.la_stage_count, l.use_at, l.is_valid)
        else:
            args = (None, 0, None, 0, None, 0, 0, 0)
        if getattr(perturb1, "pextras", 0) == 2:
            return int(self._lib.refh_init_perturb_rc(self._h, self._iter_bytes, int(perturb1.numeric), generation1,
                                                      d.elements, d.compressed_count, d.uncompressed_count,
                                                      d.period_maybe_zero, d.orbit_x_low, d.orbit_y_low, *args))
        return int(self._lib.refh_init_perturb(self._h, self._iter_bytes, int(perturb1.numeric), generation1,
                                               d.elements, d.uncompressed_count, d.period_maybe_zero, d.orbit_x_low,
                                               d.orbit_y_low, *args))

    def ClearMemory(self):
        self._lib.refh_clear(self._h, self._iter_bytes)

    def Render(self, algorithm, coords, n_iterations, iteration_precision=1):
        t = traits(algorithm)
        return int(self._lib.refh_render_direct(self._h, self._iter_bytes, int(algorithm), int(t.numeric),
                                                _buf(coords["cx"]), _buf(coords["cy"]), _buf(coords["dx"]),
                                                _buf(coords["dy"]), n_iterations, iteration_precision))

    def RenderPerturbLAv2(self, algorithm, coords, n_iterations):
        t = traits(algorithm)
        fn = self._lib.refh_render_lav2_rc if int(t.pextras) == 2 else self._lib.refh_render_lav2
        return int(fn(self._h, self._iter_bytes, int(algorithm), int(t.numeric), int(t.mode),
                                              _buf(coords["cx"]), _buf(coords["cy"]), _buf(coords["dx"]),
                                              _buf(coords["dy"]), _buf(coords["center_x"]), _buf(coords["center_y"]),
                                              n_iterations))

    def RenderPerturbBLA(self, algorithm, perturb, blas, coords, n_iterations, iteration_precision=1):
        t = traits(algorithm)
        d, b = perturb.descriptor(), blas.descriptor()
        return int(self._lib.refh_render_bla(self._h, self._iter_bytes, int(algorithm), int(t.numeric), d.elements,
                                             d.uncompressed_count, d.period_maybe_zero, b.levels, b.level_counts,
                                             b.num_levels, b.lm2, _buf(coords["cx"]), _buf(coords["cy"]),
                                             _buf(coords["dx"]), _buf(coords["dy"]), _buf(coords["center_x"]),
                                             _buf(coords["center_y"]), n_iterations))

    def RenderPerturbBLAScaled(self, algorithm, double_perturb, float_perturb, coords, n_iterations,
                               iteration_precision=1):
        t = traits(algorithm)
        d, f = double_perturb.descriptor(), float_perturb.descriptor()
        return int(self._lib.refh_render_scaled(self._h, self._iter_bytes, int(algorithm), int(t.numeric), d.elements,
                                                f.elements, d.uncompressed_count, d.period_maybe_zero,
                                                _buf(coords["cx"]), _buf(coords["cy"]), _buf(coords["dx"]),
                                                _buf(coords["dy"]), _buf(coords["center_x"]),
                                                _buf(coords["center_y"]), n_iterations))

    def RenderCurrent(self, n_iterations, want_iters=True, want_colors=False, progressive=False):
        hp, wp = _round_up(self._h_px, NB_THREADS_H), _round_up(self._w, NB_THREADS_W)
        dt = np.uint32 if self._iter_bytes == 4 else np.uint64
        iters = np.empty((hp, wp), dtype=dt) if want_iters else None
        colors = None
        if want_colors:
            colors = np.empty((_round_up(self._h_px // self._aa, NB_THREADS_H),
                               _round_up(self._w // self._aa, NB_THREADS_W), 4), dtype=np.uint16)
        red = (C.c_uint64 * 3)()
        rc = int(self._lib.refh_render_current(self._h, self._iter_bytes, n_iterations,
                                               iters.ctypes.data if want_iters else None,
                                               colors.ctypes.data if want_colors else None, red))
        return rc, iters, colors, {"Min": int(red[0]), "Max": int(red[1]), "Sum": int(red[2])}

    def SyncComputeStream(self):
        return int(self._lib.refh_sync(self._h))

    def LastRenderMs(self):
        ms = C.c_float(0)
        rc = self._lib.refh_last_render_ms(self._h, C.byref(ms))
        if rc:
            raise RuntimeError(f"cuda error {rc}")
        return float(ms.value)
