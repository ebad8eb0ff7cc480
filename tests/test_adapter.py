"""The adapter header include/fs_gpu_adapter.hpp -- the reference's `GPURenderer` class on top of the C-ABI -- as code:
compiled against the reference's own headers with every member template instantiated (CPU), and driven from C++ with the
reference's own PerturbationResults / LAReference / BLAS objects on the GPU box (oracle/adapter_driver.cpp)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import cases
from fractalshark_b200 import RenderAlgorithm as A
from fractalshark_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "adapter_driver")
REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "FractalSharkLib")), reason="reference tree not mounted")
def test_adapter_header_compiles_against_the_reference_headers(built):
    """`make -C oracle _ref/adapter_driver`: g++ -std=c++23 on oracle/adapter_driver.cpp, whose instantiate_everything()
    calls every member of the adapter class with the argument types of the reference's explicit instantiations
    (GPU_Render.cu:227-1818) against GPU_Types.h / LAReference.h / BLAS.h / RenderAlgorithm.h where they lie."""
    subprocess.run(["touch", os.path.join(ROOT, "include", "fs_gpu_adapter.hpp")], check=True)
    res = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/adapter_driver"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert os.path.exists(DRIVER)


def _write_case(path, kind, w, h, n_iter, coords, orbit):
    with open(path, "wb") as f:
        count = orbit.count if orbit is not None else 0
        period = orbit.period if orbit is not None else 0
        f.write(struct.pack("<IIIQQQ", kind, w, h, n_iter, count, period))
        if kind == 3:
            f.write(coords["cx"] + coords["cy"] + coords["dx"] + coords["dy"])
        else:
            import ctypes as C
            radius = bytes((C.c_ubyte * 8).from_address(N.host_lib().fsh_orbit_max_radius(orbit._h)))
            f.write(coords["dx"] + coords["dy"] + coords["center_x"] + coords["center_y"] + radius)
            f.write(orbit.as_numpy().tobytes())


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(DRIVER), reason="oracle/_ref/adapter_driver not built (needs the reference tree)")
@pytest.mark.parametrize("name,kind", [("v5_hdr32_lav2", 1), ("v1_hdr32_lav2", 1), ("v5_hdr32_bla", 2), ("v0_gpu1x64", 3)])
def test_cpp_caller_renders_through_the_adapter(tmp_path, name, kind):
    """Fractal.cpp's call sequence in C++ (InitializeMemory, InitializePerturb with the reference's own LAReference,
    ClearMemory, RenderPerturbLAv2 / RenderPerturbBLA / Render, RenderCurrent) through the adapter class: the frame equals
    the committed fixture produced by the reference's kernels."""
    case = next(c for c in cases.ALL_SMALL_CASES if c[0] == name)
    _, view_id, w, h, alg, n_iter, ib = case
    _, coords, orbit, la, n = cases.make_inputs(view_id, w, h, alg, n_iter, ib)
    src, out = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    _write_case(src, kind, w, h, n, coords, orbit)
    res = subprocess.run([DRIVER, src, out], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = open(out, "rb").read()
    wp, hp, vmin, vmax, vsum = struct.unpack("<5Q", raw[:40])
    iters = np.frombuffer(raw[40:], dtype=np.uint32).reshape(hp, wp)
    want = cases.load_goldens()[name]
    np.testing.assert_array_equal(iters[:h, :w], want)
    assert (vmin, vmax, vsum) == (int(want.min()), int(want.max()), int(want.astype(np.uint64).sum()))
    assert not iters[h:, :].any() and not iters[:, w:].any()      # padding cells cleared, as the reference leaves them
