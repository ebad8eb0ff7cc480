"""Generates tests/golden/ref_gpu_small.npz: iteration buffers produced by the REFERENCE's own CUDA kernels
(oracle/_ref/libref_gpurender.so, built from /root/reference for sm_100a by oracle/Makefile) for the
small parity cases of tests/cases.py.  Run on a B200 box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/ref_gpu_small.npz 1; python tests/golden/make_golden.py gpurun_out/ref_gpu_small2.npz 2; python tests/golden/make_golden.py gpurun_out/ref_gpu_small3.npz 3'

then copy the file to tests/golden/.  Inputs (orbit, LA table, coordinates) come from the in-tree
generator (libfshost.so) and are deterministic; their CRC32 is stored beside each buffer so a drift of
the generator is detected instead of silently invalidating the fixture.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
import ref_renderer  # noqa: E402


inputs_crc = cases.inputs_crc


def main(out_path, which="1"):
    out = {}
    for name, view_id, w, h, alg, n_iter, ib in cases.CASE_SETS[which]:
        _, coords, orbit, la, n = cases.make_inputs(view_id, w, h, alg, n_iter, ib)
        iters, _, red = cases.render(ref_renderer.RefGPURenderer, w, h, alg, coords, orbit, la, n, ib,
                                     precision=cases.CASE_PRECISION.get(name, 1))
        out[name] = iters[:h, :w].copy()
        out[name + "__crc"] = np.array([inputs_crc(coords, orbit, la)], dtype=np.uint64)
        print(name, iters[:h, :w].shape, red, flush=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path)


if __name__ == "__main__":
    # usage: make_golden.py OUT.npz [1|2]   (set k of cases.CASE_SETS -> cases.GOLDEN_FILES[k])
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_gpu_small.npz"),
         sys.argv[2] if len(sys.argv) > 2 else "1")
