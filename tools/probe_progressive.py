"""Dev tool: latency of a progressive RenderCurrent issued while a long render runs (View 5 PO)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cases
from fractalshark_b200 import RenderAlgorithm as A
from fractalshark_b200.gpu_renderer import GPURenderer

w, h, alg, n_iter = 3840, 2160, A.GpuHDRx32PerturbedLAv2PO, 250000
_, coords, orbit, la, n = cases.make_inputs(5, w, h, alg, n_iter, 4)
r = GPURenderer()
assert r.InitializeMemory(w, h, 1, iter_bytes=4) == 0
assert r.InitializePerturb(1, orbit, 0, None, la) == 0
r.ClearMemory(); r.RenderPerturbLAv2(alg, coords, n); r.SyncComputeStream()
full = r.LastRenderMs()
print("full render ms", full, flush=True)
import ctypes as C
rt = C.CDLL("libcudart.so.12")
stream = C.c_void_p()
lo, hi = C.c_int(), C.c_int()
rt.cudaDeviceGetStreamPriorityRange(C.byref(lo), C.byref(hi))
assert rt.cudaStreamCreateWithPriority(C.byref(stream), 1, hi.value) == 0
dbuf = C.c_void_p()
assert rt.cudaMalloc(C.byref(dbuf), 1 << 26) == 0
for delay in (0.05, 0.3):
    r.ClearMemory(); r.SyncComputeStream()
    t0 = time.perf_counter()
    r.RenderPerturbLAv2(alg, coords, n)
    time.sleep(delay * full * 1e-3)
    t1 = time.perf_counter()
    rt.cudaMemsetAsync(dbuf, 1, 1 << 26, stream)      # a memset kernel on an unrelated high-priority stream
    rt.cudaStreamSynchronize(stream)
    t2 = time.perf_counter()
    r.SyncComputeStream()
    t3 = time.perf_counter()
    print(f"raw 64 MiB memset on own stream: asked at {1e3*(t1-t0):7.1f} ms  back after {1e3*(t2-t1):7.1f} ms  render done at {1e3*(t3-t0):7.1f} ms", flush=True)
for label, kw in (("reduction only", dict(want_iters=False, want_colors=False)),
                  ("iters", dict(want_iters=True, want_colors=False)),
                  ("iters+colors", dict(want_iters=True, want_colors=True))):
    for delay in (0.05, 0.3):
        r.ClearMemory(); r.SyncComputeStream()
        t0 = time.perf_counter()
        r.RenderPerturbLAv2(alg, coords, n)
        time.sleep(delay * full * 1e-3)
        t1 = time.perf_counter()
        rc, it, col, red = r.RenderCurrent(n, progressive=True, **kw)
        t2 = time.perf_counter()
        running = r.QueryComputeStream()
        r.SyncComputeStream()
        t3 = time.perf_counter()
        print(f"{label:16s} asked at {1e3*(t1-t0):7.1f} ms  back after {1e3*(t2-t1):7.1f} ms  compute query={running}  "
              f"render done at {1e3*(t3-t0):7.1f} ms  partial sum={red['Sum']}", flush=True)
