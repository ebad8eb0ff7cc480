"""Dev probe (debug build: make -C fractalshark_b200/csrc dbg; FS_GPU_LIB=fractalshark_b200/libfsgpu_dbg.so):
session counters of the lane-refill kernel on one view."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fractalshark_b200 import RenderAlgorithm as A, traits, _native
from fractalshark_b200.gpu_renderer import GPURenderer
from fractalshark_b200.host_inputs import View, Orbit, LaTable
from fractalshark_b200.views import PRESETS

view_id = int(sys.argv[1]) if len(sys.argv) > 1 else 14
alg = getattr(A, sys.argv[2]) if len(sys.argv) > 2 else A.GpuHDRx32PerturbedLAv2
W, H = 3840, 2160
p = PRESETS[view_id]
n_iter = int(sys.argv[3]) if len(sys.argv) > 3 else p.num_iterations
t = traits(alg)
v = View(p.min_x, p.min_y, p.max_x, p.max_y, W, H)
orbit = Orbit(v, t.numeric, n_iter, True)
la = LaTable(orbit, 4)
coords = v.coords(t.numeric)
lib = C.CDLL(_native.GPU_LIB_PATH)
r = GPURenderer(0)
assert r.InitializeMemory(W, H, 1) == 0
assert r.InitializePerturb(1, orbit, 0, None, la) == 0
out = (C.c_uint64 * 16)()
for rep in range(2):
    r.ClearMemory()
    lib.fs_debug_pool_counters(out)
    assert r.RenderPerturbLAv2(alg, coords, n_iter) == 0
    assert r.SyncComputeStream() == 0
    ms = r.LastRenderMs()
lib.fs_debug_pool_counters(out)
d = [int(x) for x in out]
print(f"view {view_id} {alg.name}: {ms:.3f} ms")
print(f"LA  sessions {d[8]}  trips {d[0]} (lanes {d[1]/max(d[0],1):.1f})  drain trips {d[2]} (lanes {d[3]/max(d[2],1):.1f})")
print(f"PO  sessions {d[9]}  rounds {d[4]} (lanes {d[5]/max(d[4],1):.1f})  drain rounds {d[6]} (lanes {d[7]/max(d[6],1):.1f})")
tot = max(d[15], 1)
print(f"warp-cycles: total {tot/1e9:.2f} G  LA {d[11]/tot*100:.1f} % + drain {d[12]/tot*100:.1f} %   PO {d[13]/tot*100:.1f} % + drain {d[14]/tot*100:.1f} %")
