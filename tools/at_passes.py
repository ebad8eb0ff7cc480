"""Dev probe (debug build): histogram of executed AT passes per pixel.  usage: python tools/at_passes.py VIEW"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fractalshark_b200 import RenderAlgorithm as A, traits, _native
from fractalshark_b200.gpu_renderer import GPURenderer
from fractalshark_b200.host_inputs import View, Orbit, LaTable
from fractalshark_b200.views import PRESETS

view_id = int(sys.argv[1]) if len(sys.argv) > 1 else 14
alg = A.GpuHDRx32PerturbedLAv2
W, H = 3840, 2160
p = PRESETS[view_id]
t = traits(alg)
v = View(p.min_x, p.min_y, p.max_x, p.max_y, W, H)
orbit = Orbit(v, t.numeric, p.num_iterations, True)
la = LaTable(orbit, 4)
coords = v.coords(t.numeric)
lib = C.CDLL(_native.GPU_LIB_PATH)
lib.fs_debug_at_passes.argtypes = [C.c_void_p]
buf = torch.zeros(2 * W * H + 64, dtype=torch.int32, device="cuda")
r = GPURenderer(0)
assert r.InitializeMemory(W, H, 1) == 0
assert r.InitializePerturb(1, orbit, 0, None, la) == 0
r.ClearMemory()
assert lib.fs_debug_at_passes(C.c_void_p(buf.data_ptr())) == 0
assert r.RenderPerturbLAv2(alg, coords, p.num_iterations) == 0
assert r.SyncComputeStream() == 0
a = buf.cpu().numpy().astype(np.int64).reshape(-1, 2)
a = a[a[:, 1] > 0]
ex, tot = a[:, 0], a[:, 1]
full = tot == tot.max()
print(f"pixels through AT: {len(a)}; reaching the pass limit {tot.max()}: {int(full.sum())}")
e = ex[full]
print("executed passes of those: quantiles 10/50/90/99/99.9/max:", [int(x) for x in np.quantile(e, [0.1, 0.5, 0.9, 0.99, 0.999, 1.0])])
print("  executed every pass (no cycle found):", int((e == tot.max()).sum()), " sum executed", int(e.sum()), "of", int(tot[full].sum()))
for lim in (256, 1024, 4096, 16384):
    print(f"  more than {lim} passes: {int((e > lim).sum())}")
print("escaping pixels: executed passes quantiles 50/90/99/max:", [int(x) for x in np.quantile(ex[~full], [0.5, 0.9, 0.99, 1.0])])
