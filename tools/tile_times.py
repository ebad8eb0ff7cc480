"""Dev probe (debug build: make -C fractalshark_b200/csrc dbg; FS_GPU_LIB=.../libfsgpu_dbg.so): when every 8x4 tile of a
LAv2 launch started and how long it took.  usage: python tools/tile_times.py VIEW [SHARDS INDEX]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fractalshark_b200 import RenderAlgorithm as A, traits, _native
from fractalshark_b200.gpu_renderer import GPURenderer
from fractalshark_b200.host_inputs import View, Orbit, LaTable
from fractalshark_b200.views import PRESETS

view_id = int(sys.argv[1]) if len(sys.argv) > 1 else 14
shard = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1, 0)
alg = A.GpuHDRx32PerturbedLAv2
W, H = 3840, 2160
p = PRESETS[view_id]
t = traits(alg)
v = View(p.min_x, p.min_y, p.max_x, p.max_y, W, H)
orbit = Orbit(v, t.numeric, p.num_iterations, True)
la = LaTable(orbit, 4)
coords = v.coords(t.numeric)
lib = C.CDLL(_native.GPU_LIB_PATH)
lib.fs_debug_tile_times.argtypes = [C.c_void_p]
n_tiles = ((W + 7) // 8) * ((H + 3) // 4)
buf = torch.zeros(2 * n_tiles, dtype=torch.int64, device="cuda")
r = GPURenderer(0)
assert r.InitializeMemory(W, H, 1) == 0
r.SetShard(*shard)
assert r.InitializePerturb(1, orbit, 0, None, la) == 0
for rep in range(3):
    r.ClearMemory()
    buf.zero_()
    assert lib.fs_debug_tile_times(C.c_void_p(buf.data_ptr())) == 0
    assert r.RenderPerturbLAv2(alg, coords, p.num_iterations) == 0
    assert r.SyncComputeStream() == 0
    ms = r.LastRenderMs()
tt = buf.cpu().numpy().reshape(-1, 2)
tt = tt[tt[:, 1] > 0]
start = (tt[:, 0] - tt[:, 0].min()) / 1e6          # ms
dur = tt[:, 1] / 1.965e6                            # ms at 1,965 MHz
end = start + dur
print(f"view {view_id} shard {shard}: {ms:.3f} ms, {len(tt)} tiles; sum of tile durations {dur.sum():.1f} ms "
      f"(/{148*32} warps = {dur.sum()/(148*32):.3f} ms)")
q = np.quantile(dur, [0.5, 0.9, 0.99, 0.999, 1.0])
print("tile duration ms: median %.4f  p90 %.4f  p99 %.4f  p99.9 %.4f  max %.4f" % tuple(q))
order = np.argsort(-dur)[:10]
for i in order:
    print(f"  slow tile: start {start[i]:.3f} ms  duration {dur[i]:.3f} ms  end {end[i]:.3f}")
late = np.argsort(-end)[:5]
for i in late:
    print(f"  last to finish: start {start[i]:.3f} ms  duration {dur[i]:.3f} ms  end {end[i]:.3f}")
for thr in (0.05, 0.1, 0.2):
    m = dur > thr
    print(f"tiles longer than {thr} ms: {int(m.sum())}, their start times: median {np.median(start[m]) if m.any() else 0:.3f} max {start[m].max() if m.any() else 0:.3f}")
