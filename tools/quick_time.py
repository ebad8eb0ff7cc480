"""Dev probe: device time of a few HDRx32 LAv2 frames with the library FS_GPU_LIB points at (A/B of builds), plus a CRC of
every iteration buffer so that builds can be compared for identical output.  usage: python tools/quick_time.py [tag]"""
import sys, os, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fractalshark_b200 import RenderAlgorithm as A, traits
from fractalshark_b200.gpu_renderer import GPURenderer
from fractalshark_b200.host_inputs import View, Orbit, LaTable
from fractalshark_b200.views import PRESETS

W, H = 3840, 2160
tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("FS_GPU_LIB", "default"))
views = [int(v) for v in os.environ.get("QT_VIEWS", "14,19,5").split(",")]
out = []
for view_id in views:
    alg = A.GpuHDRx32PerturbedLAv2
    p = PRESETS[view_id]
    t = traits(alg)
    v = View(p.min_x, p.min_y, p.max_x, p.max_y, W, H)
    orbit = Orbit(v, t.numeric, p.num_iterations, True)
    la = LaTable(orbit, 4)
    coords = v.coords(t.numeric)
    for shard in (None, (8, 3)) if view_id == 14 else (None,):
        r = GPURenderer(0)
        assert r.InitializeMemory(W, H, 1) == 0
        if shard:
            r.SetShard(*shard)
        assert r.InitializePerturb(1, orbit, 0, None, la) == 0
        ts = []
        for _ in range(5):
            r.ClearMemory()
            assert r.RenderPerturbLAv2(alg, coords, p.num_iterations) == 0
            assert r.SyncComputeStream() == 0
            ts.append(r.LastRenderMs())
        rc, iters, _, red = r.RenderCurrent(p.num_iterations)
        crc = zlib.crc32(np.ascontiguousarray(iters[:H, :W]).tobytes())
        out.append(f"v{view_id}{'s' if shard else ''} {min(ts):.3f} ms crc {crc:08x}")
        r.close()
print(f"[{tag}] " + " | ".join(out), flush=True)
