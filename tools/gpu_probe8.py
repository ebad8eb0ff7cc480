"""Dev tool: per-shard kernel time of the View 14 frame on one GPU (shard imbalance + tail of the multi-GPU split)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
from fractalshark_b200 import RenderAlgorithm as A
from fractalshark_b200.gpu_renderer import GPURenderer
view_id = int(sys.argv[1]) if len(sys.argv) > 1 else 14
bench.set_view(view_id)
view, coords, orbit, la, n_iter, _ = bench.build_inputs(3840, 2160)
r = GPURenderer(0)
assert r.InitializeMemory(3840, 2160, 1, iter_bytes=4) == 0
assert r.InitializePerturb(1, orbit, 0, None, la) == 0
split = "--fused" not in sys.argv
assert r.SetSplitAt(split) == 0
print("split AT launch" if split else "fused launch", flush=True)
for world in (1, 2, 4, 8):
    ms = []
    for s in range(world):
        assert r.SetShard(world, s) == 0
        best = 1e9
        for rep in range(3):
            r.ClearMemory()
            assert r.RenderPerturbLAv2(A.GpuHDRx32PerturbedLAv2, coords, n_iter) == 0
            assert r.SyncComputeStream() == 0
            best = min(best, r.LastRenderMs())
        ms.append(best)
    print(f"world {world}: max {max(ms):.3f} ms  mean {sum(ms)/len(ms):.3f} ms  sum {sum(ms):.3f}  per-shard {[round(m,3) for m in ms]}", flush=True)
