"""Dev probe: lane-refill LAv2 kernel (fs_lav2_pool.cuh) against the one-tile-per-warp kernel on the same inputs:
exact comparison of the iteration buffers and device time of both.  usage: python tools/pool_ab.py [view:alg:niter ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fractalshark_b200 import RenderAlgorithm, traits
from fractalshark_b200.gpu_renderer import GPURenderer
from fractalshark_b200.host_inputs import View, Orbit, LaTable, BlaTable
from fractalshark_b200.views import PRESETS

W, H = 3840, 2160
A = RenderAlgorithm
_orbits = {}


def inputs(view_id, alg, n_iter, iter_bytes):
    global W, H
    p = PRESETS[view_id]
    t = traits(alg)
    v = View(p.min_x, p.min_y, p.max_x, p.max_y, W, H)
    key = (view_id, n_iter, iter_bytes, int(t.numeric), t.family)
    if key not in _orbits:
        orbit = Orbit(v, t.numeric, n_iter, True)
        _orbits[key] = (orbit, LaTable(orbit, iter_bytes) if t.family == "lav2" else BlaTable(orbit))
    return v.coords(t.numeric), *_orbits[key]


SWITCH = "pool"  # or "cycle"


def run(view_id, alg, n_iter=None, iter_bytes=4, reps=3, shard=None):
    n_iter = n_iter or PRESETS[view_id].num_iterations
    coords, orbit, la = inputs(view_id, alg, n_iter, iter_bytes)
    outs = {}
    for pool in (0, 1):
        r = GPURenderer(0)
        if SWITCH == "pool":
            r.SetPoolKernel(bool(pool))
        elif SWITCH == "la2":
            r.SetLaStep2(bool(pool))
        else:
            r.SetAtCycleDetection(bool(pool))
        assert r.InitializeMemory(W, H, 1, iter_bytes=iter_bytes) == 0
        if shard:
            r.SetShard(*shard)
        fam = traits(alg).family
        if fam == "lav2":
            assert r.InitializePerturb(1, orbit, 0, None, la) == 0
        best = 1e30
        for _ in range(reps):
            r.ClearMemory()
            rc = r.RenderPerturbLAv2(alg, coords, n_iter) if fam == "lav2" else r.RenderPerturbBLA(alg, orbit, la, coords, n_iter)
            assert rc == 0, rc
            assert r.SyncComputeStream() == 0
            best = min(best, r.LastRenderMs())
        rc, iters, _, red = r.RenderCurrent(n_iter)
        assert rc == 0, rc
        outs[pool] = (iters[:H, :W].copy(), best)
        r.close()
    same = np.array_equal(outs[0][0], outs[1][0])
    nd = int((outs[0][0] != outs[1][0]).sum())
    print(f"view {view_id} {alg.name} n={n_iter} u{iter_bytes*8} shard={shard}: off {outs[0][1]:.3f} ms  {SWITCH} {outs[1][1]:.3f} ms  "
          f"({outs[0][1]/outs[1][1]:.2f}x)  identical={same} differing={nd}", flush=True)
    return same


if __name__ == "__main__":
    if len(sys.argv) > 1:
        SWITCH = sys.argv[1]
    ok = True
    if SWITCH == "blafast":
        SWITCH = "la2"
        ok &= run(5, A.GpuHDRx32PerturbedBLA, reps=2)
        W, H = 1920, 1080
        _orbits.clear()
        ok &= run(14, A.GpuHDRx32PerturbedBLA, reps=2)
        ok &= run(1, A.GpuHDRx32PerturbedBLA, reps=2)
        ok &= run(19, A.GpuHDRx32PerturbedBLA, 3000000, reps=2)
        ok &= run(100, A.GpuHDRx32PerturbedBLA, reps=2)
        ok &= run(5, A.GpuHDRx32PerturbedBLA, iter_bytes=8, reps=2)
        print("ALL IDENTICAL" if ok else "MISMATCH")
        sys.exit(0 if ok else 1)
    if SWITCH == "bla":
        SWITCH = "cycle"
        W, H = 1920, 1080
        ok &= run(14, A.GpuHDRx32PerturbedBLA, reps=2)
        ok &= run(5, A.GpuHDRx32PerturbedBLA, reps=2)
        ok &= run(1, A.GpuHDRx32PerturbedBLA, reps=2)
        ok &= run(14, A.GpuHDRx64PerturbedBLA, reps=2)
        ok &= run(100, A.Gpu1x64PerturbedBLA, reps=2)
        ok &= run(5, A.GpuHDRx32PerturbedBLA, iter_bytes=8, reps=2)
        print("ALL IDENTICAL" if ok else "MISMATCH")
        sys.exit(0 if ok else 1)
    if SWITCH == "cycle":
        W, H = 1920, 1080
        ok &= run(5, A.GpuHDRx32PerturbedLAv2)
        ok &= run(19, A.GpuHDRx32PerturbedLAv2)
        ok &= run(5, A.GpuHDRx32PerturbedLAv2PO, 50000)
        ok &= run(1, A.GpuHDRx32PerturbedLAv2)
        ok &= run(5, A.GpuHDRx32PerturbedLAv2, iter_bytes=8)
        ok &= run(14, A.GpuHDRx2x32PerturbedLAv2)
        ok &= run(14, A.GpuHDRx64PerturbedLAv2)
        ok &= run(14, A.GpuHDRx32PerturbedLAv2)
        ok &= run(19, A.GpuHDRx2x32PerturbedLAv2)
        ok &= run(100, A.Gpu1x64PerturbedLAv2)
        ok &= run(101, A.Gpu1x32PerturbedLAv2)
        ok &= run(101, A.Gpu2x32PerturbedLAv2)
        ok &= run(100, A.GpuHDRx64PerturbedLAv2)
        ok &= run(5, A.GpuHDRx2x32PerturbedLAv2)
        print("ALL IDENTICAL" if ok else "MISMATCH")
        sys.exit(0 if ok else 1)
    ok &= run(14, A.GpuHDRx32PerturbedLAv2)
    ok &= run(5, A.GpuHDRx32PerturbedLAv2)
    ok &= run(19, A.GpuHDRx32PerturbedLAv2)
    ok &= run(5, A.GpuHDRx32PerturbedLAv2PO, 20000)
    ok &= run(5, A.GpuHDRx32PerturbedLAv2LAO)
    ok &= run(1, A.GpuHDRx32PerturbedLAv2)
    ok &= run(5, A.GpuHDRx32PerturbedLAv2, iter_bytes=8)
    ok &= run(14, A.GpuHDRx64PerturbedLAv2)
    ok &= run(14, A.GpuHDRx32PerturbedLAv2, iter_bytes=8)
    ok &= run(14, A.GpuHDRx32PerturbedLAv2, shard=(8, 3))
    print("ALL IDENTICAL" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
