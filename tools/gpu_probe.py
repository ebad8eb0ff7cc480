"""Ad-hoc GPU probe: run new kernels and the reference's kernels on the same inputs, print parity + device time."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from fractalshark_b200 import RenderAlgorithm, Numeric, traits
from fractalshark_b200.gpu_renderer import GPURenderer
from fractalshark_b200.host_inputs import View, Orbit, LaTable, BlaTable
from fractalshark_b200.views import PRESETS
import ref_renderer


def compare(a, b, w, h, tag):
    a = a[:h, :w].astype(np.int64); b = b[:h, :w].astype(np.int64)
    d = np.abs(a - b)
    exact = float((d == 0).mean())
    print(f"  {tag}: exact={exact*100:.4f}%  max|d|={int(d.max())}  >1: {int((d>1).sum())}  sum_new={int(a.sum())} sum_ref={int(b.sum())}", flush=True)
    return exact, int(d.max())


def run(view_id, w, h, alg, n_iter=None, iter_bytes=4, with_ref=True):
    p = PRESETS[view_id]
    n_iter = n_iter or p.num_iterations
    t = traits(alg)
    v = View(p.min_x, p.min_y, p.max_x, p.max_y, w, h)
    coords = v.coords(t.numeric, direct=(t.family == 'direct'))
    print(f"view {view_id} {w}x{h} {alg.name} n_iter={n_iter} iter_bytes={iter_bytes}", flush=True)
    orbit = la = None
    if t.family == "lav2":
        t0 = time.time(); orbit = Orbit(v, t.numeric, n_iter, True); t1 = time.time()
        if int(t.pextras) == 2:
            orbit = orbit.compress()
            print(f"  compressed: {orbit.count} waypoints for {orbit.uncompressed_count} entries", flush=True)
        la = LaTable(orbit, iter_bytes)
        print(f"  orbit count={orbit.count} period={orbit.period} ({t1-t0:.2f}s) la: n={la.num_las} stages={la.stage_count} at={la.use_at} valid={la.is_valid} ({time.time()-t1:.2f}s)", flush=True)
    if t.family == "bla":
        t0 = time.time(); orbit = Orbit(v, t.numeric, n_iter, True); t1 = time.time()
        la = BlaTable(orbit)
        print(f"  orbit count={orbit.count} period={orbit.period} ({t1-t0:.2f}s) bla: levels={la.num_levels} lm2={la.lm2} ({time.time()-t1:.2f}s)", flush=True)
    if t.family == "scaled":
        base = Orbit(v, t.numeric, n_iter, True)
        orbit, la = base.with_bad(), base.with_bad(True)
        nbad = int(orbit.as_numpy()[:, 0].sum())
        print(f"  orbit count={orbit.count} period={orbit.period} bad={nbad}", flush=True)
    outs = {}
    for name, R in (("new", GPURenderer), ("ref", ref_renderer.RefGPURenderer)):
        if name == "ref" and not with_ref:
            continue
        r = R()
        rc = r.InitializeMemory(w, h, 1, iter_bytes=iter_bytes); assert rc == 0, rc
        if orbit is not None and t.family == "lav2":
            rc = r.InitializePerturb(1, orbit, 0, None, la); assert rc == 0, rc
        for rep in range(2):
            r.ClearMemory()
            if t.family == "lav2":
                rc = r.RenderPerturbLAv2(alg, coords, n_iter)
            elif t.family == "bla":
                rc = r.RenderPerturbBLA(alg, orbit, la, coords, n_iter)
            elif t.family == "scaled":
                rc = r.RenderPerturbBLAScaled(alg, orbit, la, coords, n_iter)
            else:
                rc = r.Render(alg, coords, n_iter, 1)
            assert rc == 0, rc
            rc = r.SyncComputeStream(); assert rc == 0, rc
            ms = r.LastRenderMs()
        rc, iters, _, red = r.RenderCurrent(n_iter); assert rc == 0, rc
        print(f"  {name}: {ms:.3f} ms  sum={red['Sum']} min={red['Min']} max={red['Max']}  -> {red['Sum']/ms/1e6:.3f} G pixel-iters/s", flush=True)
        outs[name] = (iters, ms, red)
        r.close()
    if "ref" in outs:
        compare(outs["new"][0], outs["ref"][0], w, h, "parity")
        print(f"  speedup vs reference kernel: {outs['ref'][1]/outs['new'][1]:.2f}x", flush=True)
    return outs


if __name__ == "__main__":
    A = RenderAlgorithm
    small = "--small" in sys.argv
    W, H = (512, 288) if small else (3840, 2160)
    run(0, W, H, A.Gpu1x64, 2048 if small else 65536)
    run(0, W, H, A.Gpu1x32, 2048 if small else 65536)
    run(5, W, H, A.GpuHDRx32PerturbedLAv2PO, 20000)
    run(5, W, H, A.GpuHDRx32PerturbedLAv2)
    run(5, W, H, A.GpuHDRx32PerturbedLAv2LAO)
    run(5, W, H, A.GpuHDRx32PerturbedLAv2, iter_bytes=8)
    run(5, W, H, A.GpuHDRx64PerturbedLAv2)
    run(1, W, H, A.GpuHDRx32PerturbedLAv2)
