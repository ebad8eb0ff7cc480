"""Dev tool: where the host-side time of one end-to-end step goes (View 14, 3840x2160): per-call blocking time of the
public call sequence, pinned vs pageable inputs, result sink on/off."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fractalshark_b200 import RenderAlgorithm as A
from fractalshark_b200.gpu_renderer import GPURenderer
from fractalshark_b200.host_inputs import ReplicatedInputs
bench.set_view(14)
view, coords, orbit, la, n_iter, _ = bench.build_inputs(3840, 2160)
meta, blobs = ReplicatedInputs.pack(coords, orbit, la, n_iter)
pinned = []
for b in blobs:
    t = torch.empty(max(int(b.size), 1), dtype=torch.uint8, pin_memory=True)
    t[:b.size] = torch.from_numpy(b)
    pinned.append(t.numpy()[:b.size])
_, orbit_p, la_p, _ = ReplicatedInputs.unpack(meta, pinned)
r = GPURenderer(0)
assert r.InitializeMemory(3840, 2160, 1, iter_bytes=4) == 0
hp, wp = r.buffer_shape()
frame = torch.empty((hp, wp), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
alg = A.GpuHDRx32PerturbedLAv2
for label, o, l, sink in (("pageable inputs, copy after render", orbit, la, False), ("pinned inputs, copy after render", orbit_p, la_p, False),
                          ("pinned inputs, result sink", orbit_p, la_p, True)):
    assert r.SetResultSink(frame if sink else None) == 0
    acc = np.zeros(6)
    gen = 100
    reps = 12
    for rep in range(reps + 2):
        gen += 1
        t = [time.perf_counter()]
        assert r.InitializePerturb(gen, o, 0, None, l) == 0; t.append(time.perf_counter())
        r.ClearMemory(); t.append(time.perf_counter())
        assert r.RenderPerturbLAv2(alg, coords, n_iter) == 0; t.append(time.perf_counter())
        rc, it, _, rd = r.RenderCurrent(n_iter, iters_out=frame); t.append(time.perf_counter())
        if rep >= 2:
            acc[:4] += np.diff(t); acc[4] += t[-1] - t[0]; acc[5] += r.LastRenderMs() * 1e-3
    acc *= 1e3 / reps
    print(f"{label}: InitializePerturb {acc[0]:.3f}  ClearMemory {acc[1]:.3f}  RenderPerturbLAv2(call) {acc[2]:.3f}  "
          f"RenderCurrent(+sync) {acc[3]:.3f}  total {acc[4]:.3f} ms   kernel {acc[5]:.3f} ms", flush=True)
# the pieces on the device, one at a time (synchronised)
assert r.SetResultSink(None) == 0
for name, fn in (("InitializePerturb + sync", lambda g: r.InitializePerturb(g, orbit_p, 0, None, la_p)),
                 ("ClearMemory + sync", lambda g: r.ClearMemory()),
                 ("RenderCurrent (post + 24 B) + sync", lambda g: r.RenderCurrent(n_iter, want_iters=False))):
    best = 1e9
    for rep in range(5):
        gen += 1
        r.SyncComputeStream()
        t0 = time.perf_counter(); fn(gen); r.SyncComputeStream(); best = min(best, time.perf_counter() - t0)
    print(f"{name}: {best*1e3:.3f} ms", flush=True)
