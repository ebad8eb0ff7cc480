#!/usr/bin/env python
"""bench.py -- headline benchmark of the per-pixel render path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (libfsgpu.so on N B200s)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU port on host cores

Workload (config.workload): View #14 (default; north_star's target view: 6,632-digit coordinates, zoom 4.7e6516,
maxIter 2,147,483,646) or View #5 (`--view 5`, maxIter 4,718,592), GpuHDRx32PerturbedLAv2, 3840x2160, AA 1,
u32 iterations (BASELINE.json metric "View 5/14 perturb+LA", configs[2]/[3]).  The reference orbit comes from
the in-tree GMP loop (View 14: 21.7 kbit, period 116,695, a few seconds; untimed).  One step = ClearMemory + one
full render of the frame.

* value       pixel-iterations/s = ReductionResults.Sum / device time of the render kernel(s), inputs
              (orbit, LA table) already resident in HBM; CUDA events on the launching stream, max over ranks.
* e2e         same metric through the public C-ABI call sequence with HOST buffers inside the timed region:
              InitializePerturb (H2D orbit + LA table), ClearMemory, RenderPerturbLAv2, RenderCurrent
              (AA/palette/reduction + D2H of the iteration buffer and the 24-byte reduction).
* roofline    FP32 issue: executed steps by kind (AT passes, LA steps, perturbation steps; device counters) x FP32
              instructions per step (9 for an AT pass) / kernel time, against the FFMA issue
              peak measured live by a micro-kernel on the same GPU.
* cpu_baseline the oracle's CPU port of the same kernel on a bounded pixel sample (all host threads).
N > 1: 4-row tile bands are dealt round-robin to ranks (no data-path collective inside the render);
the orbit/LA blob is replicated by an NCCL broadcast.  The device-timed arm checks the merged frame with an NCCL
reduce (disjoint rows, so SUM == gather; untimed).  The e2e arm assembles the frame on the HOST: one shared-memory
frame, page-locked by every rank, into which each rank copies only the bands it rendered (fs_render_current_shard).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 3840, 2160
VIEW_ID = 14
WORKLOAD = METRIC = ""
FP32_INSTR_PER_PERTURB_STEP = 20  # HDRx32 perturbation step, mantissa ops only
FP32_INSTR_PER_LA_STEP = 22
FP32_INSTR_PER_AT_PASS = 9       # z <- z^2 + c with |z|^2 test: rr, ii, rr+ii, rr-ii, re*im, 3 FMA + compare (ATInfo.h:155-188)


def set_view(view_id):
    global VIEW_ID, WORKLOAD, METRIC
    from fractalshark_b200.views import PRESETS
    VIEW_ID = view_id
    WORKLOAD = f"view{view_id}_GpuHDRx32PerturbedLAv2_{WIDTH}x{HEIGHT}_aa1_u32_maxiter{PRESETS[view_id].num_iterations}"
    METRIC = f"pixel-iters/sec (device-timed) for View {view_id} perturb+LA"


def _clock_sampler(stop, samples, gpu_index):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            out = subprocess.run(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip().splitlines()
            if out:
                samples.append([x.strip() for x in out[0].split(",")])
        except Exception:
            pass
        stop.wait(0.2)


def _clock_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None, "reasons": reasons,
            "samples": len(samples)}


def build_inputs(width, height):
    from fractalshark_b200 import Numeric
    from fractalshark_b200.host_inputs import LaTable, Orbit, View
    from fractalshark_b200.views import PRESETS
    p = PRESETS[VIEW_ID]
    view = View(p.min_x, p.min_y, p.max_x, p.max_y, width, height)
    t0 = time.time()
    orbit = Orbit(view, Numeric.HDR32, p.num_iterations, True)
    t1 = time.time()
    la = LaTable(orbit, 4)
    t2 = time.time()
    return view, view.coords(Numeric.HDR32), orbit, la, p.num_iterations, {"orbit_s": t1 - t0, "la_s": t2 - t1}


def cpu_port_sample(coords, orbit, la, n_iter, threads, stride=None):
    """Oracle CPU port (the checker, timed as a baseline only) on a regular sub-grid of the frame."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_cpu
    from fractalshark_b200 import RenderAlgorithm
    if stride is None:
        stride = 6 if VIEW_ID == 5 else 1  # ~10-30 s of CPU work on 16 host threads
    row_step = col_step = stride
    t0 = time.time()
    iters, steps = oracle_cpu.render_lav2(RenderAlgorithm.GpuHDRx32PerturbedLAv2, WIDTH, HEIGHT, coords, orbit, la,
                                          n_iter, rows=(0, HEIGHT), col_step=col_step, row_step=row_step,
                                          threads=threads)
    dt = time.time() - t0
    total = int(iters[0:HEIGHT:row_step, 0:WIDTH:col_step].sum())
    what = "the whole frame" if row_step == 1 and col_step == 1 else \
        f"a regular sub-grid of the frame: 1 of every {row_step} rows x 1 of every {col_step} columns"
    return total / dt, dt, f"{what} ({(HEIGHT // row_step) * (WIDTH // col_step)} pixels, {steps} executed steps, {dt:.1f} s)"


def run_reference_arm(args, rank, world):
    """Reference arm of this tier: the CPU port of the path on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_cpu
    _, coords, orbit, la, n_iter, _ = build_inputs(WIDTH, HEIGHT)
    threads = oracle_cpu.hardware_threads()
    vals, times, sample = [], [], ""
    # bounded sample: a coarse probe pass sizes the sub-grid so that warmup + steps passes end within ~4 minutes
    _, probe_dt, _ = cpu_port_sample(coords, orbit, la, n_iter, threads, stride=16)
    budget = 240.0 / max(args.warmup + args.steps, 1)
    stride = next((k for k in (1, 2, 3, 4, 6, 8, 12) if probe_dt * 256.0 / (k * k) <= budget), 16)
    for i in range(args.warmup + args.steps):
        v, dt, sample = cpu_port_sample(coords, orbit, la, n_iter, threads, stride=stride)
        if i >= args.warmup:
            vals.append(v)
            times.append(dt)
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pixel-iters/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+i32 (HDRx32)",
            "data": f"synthetic (View #{VIEW_ID} preset coordinates, orbit + LA table generated in-process)",
            "config": {"workload": WORKLOAD, "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": "pixel-iters/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "pixel-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--view", type=int, default=14, choices=[5, 14], help="view preset of the workload")
    args = ap.parse_args()
    set_view(args.view)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from fractalshark_b200 import RenderAlgorithm
    from fractalshark_b200.gpu_renderer import GPURenderer

    if not torch.cuda.is_available() or not GPURenderer.TestCudaIsWorking():
        raise SystemExit("bench.py: no CUDA device -- the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    alg = RenderAlgorithm.GpuHDRx32PerturbedLAv2
    from fractalshark_b200.host_inputs import ReplicatedInputs

    # ---- inputs: produced once on rank 0 (GMP orbit, LA table), replicated to the other ranks with NCCL broadcasts
    # of the packed blobs (north_star); receivers upload straight from the received host buffers -------------------
    bcast_ms = None
    if rank == 0:
        view, coords, orbit, la, n_iter, gen_times = build_inputs(WIDTH, HEIGHT)
    if world > 1:
        meta_box = [None]
        if rank == 0:
            meta, blobs = ReplicatedInputs.pack(coords, orbit, la, n_iter)
            meta["gen_times"] = gen_times
            meta_box = [meta]
        dist.broadcast_object_list(meta_box, src=0)
        meta = meta_box[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dev = []
        for i, size in enumerate(meta["sizes"]):
            t = torch.from_numpy(blobs[i]).cuda() if rank == 0 else torch.empty(size, dtype=torch.uint8, device="cuda")
            dev.append(t)
        torch.cuda.synchronize()
        e0.record()
        for t in dev:
            if t.numel():
                dist.broadcast(t, src=0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        if rank != 0:
            blobs = [t.cpu().numpy() for t in dev]
            coords, orbit, la, n_iter = ReplicatedInputs.unpack(meta, blobs)
            gen_times = meta["gen_times"]
        del dev
    else:
        meta, blobs = ReplicatedInputs.pack(coords, orbit, la, n_iter)
    # the e2e arm uploads from page-locked host memory: the same tables, copied once into pinned buffers
    pinned_blobs = []
    for b in blobs:
        t = torch.empty(max(int(b.size), 1), dtype=torch.uint8, pin_memory=True)
        t[:b.size] = torch.from_numpy(b)
        pinned_blobs.append(t.numpy()[:b.size])
    _, orbit_pinned, la_pinned, _ = ReplicatedInputs.unpack(meta, pinned_blobs)

    r = GPURenderer(local_rank)
    assert r.InitializeMemory(WIDTH, HEIGHT, 1, iter_bytes=4) == 0
    assert r.SetShard(world, rank) == 0
    gen = 1
    assert r.InitializePerturb(gen, orbit, 0, None, la) == 0
    assert r.SyncComputeStream() == 0
    launches0 = r.KernelLaunchCount()

    h2d = orbit.count * orbit.elem_bytes + la.num_las * la.las_elem_bytes + la.num_stages * 8 + la.at_bytes
    hp, wp = r.buffer_shape()
    d2h = hp * wp * 4 + 24
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step_resident():
        flush.zero_()
        torch.cuda.synchronize()
        r.ClearMemory()
        rc = r.RenderPerturbLAv2(alg, coords, n_iter)
        assert rc == 0, GPURenderer.ConvertErrorToString(rc)
        assert r.SyncComputeStream() == 0
        return r.LastRenderMs()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()

    # ---- timed region: K steps, device-timed ---------------------------------------------------------------
    stop, samples = threading.Event(), []
    sampler = threading.Thread(target=_clock_sampler, args=(stop, samples, local_rank), daemon=True)
    sampler.start()
    barrier()
    t_wall0 = time.time()
    kernel_ms = [step_resident() for _ in range(args.steps)]
    barrier()
    wall_s = time.time() - t_wall0
    launches_timed = r.KernelLaunchCount() - launches0 - args.warmup
    # executed-step count for the roofline: one extra, untimed launch of the counting variant of the kernel
    r.EnableStepCounter(True)
    step_resident()
    kinds = r.ReadStepCounters()
    exec_steps = float(kinds["total"])
    credited = float(kinds["at"] * FP32_INSTR_PER_AT_PASS + kinds["la"] * FP32_INSTR_PER_LA_STEP +
                     kinds["perturbation"] * FP32_INSTR_PER_PERTURB_STEP)
    r.EnableStepCounter(False)
    stop.set()
    sampler.join()

    rc, iters, _, red = r.RenderCurrent(n_iter)
    assert rc == 0
    local_sum = int(iters[:HEIGHT, :WIDTH].sum())  # this rank's rows (others are zero)

    # ---- merge the iteration buffer on rank 0 (NCCL reduce of disjoint rows == gather) ----------------------
    total_sum, max_ms = local_sum, sum(kernel_ms)
    gather_ms = None
    if world > 1:
        dev_iters = torch.from_numpy(iters.view(np.int32)).cuda()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.reduce(dev_iters, dst=0, op=dist.ReduceOp.SUM)
        e1.record()
        torch.cuda.synchronize()
        gather_ms = e0.elapsed_time(e1)
        t = torch.tensor([float(max_ms)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        max_ms = float(t.item())
        s = torch.tensor([local_sum], device="cuda", dtype=torch.int64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        total_sum = int(s.item())
        es = torch.tensor([exec_steps, credited, kinds["at"], kinds["la"], kinds["perturbation"]], device="cuda",
                          dtype=torch.float64)
        dist.all_reduce(es, op=dist.ReduceOp.SUM)
        exec_steps, credited = float(es[0].item()), float(es[1].item())
        kinds = {"at": int(es[2].item()), "la": int(es[3].item()), "perturbation": int(es[4].item())}
        if rank == 0:
            merged = dev_iters.cpu().numpy().view(np.uint32)
            assert int(merged[:HEIGHT, :WIDTH].astype(np.int64).sum()) == total_sum

    ms_per_step = max_ms / args.steps
    value = total_sum / (ms_per_step * 1e-3)

    # ---- e2e: public call sequence with host buffers (every step re-uploads orbit + LA, reads results) ------
    # N=1: the frame lands in a pinned buffer.  N>1: ONE host frame in POSIX shared memory, page-locked in every
    # rank; each rank copies just the 4-row bands it rendered into their place (fs_render_current_shard), so the
    # frame is assembled on the host with no collective and 1/N of the frame crosses each GPU's PCIe link.
    shm, frame_kind = None, "pinned"
    if world == 1:
        frame = torch.empty((hp, wp), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    else:
        from fractalshark_b200.sharding import SharedFrame
        name = "fsb200_frame_%s" % os.environ.get("MASTER_PORT", "0")
        shared_ok = [True]
        if rank == 0:
            try:
                shm = SharedFrame(name, (hp, wp), np.uint32, create=True)
            except OSError:
                shared_ok = [False]
        dist.broadcast_object_list(shared_ok, src=0)
        if shared_ok[0]:
            if rank != 0:
                shm = SharedFrame(name, (hp, wp), np.uint32)
            frame = shm.array
            frame_kind = "shared memory frame every rank writes its bands into"
        else:  # no POSIX shared memory on this host: every rank keeps its bands in a pinned frame of its own
            frame = torch.zeros((hp, wp), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
            frame_kind = "per-rank pinned frames (no shared memory on this host)"
        dist.barrier()
    # result sink: the kernel stores finished pixels into the host frame while it runs (fs_set_result_sink page-locks
    # and maps the frame if it is not yet), RenderCurrent then only fetches the 24-byte reduction.  FS_BENCH_SINK=0
    # measures the copy-after-render path instead.
    use_sink = os.environ.get("FS_BENCH_SINK", "1") != "0"
    if use_sink:
        rc = r.SetResultSink(frame)
        assert rc == 0, GPURenderer.ConvertErrorToString(rc)
        frame_kind += "; streamed by the render kernel (result sink)"
    else:
        if shm is not None:
            reg = int(torch.cuda.cudart().cudaHostRegister(frame.ctypes.data, hp * wp * 4, 0))
            assert reg == 0
        frame_kind += "; copied after the render"

    def step_e2e(g):
        rc = r.InitializePerturb(g, orbit_pinned, 0, None, la_pinned)
        assert rc == 0
        r.ClearMemory()
        assert r.RenderPerturbLAv2(alg, coords, n_iter) == 0
        if world == 1:
            rc, it, _, rd = r.RenderCurrent(n_iter, iters_out=frame)
        else:
            rc, rd = r.RenderCurrentShard(n_iter, frame)
        assert rc == 0
        return rd["Sum"]

    for _ in range(2):
        gen += 1
        step_e2e(gen)
    barrier()
    t0 = time.time()
    e2e_sum = 0
    for _ in range(args.steps):
        gen += 1
        e2e_sum = step_e2e(gen)
    barrier()
    e2e_s = (time.time() - t0) / args.steps
    assert e2e_sum == local_sum
    # untimed: one more step into a zeroed host frame -- what arrives there is the whole picture of THIS step
    if rank == 0 or shm is None:
        frame[:] = 0
    barrier()
    gen += 1
    step_e2e(gen)
    barrier()
    if shm is None and world > 1:
        assert int(frame[:HEIGHT, :WIDTH].astype(np.int64).sum()) == local_sum
    elif rank == 0:
        assert int(frame[:HEIGHT, :WIDTH].astype(np.int64).sum()) == total_sum
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        dist.barrier()
    if use_sink:
        assert r.SetResultSink(None) == 0
    elif shm is not None:
        torch.cuda.cudart().cudaHostUnregister(frame.ctypes.data)
    if world > 1:
        del frame
        dist.barrier()
        if shm is not None:
            shm.close()
    # whole-job host<->device bytes per step: every rank uploads its own copy of the tables, the frame leaves once
    h2d, d2h = h2d * world, hp * wp * 4 + 24 * world
    e2e_value = total_sum / e2e_s

    # ---- roofline -----------------------------------------------------------------------------------------------
    peak = GPURenderer.MeasureFp32IssuePeak(local_rank)  # FFMA thread-instr/s, measured live
    achieved = credited / (ms_per_step * 1e-3)
    roofline = {"bound": "fp32_issue", "achieved": achieved / 1e12, "peak": peak * world / 1e12, "unit": "T FP32 instr/s",
                "frac": achieved / (peak * world),
                # dram__bytes_read.sum + dram__bytes_write.sum of the render kernel, one `ncu --set full` capture per view
                # (profiles/r1_lav2_v12_view14_summary.md, r1_lav2_v7_summary.md); bytes per launch
                "traffic": {14: 5.47e6 + 0.15e6, 5: 0.69e6 + 0.09e6}.get(VIEW_ID),
                "note": "compute-bound scalar path (SURVEY.md 8d): not HBM, not tensor. achieved = executed "
                        "steps/launch by kind (device counters) x FP32 mantissa instr per step (AT pass 9, LA step 22, "
                        "HDRx32 perturbation step 20) / kernel time (the credit is the reference formulation's work per "
                        "pass; the kernel's lean chunk loop issues 7 FP32 + 2/16 for the escape test per AT pass); "
                        "peak = FFMA issue rate measured live by fs_measure_fp32_issue_peak on this GPU "
                        "(MEASURED_PEAKS.json has only HBM/bf16 peaks).",
                "executed_steps_per_launch": exec_steps,
                "executed_steps_by_kind": {k: int(kinds[k]) for k in ("at", "la", "perturbation")},
                "skip_factor": total_sum / max(exec_steps, 1.0)}

    line = {"metric": METRIC, "value": value, "unit": "pixel-iters/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32+i32 (HDRx32)",
            "data": f"synthetic (View #{VIEW_ID} preset coordinates; orbit via GMP + LA table generated in-process, untimed: "
                    f"{gen_times['orbit_s']:.3f}s + {gen_times['la_s']:.3f}s)",
            "config": {"workload": WORKLOAD, "l2": "flushed between timed iterations (256 MiB memset)",
                       "sharding": f"4-row tile bands round-robin over {world} rank(s)",
                       "orbit_entries": orbit.count, "la_records": la.num_las, "la_stages": la.stage_count},
            "clocks": _clock_summary(samples),
            "e2e": {"value": e2e_value, "unit": "pixel-iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "host_frame": frame_kind},
            "gpu_launches": int(launches_timed),
            "roofline": roofline,
            "wall_ms_per_step_incl_flush": wall_s * 1e3 / args.steps,
            "sum_pixel_iters": total_sum}
    if gather_ms is not None:
        line["gather_ms"] = gather_ms
        line["input_broadcast_ms"] = bcast_ms

    if rank == 0:
        # reference CUDA kernels (oracle/_ref, the checker) timed on the same GPU and inputs, for context
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import ref_renderer
            if ref_renderer.available() and world == 1:
                rr = ref_renderer.RefGPURenderer()
                assert rr.InitializeMemory(WIDTH, HEIGHT, 1, iter_bytes=4) == 0
                assert rr.InitializePerturb(1, orbit, 0, None, la) == 0
                ms = []
                for i in range(3):
                    flush.zero_()
                    torch.cuda.synchronize()
                    rr.ClearMemory()
                    assert rr.RenderPerturbLAv2(alg, coords, n_iter) == 0
                    assert rr.SyncComputeStream() == 0
                    ms.append(rr.LastRenderMs())
                rc, ref_iters, _, ref_red = rr.RenderCurrent(n_iter)
                exact = float((ref_iters[:HEIGHT, :WIDTH] == iters[:HEIGHT, :WIDTH]).mean())
                line["reference_cuda_kernel"] = {"ms_per_step": min(ms[1:]), "value": ref_red["Sum"] / (min(ms[1:]) * 1e-3),
                                                 "unit": "pixel-iters/s", "pixels_exact_vs_ours": exact,
                                                 "what": "reference mandel_1xHDR_float_perturb_lav2 built for sm_100a"}
                rr.close()
        except Exception as e:  # the checker is optional here
            line["reference_cuda_kernel"] = {"unavailable": repr(e)[:200]}
        if not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_cpu
            threads = oracle_cpu.hardware_threads()
            v, dt, sample = cpu_port_sample(coords, orbit, la, n_iter, threads)
            line["cpu_baseline"] = {"value": v, "unit": "pixel-iters/s", "cores": threads, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
