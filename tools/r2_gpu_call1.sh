#!/bin/bash
# round 2, call 1: fixture set 7 from the reference kernels, the whole GPU suite, a bench line, ncu with the pipe split
set -x
mkdir -p gpurun_out
python tests/golden/make_golden.py gpurun_out/ref_gpu_small7.npz 7 > gpurun_out/golden7.log 2>&1 && cp gpurun_out/ref_gpu_small7.npz tests/golden/
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_c1.log 2>&1
tail -15 gpurun_out/gputest_c1.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
tail -c 600 gpurun_out/bench_c1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lav2_kernel -s 1 -c 1 -o gpurun_out/prof_c1 python tools/gpu_one.py 14 GpuHDRx32PerturbedLAv2 0 --noref > gpurun_out/ncu_c1.log 2>&1
ncu -i gpurun_out/prof_c1.ncu-rep --page raw --csv > gpurun_out/prof_c1_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
