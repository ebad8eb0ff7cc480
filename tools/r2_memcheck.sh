#!/bin/bash
# Round 2: compute-sanitizer memcheck of the kernels added or changed this round, small frames (run under gpurun).
run() { echo "== $*"; compute-sanitizer --tool memcheck --error-exitcode 9 "$@" 2>&1 | grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|new:" | head -6; echo "rc=${PIPESTATUS[0]}"; }
run python tools/gpu_one.py 14 GpuHDRx32PerturbedLAv2 0 96 54 --noref
run python tools/gpu_one.py 5 GpuHDRx32PerturbedLAv2 0 160 96 --noref
FS_LAV2_POOL=1 run python tools/gpu_one.py 14 GpuHDRx32PerturbedLAv2 0 96 54 --noref
FS_LAV2_POOL=1 run python tools/gpu_one.py 5 GpuHDRx32PerturbedLAv2PO 2000 100 37 --noref
FS_PROBE_PASSES=256 run python tools/gpu_one.py 14 GpuHDRx32PerturbedLAv2 0 96 54 --noref
run python tools/gpu_one.py 14 GpuHDRx32PerturbedBLA 0 64 36 --noref
run python tools/gpu_one.py 5 GpuHDRx32PerturbedBLA 0 96 54 --noref
run python tools/gpu_one.py 14 GpuHDRx2x32PerturbedLAv2 0 64 36 --noref
run python tools/gpu_one.py 14 GpuHDRx64PerturbedLAv2 0 64 36 --noref
