#!/bin/bash
# Round 2: ncu --set full captures of the kernels behind the other BASELINE configs (run under gpurun); only the text
# extracts are kept (the reports together exceed what gpurun copies back)
mkdir -p gpurun_out/r2_others
cap() { name=$1; kern=$2; shift 2
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:$kern -s 1 -c 1 -o /tmp/prof_$name python tools/gpu_one.py "$@" --noref > gpurun_out/r2_others/$name.log 2>&1
  python tools/ncu_brief.py /tmp/prof_$name.ncu-rep > gpurun_out/r2_others/$name.metrics.txt 2>&1
  python tools/ncu_phase_split.py /tmp/prof_$name.ncu-rep 2>/dev/null | grep -v " 0.0 M warp" > gpurun_out/r2_others/$name.sass_regions.txt
  rm -f /tmp/prof_$name.ncu-rep; grep "gpu__time_duration" gpurun_out/r2_others/$name.metrics.txt; }
cap v5_lav2 lav2_kernel 5 GpuHDRx32PerturbedLAv2 0
cap v5_bla bla_kernel 5 GpuHDRx32PerturbedBLA 0
cap v14_bla bla_kernel 14 GpuHDRx32PerturbedBLA 0 1920 1080
cap v14_2x32 lav2_kernel 14 GpuHDRx2x32PerturbedLAv2 0 1920 1080
cap v0_f32 direct_kernel 0 Gpu1x32 65536
cap v0_f64 direct_kernel 0 Gpu1x64 65536
