#!/bin/bash
# Dev: bench lines of the other BASELINE workloads at N GPUs (run under gpurun --gpus N): tools/scale_others.sh N wl...
N=$1; shift
for wl in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --no-cpu-baseline 2> gpurun_out/bench_${wl}_n$N.err | grep "^{" > gpurun_out/bench_${wl}_n$N.json
  python -c "
import json
d=json.load(open('gpurun_out/bench_${wl}_n$N.json'))
print(d['config']['workload'], d['n_gpus'], round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"
done
