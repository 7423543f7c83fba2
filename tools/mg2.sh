#!/bin/bash
# 2-GPU variants on the per-rank load of the 8-GPU C3 run: bash tools/mg2.sh "<bench args A>" ...
i=0
for args in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --nx 500 --ny 500 --nobs 250000 --steps 3 --warmup 2 --no-cpu --no-e2e $args > gpurun_out/bench_2gpu_$i.json 2> gpurun_out/bench_2gpu_$i.err
  echo "exit $?"; tail -4 gpurun_out/bench_2gpu_$i.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_2gpu_$i.json').read().strip().splitlines()[-1]); print('[$args]', 'value %.0f'%d['value'], 'ms %.2f'%d['ms_per_step'], d['config']['parallelism'][:200], d['roofline']['kernel_ms_per_step'])"
  i=$((i+1))
done
