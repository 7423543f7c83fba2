#!/bin/bash
# 8-GPU pipeline variants: bash tools/mg8.sh "<bench args A>" "<bench args B>" ...
i=0
for args in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu $args > gpurun_out/bench_8gpu_$i.json 2> gpurun_out/bench_8gpu_$i.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_8gpu_$i.json').read().strip().splitlines()[-1]); print('[$args]', 'value %.0f'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e', d['e2e'] and '%.0f'%d['e2e']['value'], d['roofline']['kernel_ms_per_step'])"
  i=$((i+1))
done
