#!/bin/bash
# Round-2 multi-GPU: copy-engine pushes against the push kernel (k_push).  usage: gpurun --gpus N -- 'bash tools/r2_mg2.sh N "<variants>"'
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
LOG=gpurun_out/r2_mg2_${N}gpu.log
: > $LOG
run() {
  echo "== $*" | tee -a $LOG
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-e2e "$@" 2>>gpurun_out/r2_mg2.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f columns/s  ms/step %.2f  %s  parity %s' % (d['value'], d['ms_per_step'], d['config']['parallelism'][14:70] + ' ... ' + d['config']['parallelism'][-12:], d.get('parity', {}).get('ok')))" | tee -a $LOG
}
IFS=';' read -ra VARS <<< "${2:-;--push-kernel 1;--push-kernel 1 --push-ctas 16;--push-kernel 1 --push-ctas 64}"
for v in "${VARS[@]}"; do run $v; done
