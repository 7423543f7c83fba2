#!/bin/bash
# Round-2 multi-GPU experiments (run through `gpurun --gpus N`, N = 2 develops, 8 confirms): the fused gather with
# the apply + push of every batch in pieces (option push_pieces), with the default kernels and with the variants
# tools/r2_ab.sh found faster on one GPU (pass them as VARIANT, e.g. VARIANT="--gram-kernel 1 --fuse-apply 1").
#   usage: gpurun --gpus 8 --timeout 1500 -- 'VARIANT="--gram-kernel 1 --fuse-apply 1" bash tools/r2_mg.sh 8'
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
LOG=gpurun_out/r2_mg_${N}gpu.log
: > $LOG
run() {
  echo "== $*" | tee -a $LOG
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-e2e "$@" 2>>gpurun_out/r2_mg.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f columns/s  ms/step %.2f  %s' % (d['value'], d['ms_per_step'], d['config']['parallelism'][-60:]))" | tee -a $LOG
}
run
run --push-pieces 2
run --push-pieces 4
if [ -n "$VARIANT" ]; then
  run $VARIANT
  run $VARIANT --push-pieces 2
  run $VARIANT --push-pieces 4
fi
run --peer-mode 2   # experiment: nothing pushed (compute-only lower bound, not a valid bench line)
