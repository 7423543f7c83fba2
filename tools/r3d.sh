#!/bin/bash
# 8-GPU visit "r3d": strong scaling of C3 with the round-2 kernels (multicast gather), zones per batch at 125 k zones per rank
N=${1:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
LOG=gpurun_out/r3d_${N}gpu.log
: > $LOG
run() {
  label=$1; shift
  echo "== $label" | tee -a $LOG
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 4 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/r3d.err | tail -1 | tee gpurun_out/r3d_${N}gpu_$label.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f columns/s  ms/step %.2f  %s  parity %s' % (d['value'], d['ms_per_step'], d['config']['parallelism'][14:60], d.get('parity', {}).get('ok')), {k: round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()})" | tee -a $LOG
}
if [ "$2" = "second" ]; then
run taper1 OAK_B200_OPTIONS=taper=1
run zb5920 OAK_B200_ZB=5920
run taper1_zb5920 OAK_B200_OPTIONS=taper=1 OAK_B200_ZB=5920
run nslot8 OAK_B200_LIB=$PWD/oak_b200/variants/liboak_nslot8.so
else
run default A=1
run zb16k OAK_B200_ZB=15984
run zb10k OAK_B200_ZB=10656
fi
