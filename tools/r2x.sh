#!/bin/bash
# GPU visit "r2x": C3 sweeps around the faster k_tql: zones per batch, Gram variant with half the shared memory, k_tql in stream
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=r2x
run() {
  label=$1; shift
  echo "== $label"
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e $EXTRA 2>>gpurun_out/${TAG}_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), {k: round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d.get('parity',{}).get('ok'), d['gpu_launches'])"
}
run zb8k OAK_B200_ZB=8192
run zb12k OAK_B200_ZB=12288
run zb24k OAK_B200_ZB=24576
run tql_in_stream OAK_B200_OPTIONS=tql_side=0
EXTRA="--gram-kernel 3" run gram3 A=1
EXTRA="--gram-kernel 4" run gram4 A=1
run halves2 OAK_B200_EIG_HALVES=2
