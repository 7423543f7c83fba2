#!/bin/bash
# GPU visit "r2y": k_tql without shared memory (global scratch + register prefetch queue) against the shared-memory form
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=r2y
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -2
run() {
  label=$1; shift
  echo "== $label"
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e $EXTRA 2>>gpurun_out/${TAG}_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), {k: round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d.get('parity',{}).get('ok'), d['gpu_launches'])"
}
run global_pf6 A=1
run smem OAK_B200_LIB=$PWD/oak_b200/variants/liboak_tqlsmem.so
run global_pf10 OAK_B200_LIB=$PWD/oak_b200/variants/liboak_tqlpf10.so
run global_pf3 OAK_B200_LIB=$PWD/oak_b200/variants/liboak_tqlpf3.so
run global_pf6_zb24k OAK_B200_ZB=24576
run global_pf6_zb33k OAK_B200_ZB=33024
run smem_zb33k OAK_B200_ZB=33024 OAK_B200_LIB=$PWD/oak_b200/variants/liboak_tqlsmem.so
NCU_KERNELS="k_tql" bash tools/r2_check.sh $TAG ncufull
