#!/bin/bash
# GPU visit "r2z": is the exposed half of k_tql false serialisation between streams that share a hardware queue?
# (CUDA_DEVICE_MAX_CONNECTIONS, default 8; the library uses 4 slot streams + side, push and peer streams) ; more stream slots
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=r2z
run() {
  label=$1; shift
  echo "== $label"
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e $EXTRA 2>>gpurun_out/${TAG}_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), {k: round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d.get('parity',{}).get('ok'), d['gpu_launches'])"
}
run default A=1
run conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_instream CUDA_DEVICE_MAX_CONNECTIONS=32 OAK_B200_OPTIONS=tql_side=0
run conn32_zb24k CUDA_DEVICE_MAX_CONNECTIONS=32 OAK_B200_ZB=24576
run nslot6_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32 OAK_B200_LIB=$PWD/oak_b200/variants/liboak_nslot6.so
run nslot8_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32 OAK_B200_LIB=$PWD/oak_b200/variants/liboak_nslot8.so
run nslot8_conn32_zb8k CUDA_DEVICE_MAX_CONNECTIONS=32 OAK_B200_ZB=8192 OAK_B200_LIB=$PWD/oak_b200/variants/liboak_nslot8.so
