#!/bin/bash
# GPU visit "r2w": C3 A/B of (a) product build (flat QL loop + tracked block ends), (b) PWK_SHORT_CHAIN=1, (c..e) caps on the
# residency of k_tridiag_warp / k_tvec (unused dynamic shared memory) so that other streams' CTAs fit beside them
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=r2w
run() {  # label, env...
  label=$1; shift
  echo "== $label"
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/${TAG}_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), {k: round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d.get('parity',{}).get('ok'))"
}
run product A=1
run short_chain OAK_B200_LIB=$PWD/oak_b200/variants/liboak_short.so
run tri_pad_6 OAK_B200_TRI_PAD=27000
run tvec_pad_3 OAK_B200_TVEC_PAD=16000
run tri6_tvec3 OAK_B200_TRI_PAD=27000 OAK_B200_TVEC_PAD=16000
run tri_pad_7 OAK_B200_TRI_PAD=23000
