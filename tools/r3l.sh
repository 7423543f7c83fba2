#!/bin/bash
# GPU visit "r3l": k_gram_mma c vector with 4 partial sums (GRAM_C4) against one chain
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "rrsqrt or local_analysis_matches or localise_obs_false or edge" 2>&1 | tail -2
run() {
  label=$1; shift
  echo "== $label"
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/r3l_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), {k: round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d.get('parity',{}).get('ok'))"
}
run c4 A=1
run c1 OAK_B200_LIB=$PWD/oak_b200/variants/liboak_c1.so
