#!/bin/bash
# GPU visit "r3k": k_tvec with 8 / 4 accumulation chains (TVEC_DOT8), per-kernel times of the 300 x 300 case and the C3 step
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "rrsqrt or degenerate or local_analysis_matches or dynamic" 2>&1 | tail -2
run() {
  label=$1; shift
  echo "== $label"
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/r3k_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), {k: round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d.get('parity',{}).get('ok'))"
}
run dot8 A=1
run dot4 OAK_B200_LIB=$PWD/oak_b200/variants/liboak_dot4.so
