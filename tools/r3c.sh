#!/bin/bash
# GPU visit "r3c": hgen tests + bench + ncu with the table-driven k_cinterp; C3 A/B of the candidate prefetch in k_gram_mma
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=r3c
timeout 600 python -m pytest tests/test_hgen_gpu.py tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -2
timeout 900 python bench.py --config hgen --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_hgen.json 2> gpurun_out/${TAG}_bench_hgen.err
echo "hgen exit $?"; tail -3 gpurun_out/${TAG}_bench_hgen.err; cut -c1-200 gpurun_out/${TAG}_bench_hgen.json
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_hgen.json").read().strip().splitlines()[-1])
print("hgen value %.3e ms %.3f" % (d["value"], d["ms_per_step"]), d["roofline"]["achieved"], d["roofline"]["frac"], d["parity"], d["e2e"]["value"], d["cpu_baseline"]["value"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cinterp -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_k_cinterp \
   python bench.py --config hgen --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncufull_cinterp.log 2>&1
echo "ncu cinterp exit $?"
run() {
  label=$1; shift
  echo "== $label"
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e $EXTRA 2>>gpurun_out/${TAG}_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), {k: round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d.get('parity',{}).get('ok'), d['gpu_launches'])"
}
run gram_prefetch A=1
run gram_no_prefetch OAK_B200_LIB=$PWD/oak_b200/variants/liboak_gram_nopf.so
run gram_prefetch_again A=1
NCU_KERNELS="k_gram_mma" bash tools/r2_check.sh $TAG ncufull
