#!/bin/bash
# GPU visit "r3a": full parity suite (incl. the observation-operator weights), bench --config hgen + ncu of k_cinterp,
# C3 with batches up to 24 k zones: 4 / 6 / 8 stream slots
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=r3a
bash tools/r2_check.sh $TAG tests smoke
timeout 900 python bench.py --config hgen --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_hgen.json 2> gpurun_out/${TAG}_bench_hgen.err
echo "hgen exit $?"; tail -3 gpurun_out/${TAG}_bench_hgen.err; cut -c1-1500 gpurun_out/${TAG}_bench_hgen.json
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
run nslot4 A=1
run nslot6 OAK_B200_LIB=$PWD/oak_b200/variants/liboak_nslot6.so
run nslot8 OAK_B200_LIB=$PWD/oak_b200/variants/liboak_nslot8.so
run nslot8_zb16k OAK_B200_ZB=16576 OAK_B200_LIB=$PWD/oak_b200/variants/liboak_nslot8.so
