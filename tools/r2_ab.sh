#!/bin/bash
# Round-2 A/B on ONE B200 (run through gpurun): the variants prepared without a GPU at the end of round 1, each
# parity-checked under the CPU emulation (tools/cuemu) but never timed.  Order: parity of the new options on the
# real device first, then per-kernel times of the reduced bench (300 x 300 grid, 90 k zones) for every combination,
# then the full C3 line of the best one.  Results: gpurun_out/r2_ab.log
#   usage: python tools/r2_prepare.py        (here, no GPU: builds the -D variants into oak_b200/variants/, they travel)
#          gpurun --timeout 2400 -- 'bash tools/r2_ab.sh'
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
LOG=gpurun_out/r2_ab.log
: > $LOG
echo "== parity of the new options on the device" | tee -a $LOG
timeout 900 python -m pytest tests/test_variants_gpu.py -q -p no:cacheprovider 2>&1 | tail -5 | tee -a $LOG
small="--nx 300 --ny 300 --nz 30 --nobs 90000 --steps 3 --warmup 2 --no-cpu --no-e2e"
for combo in "0 0" "1 0" "2 0" "3 0" "4 0" "0 1" "1 1" "3 1"; do
  set -- $combo
  echo "== gram_kernel $1 fuse_apply $2" | tee -a $LOG
  timeout 600 python bench.py $small --gram-kernel $1 --fuse-apply $2 2>>gpurun_out/r2_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), d['roofline']['kernel_ms_per_step'])" | tee -a $LOG
done
echo "== apply_kernel 1 (gram 0, fuse 0)" | tee -a $LOG
timeout 600 python bench.py $small --apply-kernel 1 2>>gpurun_out/r2_ab.err | tail -1 | cut -c1-400 | tee -a $LOG
echo "== global scheme, state resident (tools/time_global.py), apply_kernel 0 / 1" | tee -a $LOG
timeout 600 python tools/time_global.py 2>>gpurun_out/r2_ab.err | tail -1 | tee -a $LOG
timeout 600 python tools/time_global.py --apply-kernel 1 2>>gpurun_out/r2_ab.err | tail -1 | tee -a $LOG
for sp in "0 0" "0 1" "1 1"; do
  set -- $sp
  echo "== tvec_split 1, gram_kernel $1 fuse_apply $2" | tee -a $LOG
  timeout 600 python bench.py $small --tvec-split 1 --gram-kernel $1 --fuse-apply $2 2>>gpurun_out/r2_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), d['roofline']['kernel_ms_per_step'])" | tee -a $LOG
done
# twisted_vector2 (interleaved pivot recurrences) as a build variant of the same sources
[ -f oak_b200/variants/liboak_tw2.so ] || python tools/r2_prepare.py 2>&1 | tail -3 | tee -a $LOG
echo "== TVEC_TWISTED2=1 (gram 0, fuse 0)" | tee -a $LOG
OAK_B200_LIB=$PWD/oak_b200/variants/liboak_tw2.so timeout 600 python bench.py $small 2>>gpurun_out/r2_ab.err | tail -1 | cut -c1-400 | tee -a $LOG
echo "== TQL_LOCAL=1 (k_tql without shared memory; gram 0, fuse 0)" | tee -a $LOG
OAK_B200_LIB=$PWD/oak_b200/variants/liboak_tqll.so timeout 600 python bench.py $small 2>>gpurun_out/r2_ab.err | tail -1 | cut -c1-400 | tee -a $LOG
echo "== full C3 line with gram_kernel 1 + fuse_apply 1 (compare with profiles/r1_bench_c3_1gpu.json)" | tee -a $LOG
timeout 900 python bench.py --steps 3 --warmup 3 --gram-kernel 1 --fuse-apply 1 --no-cpu > gpurun_out/r2_bench_c3_g1f1.json 2>>gpurun_out/r2_ab.err
cut -c1-600 gpurun_out/r2_bench_c3_g1f1.json | tee -a $LOG
