#!/bin/bash
# GPU visit "r2u": parity suite, A/B of the flat QL loop (PWK_FLAT 1 = product build, 0 = variant), C3 line, C5 line
# (ensemble entry point, fused vs three-pass), ncu of k_tql and the fused apply.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=r2u
bash tools/r2_check.sh $TAG tests smoke
small="--nx 300 --ny 300 --nz 30 --nobs 90000 --steps 3 --warmup 2 --no-cpu --no-e2e"
for lib in oak_b200/liboak_b200.so oak_b200/variants/liboak_pwk0.so; do
  echo "== $lib"
  OAK_B200_LIB=$PWD/$lib timeout 600 python bench.py $small 2>gpurun_out/${TAG}_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), d['roofline']['kernel_ms_per_step'], d.get('parity',{}).get('ok'))"
done
echo "== C3, PWK_FLAT=0 (3 steps, no cpu)"
OAK_B200_LIB=$PWD/oak_b200/variants/liboak_pwk0.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/${TAG}_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), d['roofline']['kernel_ms_per_step'], d.get('parity',{}).get('ok'))"
echo "== C3, product build"
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench_c3.err; cut -c1-200 gpurun_out/${TAG}_bench_c3.json
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().splitlines()[-1])
print(d["roofline"]["kernel_ms_per_step"], d["roofline"]["whole_step"], d.get("parity"), d.get("e2e",{}).get("value"))
PY
echo "== C5"
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
echo "c5 exit $?"; tail -3 gpurun_out/${TAG}_bench_c5.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_c5.json").read().strip().splitlines()[-1])
print("c5 value %.0f ms %.2f" % (d["value"], d["ms_per_step"]), d["roofline"]["ens_fuse"], d["parity"], d["e2e"])
PY
NCU_KERNELS="k_tql" bash tools/r2_check.sh $TAG ncufull
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_k_apply_ens \
   python bench.py --config c5 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncufull_apply_ens.log 2>&1
echo "ncu apply_ens exit $?"
