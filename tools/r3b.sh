#!/bin/bash
# GPU visit "r3b": bench --config hgen (+ ncu of k_cinterp); the full C3 line (CPU baseline, e2e, parity) and the C5 line of
# the final kernels; ncu launch list and --set full captures of the kernels that changed (k_tql, k_tvec, k_gram_mma)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=r3b
timeout 900 python bench.py --config hgen --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_hgen.json 2> gpurun_out/${TAG}_bench_hgen.err
echo "hgen exit $?"; tail -3 gpurun_out/${TAG}_bench_hgen.err; cut -c1-2500 gpurun_out/${TAG}_bench_hgen.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cinterp -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_k_cinterp \
   python bench.py --config hgen --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncufull_cinterp.log 2>&1
echo "ncu cinterp exit $?"
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench_c3.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().splitlines()[-1])
print("C3 value %.0f ms %.2f" % (d["value"], d["ms_per_step"]), d["roofline"]["kernel_ms_per_step"], d["roofline"]["whole_step"], d.get("parity"), "e2e", d.get("e2e",{}).get("value"), d.get("cpu_baseline",{}).get("value"))
PY
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
echo "c5 exit $?"; tail -3 gpurun_out/${TAG}_bench_c5.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_c5.json").read().strip().splitlines()[-1])
print("c5 value %.0f ms %.2f" % (d["value"], d["ms_per_step"]), d["roofline"]["ens_fuse"], d["parity"], d["e2e"])
PY
NCU_KERNELS="k_gram_mma k_tridiag k_tql k_tvec k_apply" bash tools/r2_check.sh $TAG ncu ncufull
