#!/bin/bash
# GPU visit "r3j" (last): whole GPU suite + smoke at HEAD; C5 line (ensemble entry point, e2e through the staged copies); the
# default bench.py line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=r3j
bash tools/r2_check.sh $TAG tests smoke
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
echo "c5 exit $?"; tail -2 gpurun_out/${TAG}_bench_c5.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_c5.json").read().strip().splitlines()[-1])
print("c5 value %.0f ms %.2f" % (d["value"], d["ms_per_step"]), d["e2e"])
PY
timeout 900 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
echo "bench exit $?"; tail -2 gpurun_out/${TAG}_bench_c3.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().splitlines()[-1])
print("C3 value %.0f ms %.2f" % (d["value"], d["ms_per_step"]), d["roofline"]["whole_step"]["frac_of_fp64_peak_per_gpu"], d.get("parity",{}).get("ok"), "e2e", d["e2e"]["value"], d["e2e"].get("pageable_columns_per_s"), d["cpu_baseline"]["value"])
PY
