#!/bin/bash
# GPU visit "r3e" (last of the round): whole GPU suite + smoke at HEAD, bench --config hgen + ncu of the final k_cinterp,
# the default bench.py line as the driver runs it
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=r3e
bash tools/r2_check.sh $TAG tests smoke
timeout 900 python bench.py --config hgen --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_hgen.json 2> gpurun_out/${TAG}_bench_hgen.err
echo "hgen exit $?"; tail -3 gpurun_out/${TAG}_bench_hgen.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_hgen.json").read().strip().splitlines()[-1])
print("hgen value %.3e ms %.3f" % (d["value"], d["ms_per_step"]), d["roofline"]["achieved"], d["roofline"]["frac"], d["parity"], d["e2e"]["value"], d["cpu_baseline"]["value"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cinterp -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_k_cinterp \
   python bench.py --config hgen --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncufull_cinterp.log 2>&1
echo "ncu cinterp exit $?"
timeout 1500 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench_c3.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_c3.json").read().strip().splitlines()[-1])
print("C3 value %.0f ms %.2f" % (d["value"], d["ms_per_step"]), d["roofline"]["kernel_ms_per_step"], d["roofline"]["whole_step"], d.get("parity"), "e2e", d.get("e2e",{}).get("value"), d.get("cpu_baseline",{}).get("value"), d["roofline"]["frac"], d["roofline"]["traffic"])
PY
