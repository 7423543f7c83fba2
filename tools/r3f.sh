#!/bin/bash
# GPU visit "r3f": whole GPU suite + smoke at HEAD, bench --config hgen
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=r3f
bash tools/r2_check.sh $TAG tests smoke
timeout 900 python bench.py --config hgen --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_hgen.json 2> gpurun_out/${TAG}_bench_hgen.err
echo "hgen exit $?"; tail -3 gpurun_out/${TAG}_bench_hgen.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_hgen.json").read().strip().splitlines()[-1])
print("hgen value %.3e ms %.3f" % (d["value"], d["ms_per_step"]), d["roofline"]["achieved"], d["roofline"]["frac"], d["e2e"]["value"])
PY
