#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for k in 1 2; do
timeout 800 python bench.py --steps 2 --warmup 3 --no-cpu --no-pageable > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3i_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"])
PY
done
