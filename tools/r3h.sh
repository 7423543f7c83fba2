#!/bin/bash
# GPU visit "r3h": pageable e2e through the staging ring (host threads 4 / 8 / 16 / 32) against direct copies; staging tests
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "staging_ring or chunked or tapered" 2>&1 | tail -2
nproc
OAK_B200_STAGE_THREADS="4,8,16,32" timeout 800 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err
tail -2 gpurun_out/r3h_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3h_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("pageable_columns_per_s"))
PY
