#!/bin/bash
# GPU visit "r3g": chunk size of the host-streaming path (e2e leg, pinned buffers)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
OAK_B200_CHUNK_MB="64,128,256,512,1024,2048" timeout 800 python bench.py --steps 2 --warmup 3 --no-cpu --no-pageable > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err
grep chunk_mb gpurun_out/r3g_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r3g_bench.json").read().strip().splitlines()[-1])
print("e2e", d["e2e"])
PY
