#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_c_abi_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "production_selection or c_program or staging" 2>&1 | tail -3
