#!/bin/bash
# last check of the round: GPU suite + smoke at HEAD
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash tools/r2_check.sh r3m tests smoke
