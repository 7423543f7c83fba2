#!/bin/bash
# durations of the streaming kernels of the ensemble branch (C5 sizes: n = 2.62e6 rows, N = 64, m = 2e5) for their HBM figures
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_mean_anom|k_epilogue|k_obsoper|k_pack_obs|k_coo' -c 60 --csv --log-file gpurun_out/r3o_stream_kernels.csv \
  python bench.py --config c5 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r3o.log 2>&1
echo "exit $?"; tail -1 gpurun_out/r3o.log | cut -c1-100
