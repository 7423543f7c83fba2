#!/bin/bash
# Round-2 GPU visit: parity suite, smoke, full C3 bench line (with CPU baseline + in-bench parity), ncu launch list
# and --set full captures of the default kernels.   usage: gpurun --timeout 1800 -- 'bash tools/r2_check.sh [tag] [what...]'
TAG=${1:-r2}; shift
WHAT=${@:-tests smoke bench ncu ncufull}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_gpu.csv 2>&1
for what in $WHAT; do
case $what in
tests)
  timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/${TAG}_tests.log
  tail -8 gpurun_out/${TAG}_tests.log ;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log ;;
bench)
  timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json ;;
benchsmall)
  timeout 900 python bench.py --nx 300 --ny 300 --nz 30 --nobs 90000 --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_small.json 2> gpurun_out/${TAG}_bench_small.err
  echo "benchsmall exit $?"; tail -3 gpurun_out/${TAG}_bench_small.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_small.json").read().strip().splitlines()[-1])
print("value %.0f  ms/step %.2f" % (d["value"], d["ms_per_step"]), d["roofline"]["kernel_ms_per_step"], d.get("parity"))
PY
  ;;
ncu)
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_|DeviceRadixSort|DeviceScan' -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv \
     python bench.py --nx 400 --ny 400 --nobs 160000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
  echo "ncu exit $?"; tail -2 gpurun_out/${TAG}_ncu_bench.log | cut -c1-200 ;;
ncufull)
  for kn in ${NCU_KERNELS:-k_gram_mma k_tridiag k_tql k_tvec k_apply}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kn -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_$kn \
     python bench.py --nx 300 --ny 300 --nobs 90000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_ncufull_$kn.log 2>&1
  echo "ncufull $kn exit $?"
  done ;;
sanitize)
  for tool in racecheck synccheck; do
  timeout 1700 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "${SAN_K:-known_answers or local_analysis_matches or edge}" > gpurun_out/${TAG}_$tool.log 2>&1
  echo "$tool exit $?"; tail -5 gpurun_out/${TAG}_$tool.log
  done ;;
esac
done
