#!/bin/bash
# One GPU-box visit: parity tests, smoke, a reduced bench, optional full bench / ncu passes.
# usage: tools/gpu_check.sh [tests] [smoke] [sanitize] [benchsmall] [bench] [ncu] [ncufull] [ncuvar]
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
for what in "$@"; do
case $what in
tests)
  timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/tests.log
  tail -25 gpurun_out/tests.log ;;
testsall)
  timeout 1800 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/tests.log
  tail -40 gpurun_out/tests.log ;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log ;;
sanitize)
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize.log 2>&1
  echo "sanitize exit $?"; tail -15 gpurun_out/sanitize.log ;;
benchsmall0)
  timeout 900 python bench.py --nx 300 --ny 300 --nz 30 --nobs 90000 --steps 2 --warmup 1 --no-cpu --no-e2e --eig-kernel 0 > gpurun_out/bench_small0.json 2> gpurun_out/bench_small0.err
  echo "benchsmall0 exit $?"; tail -3 gpurun_out/bench_small0.err; cat gpurun_out/bench_small0.json ;;
benchsmall)
  timeout 900 python bench.py --nx 300 --ny 300 --nz 30 --nobs 90000 --steps 2 --warmup 1 --cpu-seconds 5 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
  echo "benchsmall exit $?"; tail -3 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json ;;
bench)
  timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json ;;
benchref)
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "benchref exit $?"; tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json ;;
ncu)
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_|DeviceRadixSort|DeviceScan' -c 600 --csv --log-file gpurun_out/launches.csv \
     python bench.py --nx 400 --ny 400 --nobs 160000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"; tail -3 gpurun_out/ncu_bench.log ;;
ncutri)
  for kn in k_tridiag k_tvec k_gram k_tql; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kn -s 3 -c 1 -f -o gpurun_out/prof_$kn \
     python bench.py --nx 300 --ny 300 --nobs 90000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncufull_$kn.log 2>&1
  echo "ncufull $kn exit $?"
  done ;;
ncuvar)
  # --set full captures of the kernels behind the round-1-end options (VARIANT = bench flags, e.g. "--gram-kernel 1 --fuse-apply 1")
  for kn in k_gram_mma k_tvec k_apply_mma; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kn -s 3 -c 1 -f -o gpurun_out/prof_var_$kn \
     python bench.py --nx 300 --ny 300 --nobs 90000 --steps 1 --warmup 1 --no-e2e --no-cpu ${VARIANT:---gram-kernel 1 --fuse-apply 1 --apply-kernel 1} > gpurun_out/ncufull_var_$kn.log 2>&1
  echo "ncufull $kn exit $?"
  done ;;
ncufull)
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_eig_fast -s 2 -c 2 -f -o gpurun_out/prof_eig \
     python bench.py --nx 400 --ny 400 --nobs 160000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncufull_eig.log 2>&1
  echo "ncufull eig exit $?"
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_gram -s 2 -c 2 -f -o gpurun_out/prof_gram \
     python bench.py --nx 400 --ny 400 --nobs 160000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncufull_gram.log 2>&1
  echo "ncufull gram exit $?"
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_apply -s 2 -c 2 -f -o gpurun_out/prof_apply \
     python bench.py --nx 400 --ny 400 --nobs 160000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncufull_apply.log 2>&1
  echo "ncufull apply exit $?" ;;
esac
done
