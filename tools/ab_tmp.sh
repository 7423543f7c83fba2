small="--nx 300 --ny 300 --nz 30 --nobs 90000 --steps 3 --warmup 2 --no-cpu --no-e2e"
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -4
for opt in "apply_tma=1" "apply_tma=0"; do
  echo "== $opt"
  OAK_B200_OPTIONS=$opt python bench.py $small 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f  ms/step %.2f' % (d['value'], d['ms_per_step']), d['roofline']['kernel_ms_per_step'], d['parity']['ok'], d['parity']['max_rel_Sa'])"
  tail -2 gpurun_out/ab.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply_tma -s 3 -c 1 -f -o gpurun_out/r2n_prof_k_apply_tma python bench.py --nx 300 --ny 300 --nobs 90000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2n_ncu.log 2>&1
