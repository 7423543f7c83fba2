small="--nx 300 --ny 300 --nz 30 --nobs 90000 --steps 3 --warmup 2 --no-cpu --no-e2e"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -3
for tw in 1; do
  echo "== TRI_WARP $tw"
  OAK_B200_TRI_WARP=$tw python bench.py $small 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f  ms/step %.2f' % (d['value'], d['ms_per_step']), d['roofline']['kernel_ms_per_step'], d['parity']['ok'], d['parity']['max_rel_Sa'])"
  tail -2 gpurun_out/ab.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tridiag_warp -s 3 -c 1 -f -o gpurun_out/r2j_prof_k_tridiag_warp python bench.py --nx 300 --ny 300 --nobs 90000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2j_ncu.log 2>&1
