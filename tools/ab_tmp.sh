timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "local_analysis_matches or chunk or resident" 2>&1 | tail -2
for hv in 1 2; do
echo "== eig halves $hv"
OAK_B200_EIG_HALVES=$hv python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('C3 value %.0f  ms/step %.2f' % (d['value'], d['ms_per_step']), d['parity']['ok'], d['gpu_launches'])"
tail -1 gpurun_out/ab.err
done
