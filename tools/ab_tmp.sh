small="--nx 300 --ny 300 --nz 30 --nobs 90000 --steps 3 --warmup 2 --no-cpu --no-e2e"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -3
for lib in "" $EXTRA_LIBS; do
  echo "== lib $lib"
  OAKB200_DEBUG=1 OAK_B200_LIB=${lib:+$PWD/$lib} python bench.py $small $BENCH_ARGS 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f  ms/step %.2f' % (d['value'], d['ms_per_step']), d['roofline']['kernel_ms_per_step'], d['parity']['ok'], d['parity']['max_rel_Sa'])"
  grep "oak_b200" gpurun_out/ab.err | tail -1
done
