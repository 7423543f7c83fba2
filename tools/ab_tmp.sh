small="--nx 300 --ny 300 --nz 30 --nobs 90000 --steps 3 --warmup 2 --no-cpu --no-e2e"
for lib in "" oak_b200/variants/liboak_tv5.so oak_b200/variants/liboak_tw10.so; do
  echo "== lib $lib"
  OAK_B200_LIB=${lib:+$PWD/$lib} python bench.py $small 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f  ms/step %.2f' % (d['value'], d['ms_per_step']), d['roofline']['kernel_ms_per_step'], d['parity']['ok'])"
  tail -1 gpurun_out/ab.err
done
