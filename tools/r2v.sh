#!/bin/bash
# GPU visit "r2v": A/B of k_tql with tracked block ends (product build = loop nest, variant = flat loop) against the build
# of the previous visit (oak_b200/variants/liboak_pwk0.so: rescanning form, old k_tvec / k_gram_mma details), then the C3 line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=r2v
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -3
small="--nx 300 --ny 300 --nz 30 --nobs 90000 --steps 4 --warmup 2 --no-cpu --no-e2e"
for lib in oak_b200/variants/liboak_pwk0.so oak_b200/liboak_b200.so oak_b200/variants/liboak_flat.so; do
  echo "== $lib"
  OAK_B200_LIB=$PWD/$lib timeout 600 python bench.py $small 2>gpurun_out/${TAG}_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), d['roofline']['kernel_ms_per_step'], d.get('parity',{}).get('ok'))"
done
for lib in oak_b200/liboak_b200.so oak_b200/variants/liboak_flat.so; do
echo "== C3 $lib"
OAK_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e 2>>gpurun_out/${TAG}_ab.err | tee gpurun_out/${TAG}_c3_$(basename $lib).json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f columns/s  ms/step %.2f' % (d['value'], d['ms_per_step']), d['roofline']['kernel_ms_per_step'], d.get('parity',{}).get('ok'))"
done
NCU_KERNELS="k_tql k_tvec" bash tools/r2_check.sh $TAG ncufull
