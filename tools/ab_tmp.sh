timeout 900 python bench.py --config c5 --steps 3 --warmup 3 > gpurun_out/r2p_c5.json 2> gpurun_out/r2p_c5.err; echo "c5 exit $?"; tail -3 gpurun_out/r2p_c5.err; cut -c1-1500 gpurun_out/r2p_c5.json
timeout 900 python bench.py --config c4 --max-slabs 2 --warmup 2 > gpurun_out/r2p_c4.json 2> gpurun_out/r2p_c4.err; echo "c4 exit $?"; tail -3 gpurun_out/r2p_c4.err; cut -c1-2500 gpurun_out/r2p_c4.json
