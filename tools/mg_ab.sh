for zb in 0 3552 7104; do
  export OAK_B200_ZB=$zb; [ $zb = 0 ] && unset OAK_B200_ZB
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --nx 500 --ny 500 --nobs 250000 --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('zb=$zb', 'value %.0f'%d['value'], 'ms %.2f'%d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
