N=$1
for prio in -2 -5; do
ACE_SHARD_BULK_PRIO=$prio timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --mode shard 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prio $prio', d['ms_per_step'], d['value'], d['roofline']['phase_ms'], d.get('parity_vs_single_gpu',{}).get('pass'))"
done
