N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_shard_n$N.json 2> gpurun_out/r02_bench_shard_n$N.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r02_bench_shard_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', d['ms_per_step'], 'it/s', d['value'], 'e2e', d['e2e']['value'], d['scaling'])
print(d['roofline']['phase_ms'])
print(d.get('parity_vs_single_gpu',{}).get('pass'), d.get('restarts',{}).get('value'))
P
