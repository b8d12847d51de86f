set -x
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_shard_n$N.json 2> gpurun_out/r02_bench_shard_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_shard_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02_bench_shard_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', d['ms_per_step'], 'it/s', d['value'], 'e2e', d['e2e']['value'], d['scaling'])
print(d['roofline']['phase_ms'])
print(d.get('parity_vs_single_gpu'))
print(d.get('restarts'))
P
ACE_SHARD_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/shard_check.py C3 > gpurun_out/r02_shard_check_C3_w$N.log 2> gpurun_out/r02_shard_trace_w$N.log; echo "check rc=$?"
tail -2 gpurun_out/r02_shard_check_C3_w$N.log | cut -c1-1200
