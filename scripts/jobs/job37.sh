N=$1
ACE_SHARD_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/shard_check.py C3 > gpurun_out/r02_shard_check_C3_w$N.log 2> gpurun_out/r02_shard_trace_w$N.log; echo "check rc=$?"
tail -1 gpurun_out/r02_shard_check_C3_w$N.log | cut -c1-300
