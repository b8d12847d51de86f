N=$1
ACE_SHARD_HOSTTIME=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/shard_check.py C3 > gpurun_out/r02_hosttime_w$N.log 2> gpurun_out/r02_hosttime_w$N.err; echo "check rc=$?"
grep "host enqueue" gpurun_out/r02_hosttime_w$N.err | tail -6
