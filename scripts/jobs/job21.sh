set -x
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/plain21.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"grad3_kernel" -s 4 -c 1 -o gpurun_out/r02_grad3_v2 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_grad3b.log 2>&1
tail -3 gpurun_out/ncu_grad3b.log
