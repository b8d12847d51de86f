set -x
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/plain40.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu40a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"grad3_kernel|kernmat2_kernel" -s 4 -c 2 -o gpurun_out/r02_pair_kernels_v3 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu40b.log 2>&1
tail -2 gpurun_out/ncu40b.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; echo "ref rc=$?"
