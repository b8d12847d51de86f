set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1.json 2> gpurun_out/r02_bench_1.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/r02_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_kernel -s 819 -c 3 -o gpurun_out/r02_dgemm python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_bench_1.err | tail -5
