timeout 900 python -m pytest tests/test_gpu_path.py tests/test_gpu_baseline_shapes.py -x -q > gpurun_out/r02_j34_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_j34_pytest.log
for impl in 1 2; do
ACE_KERNMAT_IMPL=$impl python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('impl $impl', d['ms_per_step'], d['roofline']['phase_ms'])
for k,v in d['other_configs'].items(): print(k, round(v['ms_per_iter'],3), {a:round(b,2) for a,b in v['phase_ms'].items()})"
done
