set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_j32_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_j32_pytest.log
tail -8 gpurun_out/r02_j32_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_j32_bench.json 2> gpurun_out/r02_j32_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_j32_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_j32_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'it/s', d['value'], 'e2e', d['e2e']['value'])
print(d['roofline']['phase_ms'])
for k,v in d['other_configs'].items(): print(k, round(v['ms_per_iter'],3), {a:round(b,2) for a,b in v['phase_ms'].items()})
print(d['parity']); print(d.get('potrf_only'))
P
