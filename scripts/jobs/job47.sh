set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_final_pytest.log
tail -6 gpurun_out/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final2_bench_n1.json 2> gpurun_out/r02_final2_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_final2_bench_n1.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'it/s', d['value'], 'e2e', d['e2e']['value'], d['roofline']['phase_ms'], d['parity']['pass'])
P
timeout 300 ncu --set full --clock-control none -k regex:diag_block_kernel -s 20 -c 2 -o gpurun_out/r02_diag_final python scripts/dense_only.py 4096 > gpurun_out/ncu_diag_final.log 2>&1; tail -1 gpurun_out/ncu_diag_final.log
