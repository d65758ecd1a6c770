# round 2, first GPU pass: new parity tests (benchmark projector regime, chunked many-row projectors, bulkal) + the new bench line
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench_err.log
tail -5 gpurun_out/r2a_bench_err.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2a_bench.json'))
print(d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['kernel_ms_per_step'])
print('parity', d.get('parity'))
print('fp64', d['roofline_fp64'].get('fp64_tflops_measured'), d['roofline_fp64']['frac'])
print('au992', json.dumps(d.get('au992'))[:1500])
PY
