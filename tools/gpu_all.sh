# all GPU tests + the default bench line
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/all_pytest.log
cat gpurun_out/all_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/all_bench.json 2> gpurun_out/all_err.log
python -c "
import json; d=json.load(open('gpurun_out/all_bench.json')); print(d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['kernel_ms_per_step']); print(d['subspace_la']); print(d['tddft'])"
tail -5 gpurun_out/all_err.log
