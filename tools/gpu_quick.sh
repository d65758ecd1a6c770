set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/q_pytest.log
cat gpurun_out/q_pytest.log
timeout 300 python bench.py --workload sih4 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/q_err_sih4.log | cut -c1-700
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/q_bench.json 2> gpurun_out/q_err.log
python -c "
import json; d=json.load(open('gpurun_out/q_bench.json')); print(d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['kernel_ms_per_step'])"
tail -5 gpurun_out/q_err.log
