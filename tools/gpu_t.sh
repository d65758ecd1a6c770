set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_subspace_la.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/t_pytest.log
cat gpurun_out/t_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t_bench.json 2> gpurun_out/t_err.log
python -c "
import json; d=json.load(open('gpurun_out/t_bench.json')); print(d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['kernel_ms_per_step']); print(d['subspace_la']); print(d['roofline_fp64']); print(d['tddft'])"
tail -5 gpurun_out/t_err.log
