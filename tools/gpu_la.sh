# subspace LA (qb200_residual / qb200_gram): parity tests, then the MgO216 bench line with the subspace_la section
set -x
timeout 600 python -m pytest tests/test_subspace_la.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/la_pytest.log
cat gpurun_out/la_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/la_bench.json 2> gpurun_out/la_err.log
python -c "
import json; d=json.load(open('gpurun_out/la_bench.json')); print(d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['kernel_ms_per_step']); print(d['subspace_la'])"
tail -5 gpurun_out/la_err.log
