set -x
timeout 900 python -m pytest tests/test_subspace_la.py tests/test_gpu_parity.py -m gpu -x -q -k "psda or scf or ekin or coexist" 2>&1 | tail -8 | tee gpurun_out/r2g_pytest.log
for e in 0 1 2 4 7; do
  QB200_EXP=$e timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-sub > gpurun_out/r2g_exp$e.json 2> gpurun_out/r2g_exp_err.log
  python -c "
import json; d=json.load(open('gpurun_out/r2g_exp$e.json')); print('EXP $e', round(d['ms_per_step'],3), d['kernel_ms_per_step'])"
done
