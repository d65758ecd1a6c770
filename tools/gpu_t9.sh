set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "au992" 2>&1 | tail -3
timeout 900 python bench.py --workload au992 --nst 64 --steps 2 --warmup 1 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/t9_au.json 2>> gpurun_out/t9_err.log
python -c "
import json; d=json.load(open('gpurun_out/t9_au.json')); print('au992 rowb12', d['ms_per_step'], d['value'], d['kernel_ms_per_step'])"
tail -3 gpurun_out/t9_err.log
