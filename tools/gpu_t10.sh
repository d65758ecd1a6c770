set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "si54p or split" 2>&1 | tail -6
timeout 600 python bench.py --workload si54p --steps 3 --warmup 3 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/t10_si54p.json 2> gpurun_out/t10_err.log
python -c "
import json; d=json.load(open('gpurun_out/t10_si54p.json')); print('si54p', d['ms_per_step'], d['value'], d['kernel_ms_per_step'])"
tail -3 gpurun_out/t10_err.log
