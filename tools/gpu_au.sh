set -x
timeout 900 python bench.py --workload au992 --nst ${1:-192} --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/au_bench.json 2> gpurun_out/au_err.log
python -c "
import json; d=json.load(open('gpurun_out/au_bench.json')); print(d['ms_per_step'], d['value'], d['kernel_ms_per_step'], d['roofline_hbm'], d['roofline_fp64'], d['shape'])"
tail -5 gpurun_out/au_err.log
