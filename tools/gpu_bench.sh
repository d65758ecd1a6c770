set -x
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/b_bench.json 2> gpurun_out/b_err.log
python -c "
import json; d=json.load(open('gpurun_out/b_bench.json')); print(d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['kernel_ms_per_step']); print(d['subspace_la']); print(d['tddft']); print(d['cufft_comparison'])"
tail -5 gpurun_out/b_err.log
