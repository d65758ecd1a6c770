# first runs of the tensor-memory plane kernel: MgO216-shape parity tests, then a short bench with and without it
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mgo216 or fixture_device" 2>&1 | tail -15 > gpurun_out/t1_pytest.log
cat gpurun_out/t1_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/t1_bench.json 2> gpurun_out/t1_err.log
python -c "
import json; d=json.load(open('gpurun_out/t1_bench.json')); print(d['ms_per_step'], d['value'], d['kernel_ms_per_step'], d.get('parity'))"
tail -3 gpurun_out/t1_err.log
QB200_PLANE_T=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/t1_bench_s.json 2>> gpurun_out/t1_err.log
python -c "
import json; d=json.load(open('gpurun_out/t1_bench_s.json')); print(d['ms_per_step'], d['value'], d['kernel_ms_per_step'])"
