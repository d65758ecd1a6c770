set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mgo216 or fixture_device" 2>&1 | tail -12
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/t3_bench.json 2> gpurun_out/t3_err.log
python -c "
import json; d=json.load(open('gpurun_out/t3_bench.json')); print(d['ms_per_step'], d['value'], d['kernel_ms_per_step'], d['parity'])"
tail -3 gpurun_out/t3_err.log
