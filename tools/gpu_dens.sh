set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixture_device or mgo216 or density or shard" 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/d_bench.json 2> gpurun_out/d_err.log
python -c "
import json; d=json.load(open('gpurun_out/d_bench.json')); print(d['ms_per_step'], d['value'], d['kernel_ms_per_step'])"
tail -3 gpurun_out/d_err.log
