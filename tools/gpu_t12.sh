set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_subspace_la.py tests/test_ultrasoft.py -m gpu -x -q -k "fixture_device or projector or mgo216_all or many_rows or la or ultrasoft or residual or gram" 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/t12.json 2>> gpurun_out/t12_err.log
python -c "
import json; d=json.load(open('gpurun_out/t12.json')); k=d['kernel_ms_per_step']; print(round(d['ms_per_step'],3), k, d['roofline_fp64']['frac'])"
tail -3 gpurun_out/t12_err.log
