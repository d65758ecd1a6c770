set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mgo216 or fixture_device" 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench_err.log
tail -5 gpurun_out/r2h_bench_err.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2h_bench.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['kernel_ms_per_step'])
a = d['au992']; print('au992', a.get('error') or (a['ms_per_step'], a['value'], a['kernel_ms_per_step'], a['roofline_local_path']['frac'], a['roofline_fp64']['frac']))
PY
