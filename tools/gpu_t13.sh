set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_subspace_la.py tests/test_ultrasoft.py -m gpu -x -q -k "fixture_device or projector or mgo216_all or many_rows or la or ultrasoft or residual or gram or kpoint or bulkal" 2>&1 | tail -4
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/t13.json 2>> gpurun_out/t13_err.log
python -c "
import json; d=json.load(open('gpurun_out/t13.json')); k=d['kernel_ms_per_step']; print(round(d['ms_per_step'],3), k['k_fnl'], k['k_back']); print(d['subspace_la']); print(d['scf_iteration']); a=d.get('au992'); print('au992', a and (a.get('ms_per_step'), a.get('value'), a.get('kernel_ms_per_step')))"
tail -3 gpurun_out/t13_err.log
