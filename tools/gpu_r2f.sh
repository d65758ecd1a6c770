set -x
timeout 900 python -m pytest tests/test_subspace_la.py tests/test_gpu_parity.py -m gpu -x -q -k "psda or scf or ekin or coexist" 2>&1 | tail -8 | tee gpurun_out/r2f_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench_err.log
tail -5 gpurun_out/r2f_bench_err.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2f_bench.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e'] and {k: d['e2e'][k] for k in ('value','h2d_bytes_per_step','d2h_bytes_per_step')}, 'host', d['e2e_host_blocks'] and d['e2e_host_blocks']['value'])
print('scf', d.get('scf_iteration'))
print('parity ok', d['parity'].get('ok'), d['parity'].get('hpsi_relerr'))
PY
