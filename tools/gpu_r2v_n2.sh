# 2 GPUs: the C ABI's own NCCL collectives (tests/test_gpu_comm.py world 2) and the bench at N=2 (weak + strong sub-record + au992)
set -x
timeout 600 python -m pytest tests/test_gpu_comm.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2v_pytest_comm.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2v_bench_n2.json 2> gpurun_out/r2v_bench_n2_err.log
tail -5 gpurun_out/r2v_bench_n2_err.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2v_bench_n2.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['parity'])
s = d.get('strong_scaling'); print('strong', s and (s.get('ms_per_step'), s.get('value'), s.get('e2e') and s['e2e']['value'], s.get('error')))
a = d.get('au992'); print('au992', a and (a.get('ms_per_step'), a.get('value'), a.get('error')))
PY
