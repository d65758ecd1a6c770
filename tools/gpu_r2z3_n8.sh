set -x
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2z3_bench_n$N.json 2> gpurun_out/r2z3_bench_n${N}_err.log
tail -5 gpurun_out/r2z3_bench_n${N}_err.log
python - <<PY
import json
d = json.load(open('gpurun_out/r2z3_bench_n$N.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'host', d['e2e_host_blocks']['value'], d['parity'])
s = d.get('strong_scaling'); print('strong', s and (s.get('ms_per_step'), s.get('value'), s.get('e2e') and s['e2e']['value'], s.get('scf_iteration'), s.get('error')))
a = d.get('au992'); print('au992', a and (a.get('ms_per_step'), a.get('value'), a.get('error')))
PY
