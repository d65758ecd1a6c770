N=${1:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_err.log
echo "exit=$?"
grep -v "^\*\|OMP_NUM_THREADS" gpurun_out/n${N}_err.log | tail -30
wc -c gpurun_out/n${N}_bench.json
python -c "
import json; d=json.load(open('gpurun_out/n${N}_bench.json')); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['kernel_ms_per_step'])"
python -c "
import json; d=json.load(open('gpurun_out/n${N}_bench.json')); print(d['subspace_la']); print(d['tddft'])"
