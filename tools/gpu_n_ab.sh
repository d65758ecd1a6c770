for f in ${AB:-1 0 1 0}; do
QB200_BENCH_FUSED_ALLREDUCE=$f timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $1 --steps 20 --warmup 5 --no-sub --no-e2e --no-cpu-baseline > gpurun_out/u13.json 2> gpurun_out/u13_err.log
python -c "
import json
d=json.load(open('gpurun_out/u13.json')); print('fused', $f, d['n_gpus'], d['ms_per_step'], d['value'], sum(v for k,v in d['kernel_ms_per_step'].items() if k!='xy_density'))"
done
