for ws in 0 512 1024 2048 4096; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --workspace-mb $ws > gpurun_out/ws_$ws.json 2> gpurun_out/ws_err.log
python -c "
import json; d=json.load(open('gpurun_out/ws_$ws.json')); print($ws, d['ms_per_step'], d['value'], d['kernel_ms_per_step'], d['shape']['states_per_batch'])"
done
tail -3 gpurun_out/ws_err.log
