timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host or tag or block or slice" 2>&1 | tail -3
for ramp in 1 0; do
QB200_HOST_RAMP=$ramp timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_$ramp.json 2> gpurun_out/e2e_err.log
python -c "
import json; d=json.load(open('gpurun_out/e2e_$ramp.json')); print($ramp, d['ms_per_step'], d['value'], d['e2e']['value'])"
done
tail -3 gpurun_out/e2e_err.log
