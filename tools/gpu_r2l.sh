set -x
QB200_PLANE_H=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mgo216" 2>&1 | tail -3
for h in 0 1; do
  QB200_PLANE_H=$h timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-sub > gpurun_out/r2l_h$h.json 2> gpurun_out/r2l_err.log
  python -c "
import json; d=json.load(open('gpurun_out/r2l_h$h.json')); print('PLANE_H $h', round(d['ms_per_step'],3), d['kernel_ms_per_step'], d['parity']['integrity']['ok'])"
done
