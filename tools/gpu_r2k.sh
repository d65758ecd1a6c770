set -x
for t in 256 128; do
  QB200_Z_THREADS=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-sub > gpurun_out/r2k_z$t.json 2> gpurun_out/r2k_err.log
  python -c "
import json; d=json.load(open('gpurun_out/r2k_z$t.json')); print('ZTHREADS $t', round(d['ms_per_step'],3), d['kernel_ms_per_step'], d['parity']['integrity']['ok'])"
done
QB200_Z_THREADS=128 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mgo216" 2>&1 | tail -3
