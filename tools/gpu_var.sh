# usage: bash tools/gpu_var.sh "<ENV=VAL ...>" ["<ENV=VAL ...>" ...]   -- bench variants (kernel ms per step) + mgo216 parity under each
rm -f gpurun_out/var.log
for V in "$@"; do
  echo "== $V" >> gpurun_out/var.log
  env $V timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mgo216" 2>&1 | tail -1 >> gpurun_out/var.log
  env $V timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/var_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms_per_step'])" >> gpurun_out/var.log
done
cat gpurun_out/var.log
