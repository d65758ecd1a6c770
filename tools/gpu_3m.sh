set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/m3_pytest.log
QB200_NL_TILE=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> gpurun_out/m3_pytest.log
cat gpurun_out/m3_pytest.log
run() { echo "== $*" >> gpurun_out/m3_variants.log; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/m3_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms_per_step'], d['enl'])" >> gpurun_out/m3_variants.log; }
rm -f gpurun_out/m3_variants.log
run QB200_NL_TILE=0
run QB200_NL_TILE=1
cat gpurun_out/m3_variants.log
tail -3 gpurun_out/m3_err.log
