set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/z2_pytest.log
cat gpurun_out/z2_pytest.log
run() { echo "== $*" >> gpurun_out/z2_variants.log; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/z2_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms_per_step'])" >> gpurun_out/z2_variants.log; }
rm -f gpurun_out/z2_variants.log
run QB200_Z2=0
run QB200_Z2=1
run QB200_ZF_COLS=32
run QB200_ZF_COLS=29
run QB200_ZF_COLS=24
run QB200_ZF_COLS=16
run QB200_ZB_COLS=32
run QB200_ZB_COLS=24
run QB200_ZB_COLS=16
cat gpurun_out/z2_variants.log
