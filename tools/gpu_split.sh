set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/sp_pytest.log
cat gpurun_out/sp_pytest.log
run() { echo "== $*" >> gpurun_out/sp_variants.log; env "$@" timeout 600 python bench.py --workload au992 --nst 48 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>>gpurun_out/sp_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms_per_step'], d['enl'], d['shape']['states_per_batch'])" >> gpurun_out/sp_variants.log; }
rm -f gpurun_out/sp_variants.log
run QB200_SPLIT2=0
run QB200_SPLIT2=1
cat gpurun_out/sp_variants.log
tail -3 gpurun_out/sp_err.log
