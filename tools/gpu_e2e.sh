rm -f gpurun_out/e2e_var.log
for V in "$@"; do
  echo "== $V" >> gpurun_out/e2e_var.log
  env $V timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>>gpurun_out/e2e_err.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])" >> gpurun_out/e2e_var.log
done
cat gpurun_out/e2e_var.log
