for ns in 0 100 250 500 1000; do
  QB200_EXP=$((ns*256)) timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-sub > gpurun_out/r2m_$ns.json 2> gpurun_out/r2m_err.log
  python -c "
import json; d=json.load(open('gpurun_out/r2m_$ns.json')); print('SKEW $ns', round(d['ms_per_step'],3), d['kernel_ms_per_step']['xy_stage'], d['parity']['integrity']['ok'])"
done
