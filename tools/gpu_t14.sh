set -x
for cfg in "QB200_T_DENS=0" "QB200_T_DENS=1" "QB200_T_HPSI=1"; do
  env $cfg timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/t14.json 2>> gpurun_out/t14_err.log
  python -c "
import json; d=json.load(open('gpurun_out/t14.json')); k=d['kernel_ms_per_step']; print('$cfg', round(d['ms_per_step'],3), 'xy', k['xy_stage'], 'hpsi', round(k['xy_stage']-k['xy_density'],3), 'dens', k['xy_density'])"
done
