set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for cfg in "QB200_ZB_MW=2" "QB200_ZB_MW=4" "QB200_PLANE_G=18" "QB200_PLANE_G=37" "QB200_PLANE_G=12"; do
  env $cfg timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/t7.json 2>> gpurun_out/t7_err.log
  python -c "
import json; d=json.load(open('gpurun_out/t7.json')); k=d['kernel_ms_per_step']; print('$cfg', round(d['ms_per_step'],3), 'zb', k['k_zcol_bwd'], 'zf', k['k_zcol_fwd'], 'xy', k['xy_stage'], 'dens', k['xy_density'])"
done
tail -3 gpurun_out/t7_err.log
