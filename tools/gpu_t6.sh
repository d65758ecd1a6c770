# warp configurations of k_plane_t
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mgo216" 2>&1 | tail -3
for h in 0 1; do for d in 0 1 2 3; do
  if [ $h = 1 ] && [ $d != 0 ]; then continue; fi
  QB200_T_HPSI=$h QB200_T_DENS=$d timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/t6_$h$d.json 2>> gpurun_out/t6_err.log
  python -c "
import json; d=json.load(open('gpurun_out/t6_$h$d.json')); k=d['kernel_ms_per_step']; print('cfg h=$h d=$d', round(d['ms_per_step'],3), 'xy', k['xy_stage'], 'hpsi', round(k['xy_stage']-k['xy_density'],3), 'dens', k['xy_density'])"
done; done
tail -3 gpurun_out/t6_err.log
