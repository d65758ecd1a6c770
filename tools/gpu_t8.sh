set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/t8_smoke.log; cat gpurun_out/t8_smoke.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mgo216 or fixture_device" 2>&1 | tail -4
for cfg in "QB200_T_DENS_SMEM=1" "QB200_T_DENS_SMEM=0"; do
  env $cfg timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/t8.json 2>> gpurun_out/t8_err.log
  python -c "
import json; d=json.load(open('gpurun_out/t8.json')); k=d['kernel_ms_per_step']; print('$cfg', round(d['ms_per_step'],3), 'xy', k['xy_stage'], 'dens', k['xy_density'], 'rho_reduce', k['k_rho_reduce'], d['parity']['integrity']['nelectrons_relerr'], d['roofline_local_path']['frac'])"
done
tail -3 gpurun_out/t8_err.log
