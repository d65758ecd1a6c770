# round 2 ncu evidence: launch list of the default bench command (short), one --set full capture of the hot kernels, one of the new ones
set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub > gpurun_out/r2_ncu_launch.log 2>&1
tail -2 gpurun_out/r2_ncu_launch.log
# first step of the timed loop: split_pm, fnl, finish, back, zbwd, plane<HPSI>, zfwd, zbwd, plane<DENSITY>, rho_reduce
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_plane_s|k_zcol|k_fnl|k_back|k_split_pm' --launch-skip 0 --launch-count 9 -f -o gpurun_out/r2_full_step python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-sub > gpurun_out/r2_ncu_a.log 2>&1
tail -2 gpurun_out/r2_ncu_a.log
# the round-2 kernels: E_kin, v(r) producers, PSDA update, Jacobi (first launches only)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_psi2sum|k_ekin_sums|k_vh_|k_psda|k_jac_cols|k_jac_rows|k_rho_expand' --launch-skip 0 --launch-count 14 -f -o gpurun_out/r2_full_new python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sub > gpurun_out/r2_ncu_b.log 2>&1
tail -2 gpurun_out/r2_ncu_b.log
ls -la gpurun_out/r2_full_*.ncu-rep gpurun_out/r2_launches.csv
