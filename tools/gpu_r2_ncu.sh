# round 2 ncu evidence: launch list of the default bench command (short), one --set full capture of the hot kernels, one of the
# new ones; the reports are summarised ON the box (they exceed what gpurun copies back) and only the text comes home
set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub > gpurun_out/r2_ncu_launch.log 2>&1
tail -2 gpurun_out/r2_ncu_launch.log
python tools/launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_mgo216.txt
mkdir -p /tmp/ncu
# first step of the timed loop: split_pm, fnl, finish, back, zbwd, plane<HPSI>, zfwd, zbwd, plane<DENSITY>
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_plane_s|k_zcol|k_fnl|k_back|k_split_pm' --launch-skip 0 --launch-count 9 -f -o /tmp/ncu/r2_full_step python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-sub > gpurun_out/r2_ncu_a.log 2>&1
tail -2 gpurun_out/r2_ncu_a.log
python tools/ncu_summary.py /tmp/ncu/r2_full_step.ncu-rep > gpurun_out/r2_ncu_full_step_summary.txt 2>&1
python tools/ncu_traffic.py /tmp/ncu/r2_full_step.ncu-rep > gpurun_out/r2_ncu_traffic.json 2> gpurun_out/r2_ncu_traffic.err
# the round-2 kernels: E_kin, v(r) producers, PSDA update, Jacobi (first launches only)
timeout 900 ncu --set full --clock-control none -k regex:'k_psi2sum|k_ekin_sums|k_vh_|k_psda|k_jac_cols|k_jac_rows|k_rho_expand' --launch-skip 0 --launch-count 14 -f -o /tmp/ncu/r2_full_new python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sub > gpurun_out/r2_ncu_b.log 2>&1
tail -2 gpurun_out/r2_ncu_b.log
python tools/ncu_summary.py /tmp/ncu/r2_full_new.ncu-rep > gpurun_out/r2_ncu_full_new_kernels_summary.txt 2>&1
ls -la /tmp/ncu gpurun_out | head -30
