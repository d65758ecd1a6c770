set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/t5_launches_au.csv python bench.py --workload au992 --nst 32 --steps 1 --warmup 1 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/t5_ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/t5_launches_au.csv > gpurun_out/t5_launches_au992.txt
cat gpurun_out/t5_launches_au992.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/t5_launches_si.csv python bench.py --workload si54p --steps 1 --warmup 1 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/t5_ncu_launch2.log 2>&1
python tools/launch_summary.py gpurun_out/t5_launches_si.csv > gpurun_out/t5_launches_si54p.txt
cat gpurun_out/t5_launches_si54p.txt
mkdir -p /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ycols_t|k_xrows2' --launch-skip 3 --launch-count 3 -f -o /tmp/ncu/t5 python bench.py --workload au992 --nst 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-sub > gpurun_out/t5_ncu_a.log 2>&1
tail -2 gpurun_out/t5_ncu_a.log
python tools/ncu_summary.py /tmp/ncu/t5.ncu-rep > gpurun_out/t5_ncu_full_au992_split.txt 2>&1
