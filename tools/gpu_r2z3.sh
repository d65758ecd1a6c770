# round 2, final state (pair density, ultrasoft remainder, update_twnl, skewed GEMM warp tiles): full GPU suite, default bench,
# launch list, ncu --set full of one step, si54p bench
set -x
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2z3_pytest.log
cat gpurun_out/r2z3_pytest.log
timeout 900 python bench.py > gpurun_out/r2z3_bench.json 2> gpurun_out/r2z3_bench_err.log
tail -3 gpurun_out/r2z3_bench_err.log
python -c "
import json; d=json.load(open('gpurun_out/r2z3_bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernel_ms_per_step'], d['roofline']['frac'], d['roofline_local_path']['frac'], d['cpu_baseline']['value'], d['au992']['ms_per_step'], d.get('strong_scaling',{}).get('ms_per_step'))"
timeout 300 python bench.py --workload si54p --steps 5 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/r2z3_bench_si54p.json 2> gpurun_out/r2z3_si_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z3_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub --no-e2e > gpurun_out/r2z3_ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/r2z3_launches.csv > gpurun_out/r2z3_launches_mgo216.txt
cat gpurun_out/r2z3_launches_mgo216.txt
mkdir -p /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_plane_|k_zcol|k_fnl|k_back|k_split_pm' --launch-skip 9 --launch-count 9 -f -o /tmp/ncu/r2z3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-sub > gpurun_out/r2z3_ncu_a.log 2>&1
tail -2 gpurun_out/r2z3_ncu_a.log
python tools/ncu_summary.py /tmp/ncu/r2z3.ncu-rep > gpurun_out/r2z3_ncu_full_step_summary.txt 2>&1
python tools/ncu_traffic.py /tmp/ncu/r2z3.ncu-rep > gpurun_out/r2z3_ncu_traffic.json 2> gpurun_out/r2z3_ncu_traffic.err
