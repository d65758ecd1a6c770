# full measurement round: tests, both bench arms, Au992, launch list, ncu full captures of every hot kernel type
R=${1:-r1j}
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${R}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench_mgo216.json 2> gpurun_out/${R}_bench_err.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_ref.json 2>> gpurun_out/${R}_bench_err.log
timeout 900 python bench.py --workload au992 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_bench_au992.json 2>> gpurun_out/${R}_bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_ncu_b.log 2>&1
# matching launches per step: fnl3, back3, 13 x (zbwd, plane<HPSI>, zfwd), 13 x (zbwd, plane<DENSITY>)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_plane_s|k_zcol|k_fnl3|k_back3' --launch-skip 0 --launch-count 5 -f -o gpurun_out/${R}_full_a python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_plane_s' --launch-skip 13 --launch-count 2 -f -o gpurun_out/${R}_full_b python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_ncu_bb.log 2>&1
cat gpurun_out/${R}_pytest.log; cat gpurun_out/${R}_bench_mgo216.json
