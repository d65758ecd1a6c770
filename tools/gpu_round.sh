# full measurement round: tests, both bench arms, Au992, si54p, launch list, ncu full captures of every hot kernel type
R=${1:-r1l}
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${R}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench_mgo216.json 2> gpurun_out/${R}_bench_err.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_ref.json 2>> gpurun_out/${R}_bench_err.log
timeout 900 python bench.py --workload au992 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${R}_bench_au992.json 2>> gpurun_out/${R}_bench_err.log
timeout 300 python bench.py --workload si54p --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_si54p.json 2>> gpurun_out/${R}_bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_ncu_b.log 2>&1
# first batch of a step: split_pm, fnl, back, merge_pm, then (zbwd, plane<HPSI>, zfwd); density plane kernels follow the 13 H psi batches
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_plane_s|k_zcol|k_fnl|k_back|k_split_pm|k_merge_pm' --launch-skip 0 --launch-count 8 -f -o gpurun_out/${R}_full_a python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_plane_s' --launch-skip 13 --launch-count 2 -f -o gpurun_out/${R}_full_b python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_ncu_bb.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_pack_w3|k_fnl3|k_back3|k_la_finish|k_potrf_trail|k_trtri' --launch-skip 0 --launch-count 8 -f -o gpurun_out/${R}_full_la python tools/la_profile.py > gpurun_out/${R}_ncu_la.log 2>&1
cat gpurun_out/${R}_pytest.log; cut -c1-1500 gpurun_out/${R}_bench_mgo216.json; tail -3 gpurun_out/${R}_ncu_la.log
