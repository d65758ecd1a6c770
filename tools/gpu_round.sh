set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1c_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r1c_bench_mgo216.json 2> gpurun_out/r1c_bench_err.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1c_bench_ref.json 2>> gpurun_out/r1c_bench_err.log
timeout 600 python bench.py --workload au992 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_bench_au992.json 2>> gpurun_out/r1c_bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1c_ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_plane|k_zcol|k_fnl$|k_back|k_gemm' --launch-skip 12 --launch-count 8 -o gpurun_out/r1c_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1c_ncu_full.log 2>&1
cat gpurun_out/r1c_pytest.log; cat gpurun_out/r1c_bench_mgo216.json
