R=${1:-r1m}
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench_mgo216.json 2> gpurun_out/${R}_bench_err.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_ncu_b.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_split_pm|k_fnl|k_back' --launch-skip 0 --launch-count 3 -f -o gpurun_out/${R}_full_nl python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${R}_ncu_a.log 2>&1
cut -c1-600 gpurun_out/${R}_bench_mgo216.json; tail -2 gpurun_out/${R}_bench_err.log
