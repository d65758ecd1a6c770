set -x
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/t_bench.json 2> gpurun_out/t_err.log
python -c "
import json; d=json.load(open('gpurun_out/t_bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['tddft'])"
tail -3 gpurun_out/t_err.log
# one launch of every hot kernel type: matching launches per step = fnl3, back3, 11 x (zbwd, plane<HPSI>, zfwd), 11 x (zbwd, plane<DENSITY>)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_plane_s|k_zcol|k_fnl3|k_back3' --launch-skip 0 --launch-count 5 -f -o gpurun_out/t_full_a python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/t_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_plane_s|k_zcol|k_fnl3|k_back3' --launch-skip 35 --launch-count 2 -f -o gpurun_out/t_full_b python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/t_ncu_b.log 2>&1
tail -2 gpurun_out/t_ncu_b.log
