set -x
mkdir -p /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fnl|k_back' --launch-skip 3 --launch-count 3 -f -o /tmp/ncu/t11 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-sub > gpurun_out/t11_ncu.log 2>&1
tail -2 gpurun_out/t11_ncu.log
python tools/ncu_hot.py /tmp/ncu/t11.ncu-rep 0 60 > gpurun_out/t11_hot_fnl.txt 2>&1
python tools/ncu_phases.py /tmp/ncu/t11.ncu-rep 0 > gpurun_out/t11_phases_fnl.txt 2>&1
python tools/ncu_hot.py /tmp/ncu/t11.ncu-rep 4 60 > gpurun_out/t11_hot_back.txt 2>&1
python tools/ncu_phases.py /tmp/ncu/t11.ncu-rep 4 > gpurun_out/t11_phases_back.txt 2>&1
