# ncu --set full of the two k_plane_t launches of one step; summaries computed on the box
set -x
mkdir -p /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_plane_t' --launch-skip 2 --launch-count 2 -f -o /tmp/ncu/t2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-sub > gpurun_out/t2_ncu.log 2>&1
tail -2 gpurun_out/t2_ncu.log
python tools/ncu_summary.py /tmp/ncu/t2.ncu-rep > gpurun_out/t2_summary.txt 2>&1
python tools/ncu_phases.py /tmp/ncu/t2.ncu-rep 0 > gpurun_out/t2_phases_hpsi.txt 2>&1
python tools/ncu_phases.py /tmp/ncu/t2.ncu-rep 2 > gpurun_out/t2_phases_density.txt 2>&1
python tools/ncu_hot.py /tmp/ncu/t2.ncu-rep 0 40 > gpurun_out/t2_hot_hpsi.txt 2>&1
python tools/ncu_hot.py /tmp/ncu/t2.ncu-rep 2 30 > gpurun_out/t2_hot_density.txt 2>&1
ls -la /tmp/ncu
