# usage: bash tools/gpu_ncu.sh <kernel-regex> <out-name> [launch-skip] [launch-count] [extra bench args...]
K=$1; O=$2; S=${3:-8}; C=${4:-4}; shift 4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $S --launch-count $C -o gpurun_out/$O -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e "$@" > gpurun_out/$O.log 2>&1
tail -3 gpurun_out/$O.log
