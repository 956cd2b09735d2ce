#!/bin/bash
# ncu capture of one step-kernel launch (bench workload). Usage: bash scripts/gpu_prof.sh tag [envs]
TAG=${1:-p}
ENVS=${2:-4096}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on --warp-sampling-interval 10 --warp-sampling-buffer-size 536870912 -k regex:rp_step_kernel -s 8 -c 1 -o $OUT/prof -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --envs-per-gpu $ENVS > $OUT/ncu.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu.log
ls -la $OUT
