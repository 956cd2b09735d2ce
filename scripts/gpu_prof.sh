#!/bin/bash
# ncu evidence of one GPU session: launch list of a short bench run + one full capture of the step kernel (bench workload,
# window protocol).  Usage: bash scripts/gpu_prof.sh TAG [envs]
TAG=${1:-p}
ENVS=${2:-4096}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-full-episode --no-hand-collisions --envs-per-gpu $ENVS > $OUT/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --warp-sampling-interval 10 --warp-sampling-buffer-size 536870912 -k regex:rp_step_kernel -s 8 -c 1 -o $OUT/prof -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-full-episode --no-hand-collisions --envs-per-gpu $ENVS > $OUT/ncu.log 2>&1; echo "full capture rc=$?"; tail -2 $OUT/ncu.log
timeout 300 python bench.py --steps 40 --warmup 5 > $OUT/bench40.json 2> $OUT/bench40.err
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench20.json 2> $OUT/bench20.err
ls -la $OUT
