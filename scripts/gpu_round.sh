#!/bin/bash
# One GPU-box session: parity tests, bench, launch list, full ncu capture.  Outputs -> gpurun_out/.
# Usage (via gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -5 $OUT/pytest.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log; tail -3 $OUT/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 40 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== timing sweep" ; timeout 600 python tests/gpu_timing.py 148 1184 2368 4096 8192 > $OUT/timing.log 2>&1; cat $OUT/timing.log
echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "rc=$?"; tail -5 $OUT/launches.csv
echo "== ncu full" ; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rp_step_kernel -s 12 -c 1 -o $OUT/prof -f python bench.py --steps 12 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_full.log
ls -la $OUT
