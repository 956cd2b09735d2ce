#!/bin/bash
# Final evidence of a round on one GPU: ncu launch list + one full capture of the step kernel, bench lines (driver window with the
# CPU baseline, 40-step window, reference arm), phase timers (window / staggered workloads), BASELINE configs + the
# hand-collision workloads.  usage: gpu_final.sh TAG
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
bash scripts/gpu_prof.sh $TAG
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "rc=$?"
echo "== configs"; timeout 900 python scripts/gpu_configs.py 20 > $OUT/configs.jsonl 2> $OUT/configs.err; echo "rc=$?"; cat $OUT/configs.jsonl; tail -3 $OUT/configs.err
if [ -f robopianist_b200/variants/lib_prof.so ]; then
  ROBOPLAY_LIB=$PWD/robopianist_b200/variants/lib_prof.so timeout 200 python tests/gpu_stage_timing.py 4096 15 > $OUT/phase_timers_4096envs_step15.txt 2>&1
  ROBOPLAY_LIB=$PWD/robopianist_b200/variants/lib_prof.so timeout 200 python tests/gpu_stage_timing_staggered.py 4096 > $OUT/phase_timers_staggered_4096envs.txt 2>&1
  head -3 $OUT/phase_timers_4096envs_step15.txt; head -3 $OUT/phase_timers_staggered_4096envs.txt
fi
python - <<PY
import json
for n in ("bench20", "bench40"):
    try:
        d = json.load(open("$OUT/%s.json" % n)); print(n, d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("cpu_baseline", {}).get("value"), (d.get("full_episode") or {}).get("ms_per_step"))
    except Exception as e: print(n, "failed", e)
PY
echo "== contact-record overflow rate"; timeout 600 python scripts/gpu_overflow_rate.py 4096 200 > $OUT/overflow.txt 2>&1; cat $OUT/overflow.txt
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
