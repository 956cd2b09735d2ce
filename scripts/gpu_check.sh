#!/bin/bash
# One GPU session: parity tests, 40-step and 20-step bench, phase timers of the profile build.  usage: gpu_check.sh TAG
TAG=${1:-chk}; OUT=gpurun_out/$TAG; mkdir -p $OUT
(timeout 700 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log)
timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
try:
    d=json.load(open("$OUT/bench.json")); print("bench40 ms/step %.3f value %.0f e2e %.0f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), d.get("solver"))
except Exception as e: print("bench failed", e)
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench20.json 2> $OUT/bench20.err; python - <<PY
import json
try:
    d=json.load(open("$OUT/bench20.json")); print("bench20 ms/step %.3f value %.0f e2e %.0f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), d.get("solver"))
except Exception as e: print("bench failed", e)
PY
if [ -f robopianist_b200/variants/lib_prof.so ]; then
for n in 148 4096; do ROBOPLAY_LIB=$PWD/robopianist_b200/variants/lib_prof.so timeout 200 python tests/gpu_stage_timing.py $n 40 > $OUT/prof_$n.txt 2>&1; done
head -1 $OUT/prof_148.txt; paste <(sed -n 2,22p $OUT/prof_148.txt | cut -c1-58) <(sed -n 23,43p $OUT/prof_148.txt | cut -c32-58)
fi
