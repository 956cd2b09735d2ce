#!/bin/bash
# Quick GPU check: parity tests + short bench.  Usage: bash scripts/gpu_quick.sh tag [extra pytest args]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$OUT/bench.json")); print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["solver"])
except Exception as e: print("bench parse failed", e)
PY
tail -3 $OUT/bench.err
timeout 300 python tests/gpu_timing.py 4096 > $OUT/timing.log 2>&1; cat $OUT/timing.log
