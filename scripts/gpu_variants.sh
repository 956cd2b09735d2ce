#!/bin/bash
# Every library in robopianist_b200/variants/: parity tests (ROBOPLAY_LIB selects the build) + bench timing.
OUT=gpurun_out/${1:-variants}; mkdir -p $OUT
STEPS=${2:-20}
for lib in robopianist_b200/variants/lib_*.so; do
  n=$(basename $lib .so)
  ROBOPLAY_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/$n.pytest 2>&1; echo "$n pytest: $(tail -1 $OUT/$n.pytest)"
  ROBOPLAY_LIB=$PWD/$lib timeout 300 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > $OUT/$n.json 2> $OUT/$n.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$n.json")); print("$n", "ms/step %.2f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"])
except Exception as e: print("$n failed", e)
PY
done
