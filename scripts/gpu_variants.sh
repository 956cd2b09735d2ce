#!/bin/bash
# Times every library in robopianist_b200/variants/ with the bench workload (no CPU baseline) and the identical-action sweep.
OUT=gpurun_out/${1:-variants}; mkdir -p $OUT
for lib in robopianist_b200/variants/lib_*.so; do
  n=$(basename $lib .so)
  ROBOPLAY_LIB=$PWD/$lib timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > $OUT/$n.json 2> $OUT/$n.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$n.json")); print("$n", "ms/step %.2f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"])
except Exception as e: print("$n failed", e)
PY
done
