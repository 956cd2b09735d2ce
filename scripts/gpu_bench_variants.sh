#!/bin/bash
# Smoke + 40-step bench of every library in robopianist_b200/variants/ (skips lib_prof).  usage: gpu_bench_variants.sh TAG [steps] [warmup]
OUT=gpurun_out/${1:-variants}; mkdir -p $OUT
STEPS=${2:-40}; WARM=${3:-3}
for lib in robopianist_b200/libroboplay.so robopianist_b200/variants/lib_*.so; do
  n=$(basename $lib .so); [ "$n" = "lib_prof" ] && continue
  ROBOPLAY_LIB=$PWD/$lib timeout -k 5 90 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/$n.smoke 2>&1; rc=$?
  if [ $rc -ne 0 ]; then echo "$n smoke rc=$rc: $(tail -1 $OUT/$n.smoke)"; continue; fi
  ROBOPLAY_LIB=$PWD/$lib timeout -k 5 180 python bench.py --steps $STEPS --warmup $WARM --no-cpu-baseline > $OUT/$n.json 2> $OUT/$n.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$n.json")); print("%-16s ms/step %.3f value %.0f e2e %.0f" % ("$n", d["ms_per_step"], d["value"], d["e2e"]["value"]), d.get("solver"))
except Exception as e: print("$n failed", e)
PY
done
