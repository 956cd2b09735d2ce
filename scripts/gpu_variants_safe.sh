#!/bin/bash
# As gpu_variants.sh, with short timeouts (first runs of builds that could dead-lock on a barrier).
OUT=gpurun_out/${1:-variants}; mkdir -p $OUT
STEPS=${2:-20}
for lib in robopianist_b200/variants/lib_*.so; do
  n=$(basename $lib .so)
  ROBOPLAY_LIB=$PWD/$lib timeout -k 5 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/$n.smoke 2>&1; rc=$?; echo "$n smoke rc=$rc: $(tail -1 $OUT/$n.smoke)"
  if [ $rc -ne 0 ]; then nvidia-smi --query-gpu=name,utilization.gpu --format=csv,noheader; continue; fi
  ROBOPLAY_LIB=$PWD/$lib timeout -k 5 240 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/$n.pytest 2>&1; echo "$n pytest rc=$?: $(tail -1 $OUT/$n.pytest)"
  ROBOPLAY_LIB=$PWD/$lib timeout -k 5 180 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > $OUT/$n.json 2> $OUT/$n.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$n.json")); print("$n", "ms/step %.2f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d["solver"])
except Exception as e: print("$n failed", e)
PY
done
