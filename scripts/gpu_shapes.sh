#!/bin/bash
# Launch-shape variants (robopianist_b200/variants/lib_*.so): smoke, a parity check in the variant's own phase-locked shape,
# the driver window and the full-episode measurement.  usage: gpu_shapes.sh TAG [variant ...]
OUT=gpurun_out/${1:-shapes}; mkdir -p $OUT; shift
LIBS=${@:-$(ls robopianist_b200/variants/lib_*.so | xargs -n1 basename | sed 's/^lib_//; s/\.so$//')}
for n in $LIBS; do
  lib=robopianist_b200/variants/lib_$n.so
  ROBOPLAY_LIB=$PWD/$lib timeout -k 5 90 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/$n.smoke 2>&1; rc=$?; echo "$n smoke rc=$rc: $(tail -1 $OUT/$n.smoke)"
  if [ $rc -ne 0 ]; then nvidia-smi --query-gpu=name,utilization.gpu --format=csv,noheader; continue; fi
  ROBOPLAY_LIB=$PWD/$lib ROBOPLAY_BLOCK_ENVS=$(case $n in e7*) echo 7;; e8*) echo 8;; *) echo 3;; esac) timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -x -q -k "1000_substeps or heterogeneous or shared_key" > $OUT/$n.pytest 2>&1; echo "$n pytest rc=$?: $(tail -1 $OUT/$n.pytest)"
  ROBOPLAY_LIB=$PWD/$lib timeout -k 5 240 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/$n.json 2> $OUT/$n.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$n.json")); fe=d.get("full_episode") or {}
    print("$n", "window ms/step %.2f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "full %.2f" % fe.get("ms_per_step", float("nan")), d.get("solver"))
except Exception as e: print("$n failed", e)
PY
done
