#!/bin/bash
# First GPU session of the next round: confirm the shipped build, then measure the prepared experiments.
#   here (no GPU):  scripts/build_variants.sh base:"" ldl:"-DRP_MINV_LDL=1" \
#                       pgsbar:"-DRP_PHASE_EVERY=100 -DRP_PHASE_MASK=2" ldlpgsbar:"-DRP_MINV_LDL=1 -DRP_PHASE_EVERY=100 -DRP_PHASE_MASK=2"
#   (ldl: M^-1 through the sparse factorisation, prepared and host-checked, never run on a GPU;
#    pgsbar: the phase barrier moved from the top of the substep to the entry of the solver — untested idea:
#    fast environments would overlap their next setup with the slow ones' sweeps)
#   on the GPU box: bash scripts/gpu_next.sh [tag]
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q -s > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|SoA Gauss" $OUT/pytest.log | tail -4
echo "== bench"; timeout 600 python bench.py --steps 40 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/bench.json
echo "== variants (parity + 20-step bench each)"; bash scripts/gpu_variants_safe.sh $TAG/variants 20
echo "== ncu launch list + full capture of the shipped build"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rp_step_kernel -s 12 -c 1 -o $OUT/prof -f python bench.py --steps 12 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?"
