#!/bin/bash
# GPU session: full parity suite, bench, BASELINE configs, tuning variants, instruction-cache counters, reference arm.
TAG=${1:-r01d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -4 $OUT/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 40 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "rc=$?"; cat $OUT/bench_reference.json
echo "== configs"; timeout 900 python scripts/gpu_configs.py 20 > $OUT/configs.jsonl 2> $OUT/configs.err; echo "rc=$?"; cat $OUT/configs.jsonl; tail -3 $OUT/configs.err
echo "== variants"; bash scripts/gpu_variants.sh $TAG/variants 20
echo "== icache counters"; timeout 600 ncu --metrics sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction_lookup_hit.sum,gcc__cache_requests_type_instruction_lookup_miss.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:rp_step_kernel -s 8 -c 1 --csv --log-file $OUT/icache.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/icache.log 2>&1; echo "rc=$?"; tail -9 $OUT/icache.csv
