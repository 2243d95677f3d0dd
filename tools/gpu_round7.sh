#!/bin/bash
TAG=${1:-r01h}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_$TAG.log
echo "== bench"; python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== bench reference arm"; python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "exit $?"; cat $OUT/bench_ref_$TAG.json
echo "== ncu launches"
python bench.py --steps 2 --warmup 3 --no-cpu --no-ref-cuda > $OUT/ncu_plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-ref-cuda > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu exit $?"; tail -2 $OUT/ncu_launches_$TAG.log
