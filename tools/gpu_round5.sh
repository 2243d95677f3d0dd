#!/bin/bash
TAG=${1:-r01f}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_$TAG.log
echo "== variants"; python tools/sweep_variants.py 4096 quick 2>&1 | tee $OUT/variants_$TAG.log
echo "== bench"; python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
