#!/bin/bash
# One GPU-box session: parity tests, golden generation, bench, ncu evidence.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; free -g | head -2 >> $OUT/gpu_$TAG.txt
echo "== smoke"; python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke_$TAG.log
echo "== golden"; python tests/golden/make_golden.py ref $OUT/golden_ref > $OUT/golden_$TAG.log 2>&1; echo "golden exit $?"; tail -12 $OUT/golden_$TAG.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -x --timeout=900 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_$TAG.log
echo "== bench"; python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; BE=$?; echo "bench exit $BE"; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
