#!/bin/bash
# One GPU-box session of round 2: parity tests, bench, QV-tree timing.
# Usage (under gpurun, from the repo root): bash tools/gpu_r02.sh <tag> [steps...]
TAG=${1:-r02a}
shift
STEPS=${@:-"pytest bench qv fused"}
OUT=gpurun_out
mkdir -p $OUT
for S in $STEPS; do
case $S in
pytest)
  echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu --timeout=1200 --durations=12 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -25 $OUT/pytest_$TAG.log;;
bench)
  echo "== bench"; python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cut -c1-3000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err;;
benchref)
  echo "== bench reference arm"; python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "exit $?"; cut -c1-1500 $OUT/bench_ref_$TAG.json;;
qv)
  echo "== qv"; PP2D_POMDP_PROFILE=1 python tools/bench_pomdp.py 1250 --cpu > $OUT/qv_$TAG.log 2>&1; echo "qv exit $?"; tail -8 $OUT/qv_$TAG.log
  PP2D_POMDP_DENSE=1 python tools/bench_pomdp.py 1250 > $OUT/qv_dense_$TAG.log 2>&1; tail -2 $OUT/qv_dense_$TAG.log;;
fused)
  echo "== fused kernel timing"; python tools/time_fused.py 4096 > $OUT/fused_$TAG.log 2>&1; python tools/time_fused.py 16384 >> $OUT/fused_$TAG.log 2>&1
  for f in build/variants/libpp2d_*.so; do [ -f "$f" ] && PP2D_LIB=$PWD/$f python tools/time_fused.py 4096 >> $OUT/fused_$TAG.log 2>&1; done
  cat $OUT/fused_$TAG.log;;
esac
done
