#!/bin/bash
# A/B of alternate builds (build/variants/*.so) of the batched QV-tree planner against the default
# library, then the POMDP parity tests on every variant.
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/qv_lib_ab.log; : > $LOG
for rep in 1 2 3; do
  for lib in path_planning_2d_b200/libpp2d.so build/variants/libpp2d_*.so; do
    [ -f "$lib" ] || continue
    echo "== $lib" | tee -a $LOG
    PP2D_LIB=$PWD/$lib python tools/bench_pomdp.py 1250 2>&1 | tail -1 | cut -c1-118 | tee -a $LOG
  done
done
for lib in build/variants/libpp2d_*.so; do
  [ -f "$lib" ] || continue
  echo "== pytest with $lib" | tee -a $LOG
  PP2D_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_pomdp_gpu.py tests/test_tree_pin_gpu.py -q -m gpu -x --timeout=600 2>&1 | tail -3 | tee -a $LOG
done
