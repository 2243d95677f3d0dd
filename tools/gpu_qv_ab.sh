#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for g in 2 3 4; do echo "== groups $g"; PP2D_POMDP_GROUPS=$g python tools/bench_pomdp.py 1250 2>&1 | tail -1; done | tee $OUT/qv_groups.log
echo "== 2500 queries"; for g in 2 3; do PP2D_POMDP_GROUPS=$g python tools/bench_pomdp.py 2500 2>&1 | tail -1; done | tee -a $OUT/qv_groups.log
