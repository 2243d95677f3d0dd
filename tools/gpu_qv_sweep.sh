#!/bin/bash
# Batched QV-tree planner: pipelined groups, host threads, batch size (sensitivity of plans/s).
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/qv_sweep.log; : > $LOG
nproc | tee -a $LOG
run() { echo "== $*" | tee -a $LOG; for rep in 1 2; do env "$@" python tools/bench_pomdp.py ${N:-1250} --fixture 2>&1 | tail -1 | cut -c1-64 | tee -a $LOG; done; }
run PP2D_POMDP_GROUPS=2
run PP2D_POMDP_GROUPS=3
run PP2D_POMDP_GROUPS=4
run PP2D_HOST_THREADS=4
run PP2D_HOST_THREADS=8
run PP2D_HOST_THREADS=32
N=2500 run PP2D_POMDP_GROUPS=3
N=2500 run PP2D_POMDP_GROUPS=4
N=5000 run PP2D_POMDP_GROUPS=4
echo "== phases" | tee -a $LOG
PP2D_POMDP_PROFILE=1 python tools/bench_pomdp.py 1250 --fixture 2>&1 | tail -3 | tee -a $LOG
PP2D_POMDP_PROFILE=1 PP2D_HOST_THREADS=4 python tools/bench_pomdp.py 1250 --fixture 2>&1 | tail -3 | tee -a $LOG
