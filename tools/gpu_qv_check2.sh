#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; LOG=$OUT/qv_check2.log; : > $LOG
for n in 1250 5000; do
  echo "== $n queries" | tee -a $LOG
  python tools/bench_pomdp.py $n --fixture 2>&1 | tail -1 | cut -c1-200 | tee -a $LOG
done
for t in 4 16; do
  echo "== PP2D_HOST_THREADS=$t" | tee -a $LOG
  for rep in 1 2; do PP2D_POMDP_PROFILE=1 PP2D_HOST_THREADS=$t python tools/bench_pomdp.py 1250 --fixture 2>&1 | tail -2 | cut -c1-260 | tee -a $LOG; done
done
echo "== pytest" | tee -a $LOG
timeout 900 python -m pytest tests/test_pomdp_gpu.py tests/test_tree_pin_gpu.py tests/test_pbvi_gpu.py tests/test_checkpoint_gpu.py tests/test_host_mirror.py -q -m gpu -x --timeout=600 2>&1 | tail -3 | tee -a $LOG
