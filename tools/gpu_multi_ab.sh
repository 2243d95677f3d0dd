#!/bin/bash
# gpurun --gpus 2: edge-block shortening sweep on the single-process 2-GPU handle + the new test
OUT=gpurun_out; mkdir -p $OUT
for s in 0 20 28; do PP2D_P2P_EDGE_SHORT=$s python tools/time_multi.py 2; done 2>&1 | tee $OUT/multi_ab.log
python tools/time_multi.py 1 2>&1 | tee -a $OUT/multi_ab.log
for s in 0 20; do PP2D_P2P_EDGE_SHORT=$s python tools/time_multi.py 2 16384; done 2>&1 | tee -a $OUT/multi_ab.log
timeout 900 python -m pytest tests/test_distributed_gpu.py -q -m gpu --timeout=600 2>&1 | tail -3 | tee -a $OUT/multi_ab.log
