#!/bin/bash
# A/B timing of sweep-kernel builds (default library + build/variants/*), PDL on, 4096^2 and 16384^2;
# the J checksum must not change.  usage: bash tools/gpu_ab2.sh <tag>
OUT=gpurun_out; TAG=${1:-ab2}; mkdir -p $OUT; LOG=$OUT/ab_$TAG.log; : > $LOG
for size in 4096 16384; do
  echo "# size $size" >> $LOG
  for rep in 1 2; do
    python tools/time_fused.py $size >> $LOG 2>&1
    for f in build/variants/libpp2d_*.so; do
      [ -f "$f" ] && PP2D_LIB=$PWD/$f python tools/time_fused.py $size >> $LOG 2>&1
    done
  done
done
cat $LOG
