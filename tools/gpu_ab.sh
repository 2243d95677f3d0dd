#!/bin/bash
# A/B timing of sweep-kernel builds on the GPU box: every lib in build/variants plus the default,
# with and without programmatic dependent launch, 4096^2 and 16384^2; the J checksum must not change.
OUT=gpurun_out
TAG=${1:-ab}
mkdir -p $OUT
LOG=$OUT/ab_$TAG.log
: > $LOG
for size in 4096 16384; do
  for pdl in 1 0; do
    echo "# size $size PP2D_MDP_PDL=$pdl" >> $LOG
    PP2D_MDP_PDL=$pdl python tools/time_fused.py $size >> $LOG 2>&1
    for f in build/variants/libpp2d_*.so; do
      [ -f "$f" ] && PP2D_MDP_PDL=$pdl PP2D_LIB=$PWD/$f python tools/time_fused.py $size >> $LOG 2>&1
    done
  done
done
cat $LOG
