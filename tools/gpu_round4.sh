#!/bin/bash
TAG=${1:-r01e}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_$TAG.log
for lib in libpp2d.so libpp2d_ring2.so libpp2d_ring6.so; do
echo "== variants $lib"; PP2D_LIB=$PWD/path_planning_2d_b200/$lib python tools/sweep_variants.py 4096 quick 2>&1 | tee $OUT/variants_${TAG}_$lib.log
done
echo "== ncu"
python tools/ncu_target.py 4096 12 > $OUT/ncu_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mdp_sweep_kernel -s 3 -c 1 -f -o $OUT/prof_$TAG python tools/ncu_target.py 4096 12 > $OUT/ncu_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu_$TAG.log
