#!/bin/bash
TAG=${1:-r01d}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_$TAG.log
echo "== variants"; python tools/sweep_variants.py 4096 > $OUT/variants_$TAG.log 2>&1; echo "variants exit $?"; cat $OUT/variants_$TAG.log
echo "== ncu"
python tools/ncu_target.py 4096 12 > $OUT/ncu_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mdp_sweep_kernel -s 3 -c 1 -f -o $OUT/prof_$TAG python tools/ncu_target.py 4096 12 > $OUT/ncu_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu_$TAG.log
