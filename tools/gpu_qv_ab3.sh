#!/bin/bash
# A/B of one environment switch of the batched QV-tree planner + warm kernel times + ncu full of two kernels.
# usage: bash tools/gpu_qv_ab3.sh VAR
OUT=gpurun_out; mkdir -p $OUT
VAR=${1:-PP2D_POMDP_CSUM_TILE}
LOG=$OUT/qv_ab3.log; : > $LOG
for v in 1 0 1 0; do
  echo "== $VAR=$v" | tee -a $LOG
  env $VAR=$v python tools/bench_pomdp.py 1250 2>&1 | tail -1 | cut -c1-60 | tee -a $LOG
done
echo "== kernel times" | tee -a $LOG
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"pomdp_" -s 150 -c 200 --csv python tools/bench_pomdp.py 1250 --fixture 2>/dev/null | python -c "
import sys,csv,collections
rows=list(csv.reader(sys.stdin)); hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); agg[d['Kernel Name'][:36]].append(float(d['Metric Value'])/1e3)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print('   %-36s n=%d avg %.1f us total %.0f us'%(k,len(v),sum(v)/len(v),sum(v)))
" | tee -a $LOG
ncu --set full --import-source on --clock-control none --cache-control none -k regex:"pomdp_expand|pomdp_child_write|pomdp_child_sum|pomdp_values" -s 120 -c 4 -o $OUT/prof_qv_small -f python tools/bench_pomdp.py 1250 --fixture > $OUT/ncu_qv_small.log 2>&1; echo "ncu exit $?"
echo "== pytest" | tee -a $LOG
timeout 900 python -m pytest tests/test_pomdp_gpu.py tests/test_tree_pin_gpu.py -q -m gpu -x --timeout=600 2>&1 | tail -3 | tee -a $LOG
