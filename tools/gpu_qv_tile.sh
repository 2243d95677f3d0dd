#!/bin/bash
# A/B of the per-tile inner rows and the query ordering of the batched QV-tree
# planner (both exact); then the POMDP parity tests.
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/qv_tile.log; : > $LOG
for cfg in "1 1" "0 0"; do
  set -- $cfg
  echo "== TILE_SUPPORT=$1 SORT=$2" | tee -a $LOG
  for rep in 1 2; do
    PP2D_POMDP_PROFILE=1 PP2D_POMDP_TILE_SUPPORT=$1 PP2D_POMDP_SORT=$2 python tools/bench_pomdp.py 1250 2>&1 | tail -2 | tee -a $LOG
  done
done
echo "== kernel times (tile support + sort on)" | tee -a $LOG
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pomdp_" -s 150 -c 200 --csv python tools/bench_pomdp.py 1250 --fixture 2>/dev/null | python -c "
import sys,csv,collections
rows=list(csv.reader(sys.stdin)); hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); agg[d['Kernel Name'][:36]].append(float(d['Metric Value'])/1e3)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print('   %-36s n=%d avg %.1f us total %.0f us'%(k,len(v),sum(v)/len(v),sum(v)))
" | tee -a $LOG
echo "== pytest pomdp" | tee -a $LOG
timeout 900 python -m pytest tests/test_pomdp_gpu.py tests/test_tree_pin_gpu.py tests/test_pbvi_gpu.py tests/test_checkpoint_gpu.py tests/test_host_mirror.py -q -m gpu -x --timeout=600 2>&1 | tail -5 | tee -a $LOG
