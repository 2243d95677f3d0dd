#!/bin/bash
# Batched QV-tree planner: number of pipelined groups, warm-cache kernel times.
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/qv_sweep.log; : > $LOG
for g in 2 3 4; do
  echo "== GROUPS=$g" | tee -a $LOG
  for rep in 1 2 3; do
    PP2D_POMDP_GROUPS=$g python tools/bench_pomdp.py 1250 2>&1 | tail -1 | cut -c1-60 | tee -a $LOG
  done
done
echo "== 2500 queries" | tee -a $LOG
for g in 3 4; do PP2D_POMDP_GROUPS=$g python tools/bench_pomdp.py 2500 2>&1 | tail -1 | cut -c1-60 | tee -a $LOG; done
echo "== kernel times, caches left warm (ncu --cache-control none)" | tee -a $LOG
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"pomdp_" -s 150 -c 200 --csv python tools/bench_pomdp.py 1250 --fixture 2>/dev/null | python -c "
import sys,csv,collections
rows=list(csv.reader(sys.stdin)); hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); agg[d['Kernel Name'][:36]].append(float(d['Metric Value'])/1e3)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print('   %-36s n=%d avg %.1f us total %.0f us'%(k,len(v),sum(v)/len(v),sum(v)))
" | tee -a $LOG
