#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/qv_check.log
for lib in path_planning_2d_b200/libpp2d.so build/variants/libpp2d_cs32.so build/variants/libpp2d_cs8.so; do
  echo "== $lib" | tee -a $OUT/qv_check.log
  PP2D_LIB=$PWD/$lib python tools/bench_pomdp.py 1250 2>&1 | tail -1 | tee -a $OUT/qv_check.log
  PP2D_LIB=$PWD/$lib python tools/bench_pomdp.py 1250 2>&1 | tail -1 | tee -a $OUT/qv_check.log
  PP2D_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"child_sum|child_write|predict|prefix|rewards" -s 100 -c 60 --csv python tools/bench_pomdp.py 1250 --fixture 2>/dev/null | python -c "
import sys,csv,collections
rows=list(csv.reader(sys.stdin)); hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); agg[d['Kernel Name'][:28]].append(float(d['Metric Value'])/1e3)
for k,v in agg.items(): print('   %-28s n=%d avg %.1f us'%(k,len(v),sum(v)/len(v)))
" | tee -a $OUT/qv_check.log
done
