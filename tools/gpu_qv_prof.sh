#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
ncu --set full --import-source on --clock-control none -k regex:pomdp_child_sum -s 20 -c 1 -o $OUT/prof_child_sum -f python tools/bench_pomdp.py 1250 --fixture > $OUT/prof_child_sum.log 2>&1; echo "exit $?"
ncu -i $OUT/prof_child_sum.ncu-rep --page raw --csv > $OUT/prof_child_sum_raw.csv 2>/dev/null
ncu -i $OUT/prof_child_sum.ncu-rep --page source --csv > $OUT/prof_child_sum_src.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/prof_child_sum_raw.csv')))
d=dict(zip(rows[0],rows[-1]))
for k in ['gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_issued.avg.per_cycle_active','smsp__inst_executed.sum','dram__bytes_read.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','sm__cycles_active.avg','sm__cycles_elapsed.avg','smsp__cycles_active.avg']:
    print(k, d.get(k))
st=[(k,float(v.replace(',',''))) for k,v in d.items() if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio')]
for k,v in sorted(st,key=lambda x:-x[1])[:8]: print(k,v)
PY
