#!/bin/bash
# Quick validation on one B200: all GPU tests, a short bench line, the launch list of an e2e-heavy run.
TAG=${1:-r04e}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu --timeout=900 -x > $OUT/pytest_$TAG.log 2>&1; echo "exit $?"; tail -3 $OUT/pytest_$TAG.log
echo "== bench (short)"; python bench.py --steps 20 --warmup 3 --no-cpu --no-qv --no-syn16k --no-ref-cuda > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"
python - <<P
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("  value %.4g  ms %.3f  frac %.3f  e2e %.4g policy-only %.4g"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["e2e"]["policy_only"]["value"]))
P
echo "== ncu launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-ref-cuda --no-qv --no-syn16k > $OUT/ncu_launches_$TAG.log 2>&1; echo "exit $?"
python - <<P
import csv,collections
rows=[r for r in csv.reader(open("$OUT/launches_$TAG.csv")) if len(r)>10]
hdr=rows[0]; k=hdr.index("Kernel Name"); v=hdr.index("Metric Value"); u=hdr.index("Metric Unit")
acc=collections.defaultdict(list)
for r in rows[1:]:
    t=float(r[v].replace(",","")); t = t/1000 if r[u]=="ns" else t
    acc[r[k][:60]].append(t)
for n,l in acc.items(): print("  %-60s n=%3d avg %.1f us"%(n,len(l),sum(l)/len(l)))
P
