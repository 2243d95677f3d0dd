#!/bin/bash
# Round-end evidence on one B200: tests, bench (both arms), ncu launch lists and
# one full capture of the fused sweep kernel and of the QV-tree values kernel.
TAG=${1:-r02z}
OUT=gpurun_out; mkdir -p $OUT
echo "== smoke"; python __graft_entry__.py smoke 2>&1 | tail -1
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > $OUT/pytest_$TAG.log 2>&1; echo "exit $?"; tail -3 $OUT/pytest_$TAG.log
echo "== bench reference arm"; python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "exit $?"
echo "== bench"; python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "exit $?"; cut -c1-400 $OUT/bench_$TAG.json
echo "== ncu launch list (MDP step)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-ref-cuda --no-qv --no-syn16k > $OUT/ncu_launches_$TAG.log 2>&1; echo "exit $?"
echo "== ncu launch list (QV-tree batch, fixture alphas: the solver launches are skipped)"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pomdp_ -c 600 --csv --log-file $OUT/pomdp_launches_$TAG.csv \
  python tools/bench_pomdp.py 1250 --fixture > $OUT/pomdp_ncu_$TAG.log 2>&1; echo "exit $?"
echo "== ncu full: fused sweep kernel"
ncu --set full --import-source on --clock-control none -k regex:mdp_sweep_kernel -s 3 -c 1 -o $OUT/prof_fused_$TAG -f \
  python tools/ncu_target.py 4096 12 > $OUT/ncu_full_$TAG.log 2>&1; echo "exit $?"
echo "== ncu full: values kernel"
ncu --set full --import-source on --clock-control none -k regex:pomdp_values -s 40 -c 1 -o $OUT/prof_values_$TAG -f \
  python tools/bench_pomdp.py 1250 --fixture > $OUT/ncu_full_values_$TAG.log 2>&1; echo "exit $?"
echo "== SASS of the fused kernel (instruction histogram + one marching step)"
python tools/sass_summary.py > $OUT/sass_fused_$TAG.txt 2>&1; echo "exit $?"
