#!/bin/bash
# Why is the solve slower at N >= 2 while a download is in flight?  e2e phase trace with the P2P hand-shake switched off piece by piece
# (PP2D_P2P_DEBUG bit0 no waits, bit1 no peer stores, bit2 no fences/signals: results are WRONG, timing only).
N=${1:-2}; TAG=${2:-r04g}; OUT=gpurun_out; mkdir -p $OUT
export PP2D_E2E_TRACE=1
for dbg in 0 5 7; do
  export PP2D_P2P_DEBUG=$dbg
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-qv --no-syn16k --no-ref-cuda > $OUT/bench_dbg${dbg}_${TAG}_n$N.json 2> $OUT/bench_dbg${dbg}_${TAG}_n$N.err
  echo "dbg=$dbg exit $?"; grep -h "rank 0 e2e ms" $OUT/bench_dbg${dbg}_${TAG}_n$N.err | cut -c1-400; tail -2 $OUT/bench_dbg${dbg}_${TAG}_n$N.err | cut -c1-300
done
