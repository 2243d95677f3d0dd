#!/bin/bash
N=${1:-2}; TAG=${2:-r04d}; OUT=gpurun_out; mkdir -p $OUT
export PP2D_E2E_TRACE=1
python bench.py --steps 6 --warmup 3 --no-cpu --no-qv --no-syn16k --no-ref-cuda > $OUT/bench_trace_${TAG}_n1.json 2> $OUT/bench_trace_${TAG}_n1.err; echo "n1 exit $?"; grep "e2e ms" $OUT/bench_trace_${TAG}_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-qv --no-syn16k --no-ref-cuda > $OUT/bench_trace_${TAG}_n$N.json 2> $OUT/bench_trace_${TAG}_n$N.err; echo "n$N exit $?"; grep "e2e ms" $OUT/bench_trace_${TAG}_n$N.err
