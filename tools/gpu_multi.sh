#!/bin/bash
# usage: bash tools/gpu_multi.sh N tag
N=${1:-2}; TAG=${2:-r01m}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -8
echo "== nccl test"; timeout 600 python -m pytest tests/test_distributed_gpu.py -q --timeout=500 > $OUT/pytest_multi_$TAG.log 2>&1; echo "exit $?"; tail -5 $OUT/pytest_multi_$TAG.log
for n in 1 $N; do
echo "== bench N=$n"
if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu --no-ref-cuda > $OUT/bench_${TAG}_n$n.json 2> $OUT/bench_${TAG}_n$n.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 > $OUT/bench_${TAG}_n$n.json 2> $OUT/bench_${TAG}_n$n.err; fi
echo "exit $?"; cat $OUT/bench_${TAG}_n$n.json | cut -c1-600; tail -3 $OUT/bench_${TAG}_n$n.err
done
