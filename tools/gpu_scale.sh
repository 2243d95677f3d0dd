#!/bin/bash
# usage: bash tools/gpu_scale.sh N tag  -- the distributed tests + the bench at N GPUs
N=${1:-8}; TAG=${2:-r01s}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | wc -l
echo "== nccl/p2p test"; timeout 600 python -m pytest tests/test_distributed_gpu.py -q --timeout=500 > $OUT/pytest_multi_$TAG.log 2>&1; echo "exit $?"; tail -3 $OUT/pytest_multi_$TAG.log
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
echo "exit $?"; cat $OUT/bench_${TAG}_n$N.json | cut -c1-1500; tail -3 $OUT/bench_${TAG}_n$N.err
