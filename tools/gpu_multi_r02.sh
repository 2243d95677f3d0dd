#!/bin/bash
# Multi-GPU session (gpurun --gpus N): distributed parity tests, the bench under torchrun,
# and (N = 2) per-launch NVLink byte counters of the fused kernel's peer stores.
N=${1:-2}; TAG=${2:-r02m}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_${TAG}_n$N.txt 2>&1
echo "== pytest multi-GPU"; timeout 1200 python -m pytest tests/test_distributed_gpu.py tests/test_host_mirror.py -q -m gpu --timeout=900 > $OUT/pytest_multi_${TAG}_n$N.log 2>&1; echo "exit $?"; tail -5 $OUT/pytest_multi_${TAG}_n$N.log
if [ -z "$SKIP_REF" ]; then echo "== bench reference arm N=$N"; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus $N --steps 10 --warmup 3 > $OUT/bench_ref_${TAG}_n$N.json 2> $OUT/bench_ref_${TAG}_n$N.err; echo "exit $?"; cut -c1-300 $OUT/bench_ref_${TAG}_n$N.json; fi
echo "== bench N=$N"; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err; echo "exit $?"; cut -c1-1500 $OUT/bench_${TAG}_n$N.json; tail -3 $OUT/bench_${TAG}_n$N.err
if [ "$N" = "2" ]; then
  echo "== ncu NVLink counters, single-process 2-GPU handle"
  timeout 600 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_ltcfabric.sum \
    --clock-control none -k regex:mdp_sweep_kernel -c 24 --csv --log-file $OUT/nvlink_${TAG}.csv \
    python tools/ncu_multi_target.py 2 > $OUT/nvlink_${TAG}.log 2>&1; echo "exit $?"; tail -3 $OUT/nvlink_${TAG}.log
fi
