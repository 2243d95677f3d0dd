#!/bin/bash
# e2e at N GPUs: does the download overlap the solve?  A/B over CUDA_DEVICE_MAX_CONNECTIONS (hardware work queues per process).
N=${1:-2}; TAG=${2:-r04c}; OUT=gpurun_out; mkdir -p $OUT
python tools/pcie_probe.py $N 2>&1 | tee $OUT/pcie_${TAG}_n$N.txt
for conn in default 32; do
  if [ $conn = default ]; then unset CUDA_DEVICE_MAX_CONNECTIONS; else export CUDA_DEVICE_MAX_CONNECTIONS=$conn; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-qv --no-syn16k --no-ref-cuda > $OUT/bench_e2e_${TAG}_n${N}_conn$conn.json 2> $OUT/bench_e2e_${TAG}_n${N}_conn$conn.err
  echo "conn=$conn exit $?"
  python - <<P
import json
d=json.load(open("$OUT/bench_e2e_${TAG}_n${N}_conn$conn.json"))
print("  value %.4g  ms %.3f  frac %.3f  e2e %.4g (%s)"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["e2e"].get("map_upload")))
P
done
