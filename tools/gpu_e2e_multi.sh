#!/bin/bash
# e2e at N GPUs: copy rates of the box with all GPUs copying at once, then the bench's e2e with and without staging.
N=${1:-4}; TAG=${2:-r04b}; OUT=gpurun_out; mkdir -p $OUT
echo "== pcie, 1 GPU then $N"; python tools/pcie_probe.py 1 2>&1 | tee $OUT/pcie_${TAG}_n1.txt; python tools/pcie_probe.py $N 2>&1 | tee $OUT/pcie_${TAG}_n$N.txt
nproc; free -g | head -2
for mode in "" "--no-stage"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-qv --no-syn16k --no-ref-cuda $mode > $OUT/bench_e2e_${TAG}_n$N$mode.json 2> $OUT/bench_e2e_${TAG}_n$N$mode.err
  echo "mode=[$mode] exit $?"
  python - <<P
import json
d=json.load(open("$OUT/bench_e2e_${TAG}_n$N$mode.json"))
print("  value %.4g  ms %.3f  frac %.3f  e2e %.4g (%s)"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["e2e"].get("map_upload")))
P
done
