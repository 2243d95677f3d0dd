#!/bin/bash
N=${1:-2}; TAG=${2:-r04i}; OUT=gpurun_out; mkdir -p $OUT
export PP2D_E2E_TRACE=1
for pub in ${3:-1}; do
  export PP2D_P2P_PUBLISHER=$pub
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-qv --no-ref-cuda > $OUT/bench_pub${pub}_${TAG}_n$N.json 2> $OUT/bench_pub${pub}_${TAG}_n$N.err
  echo "publisher=$pub exit $?"; grep -h "rank 0 e2e ms" $OUT/bench_pub${pub}_${TAG}_n$N.err | cut -c1-330
  python - <<P
import json
d=json.load(open("$OUT/bench_pub${pub}_${TAG}_n$N.json"))
print("  value %.4g  ms %.3f  frac %.3f launch_us %.2f e2e %.4g policy-only %.4g | syn16k %.4g %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["roofline"]["launch_ms"]*1e3,d["e2e"]["value"],d["e2e"]["policy_only"]["value"],d["syn16k"]["cell_updates_per_sec"],d["syn16k"]["solution_checksum"]))
P
done
