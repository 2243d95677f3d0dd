#!/bin/bash
# Short multi-GPU check (gpurun --gpus N): distributed parity tests + the bench under torchrun.
N=${1:-2}; TAG=${2:-r03}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest multi-GPU"; timeout 1200 python -m pytest tests/test_distributed_gpu.py tests/test_host_mirror.py -q -m gpu --timeout=900 > $OUT/pytest_multi_${TAG}_n$N.log 2>&1; echo "exit $?"; tail -3 $OUT/pytest_multi_${TAG}_n$N.log
echo "== bench N=$N"; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err; echo "exit $?"; tail -3 $OUT/bench_${TAG}_n$N.err
python - <<'P'
import json,sys,glob
for f in sorted(glob.glob('gpurun_out/bench_*_n*.json')):
    try: d=json.load(open(f))
    except Exception as e: continue
    if 'qv_tree' not in d: continue
    print(f, 'value %.4g ms %.3f frac %.3f e2e %.4g | syn16k %.4g %s | qv %.4g %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['syn16k']['cell_updates_per_sec'],d['syn16k']['solution_checksum'],d['qv_tree']['plans_per_sec'],d['qv_tree'].get('seconds_of_3_batches')))
P
