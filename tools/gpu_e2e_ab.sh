#!/bin/bash
# e2e A/B: map upload staged under the previous solve vs inside pp2d_mdp_reset; PCIe rates of the box.
TAG=${1:-r04a}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest (staged upload, async download, reset)"
timeout 600 python -m pytest tests/test_mdp_gpu.py -q -m gpu -k "staged or asynchronous or reset_reuses or multi" --timeout=500 > $OUT/pytest_e2e_$TAG.log 2>&1; echo "exit $?"; tail -3 $OUT/pytest_e2e_$TAG.log
echo "== pcie"; python tools/pcie_probe.py 1 2>&1 | tee $OUT/pcie_$TAG.txt
for mode in "" "--no-stage"; do
  for rep in 1 2; do
    python bench.py --steps 10 --warmup 3 --no-cpu --no-qv --no-syn16k --no-ref-cuda $mode > $OUT/bench_e2e_${TAG}${mode}_$rep.json 2> $OUT/bench_e2e_${TAG}${mode}_$rep.err
    echo "mode=[$mode] rep=$rep exit $?"
    python - <<P
import json
d=json.load(open("$OUT/bench_e2e_${TAG}${mode}_$rep.json"))
print("  value %.4g  ms %.3f  frac %.3f  e2e %.4g (%s)"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["e2e"].get("map_upload")))
P
  done
done
