#!/bin/bash
# GPU box: records of the reference's own QV-tree host code + the pin tests.
OUT=gpurun_out; mkdir -p $OUT/golden_ref
python tests/golden/make_golden.py tree $OUT/golden_ref > $OUT/golden_tree.log 2>&1; echo "golden tree exit $?"; grep -v "^$" $OUT/golden_tree.log | tail -5
cp $OUT/golden_ref/tree_*.npz tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests/test_tree_pin_cpu.py tests/test_tree_pin_gpu.py -q --timeout=600 2>&1 | tail -15
