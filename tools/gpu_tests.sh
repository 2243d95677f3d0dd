#!/bin/bash
# All GPU-tier tests (+ the CPU tier, which is cheap) on the GPU box.
TAG=${1:-run}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -q -m gpu --timeout=900 > $OUT/pytest_$TAG.log 2>&1; echo "pytest gpu exit $?"; tail -12 $OUT/pytest_$TAG.log
