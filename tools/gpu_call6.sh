#!/bin/bash
TAG=${1:-r03g}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_wrappers_gpu.py -m gpu -q -x -k "compat or evaluation" > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu_$TAG.log
for i in 1 2; do timeout 900 python tools/build_variants.py --bench > $OUT/variants${i}_$TAG.txt 2>&1; cat $OUT/variants${i}_$TAG.txt | cut -c1-150; done
