#!/bin/bash
TAG=${1:-r03h}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu_$TAG.log
for i in 1 2; do timeout 900 python tools/build_variants.py --bench > $OUT/variants${i}_$TAG.txt 2>&1; cat $OUT/variants${i}_$TAG.txt | cut -c1-150; done
