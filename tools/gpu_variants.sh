#!/bin/bash
# GPU: parity tests on the default build, then time every tuning variant.  bash tools/gpu_variants.sh tag
TAG=${1:-var}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 1500 python tools/build_variants.py --bench > $OUT/variants_$TAG.txt 2>&1
cat $OUT/variants_$TAG.txt
