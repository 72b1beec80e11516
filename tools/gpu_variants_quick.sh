#!/bin/bash
# GPU: time every tuning variant under cage_challenge_4_b200/variants twice (drift check) and parity-test the named ones.  bash tools/gpu_variants_quick.sh tag [variant ...]
TAG=${1:-r04a}; shift
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 1500 python tools/build_variants.py --bench > $OUT/variants_$TAG.txt 2>&1
timeout 1500 python tools/build_variants.py --bench >> $OUT/variants_$TAG.txt 2>&1
cat $OUT/variants_$TAG.txt
for v in "$@"; do
CC4_B200_LIB=$PWD/cage_challenge_4_b200/variants/libcc4_b200_$v.so timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q > $OUT/pytest_${v}_$TAG.log 2>&1; echo "pytest $v rc=$?"; tail -3 $OUT/pytest_${v}_$TAG.log
done
