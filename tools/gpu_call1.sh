#!/bin/bash
# round-2 call 1: GPU parity tests (default build), flow-variant parity, variant timings, phase walls
TAG=${1:-r03a}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
for v in flow31 flow5; do
CC4_B200_LIB=$PWD/cage_challenge_4_b200/variants/libcc4_b200_$v.so timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "philox or seam or tape or host" > $OUT/pytest_${v}_$TAG.log 2>&1; echo "pytest $v rc=$?"; tail -3 $OUT/pytest_${v}_$TAG.log
done
timeout 900 python tools/build_variants.py --bench > $OUT/variants_$TAG.txt 2>&1
cat $OUT/variants_$TAG.txt
timeout 300 python tools/phase_wall.py base flow31 > $OUT/phase_wall_$TAG.txt 2>&1
cat $OUT/phase_wall_$TAG.txt
