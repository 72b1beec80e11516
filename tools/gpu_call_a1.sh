#!/bin/bash
# scratch GPU call: time the tuning variants once
TAG=${1:-r04a}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python tools/build_variants.py --bench > $OUT/variants_$TAG.txt 2>&1
cat $OUT/variants_$TAG.txt
