#!/bin/bash
# GPU: parity tests on the default build, time every tuning variant, one ncu full capture of k_step.  bash tools/gpu_variants.sh tag
TAG=${1:-var}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 1500 python tools/build_variants.py --bench > $OUT/variants_$TAG.txt 2>&1
cat $OUT/variants_$TAG.txt
timeout 600 python bench.py --no-cpu --steps 200 --e2e-steps 30 > $OUT/bench_$TAG.json 2>$OUT/bench_$TAG.err; cat $OUT/bench_$TAG.json
if [ -z "$NO_NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 150 -c 1 -o $OUT/prof_kstep_$TAG -f \
    python bench.py --steps 160 --warmup 3 --no-cpu --e2e-steps 3 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
