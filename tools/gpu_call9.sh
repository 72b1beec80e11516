#!/bin/bash
TAG=${1:-r03i}
OUT=gpurun_out; mkdir -p $OUT
for i in 1 2; do timeout 900 python tools/build_variants.py --bench > $OUT/variants${i}_$TAG.txt 2>&1; cat $OUT/variants${i}_$TAG.txt | cut -c1-150; done
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -q -x > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
# launch list of a short bench run (per-launch device time, cold-cache and serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu --no-configs --e2e-steps 3 > $OUT/bench_under_ncu_$TAG.log 2>&1; echo "ncu launches rc=$?"
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -o $OUT/prof_kstep_$TAG -f \
    python bench.py --steps 40 --warmup 3 --no-cpu --no-configs --e2e-steps 3 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:k_reset_fast -c 1 -o $OUT/prof_kreset_$TAG -f \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --e2e-steps 3 > $OUT/ncu_reset_$TAG.log 2>&1; echo "ncu reset rc=$?"
ls -la $OUT | tail -8
