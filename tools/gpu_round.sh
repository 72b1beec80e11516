#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the bench command, one full capture of k_step.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
tail -c 3000 $OUT/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "bench ref rc=$?"
# launch list of the same bench command (short): per-launch device time, cold-cache and serialised
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu --e2e-steps 3 > $OUT/bench_under_ncu_$TAG.log 2>&1; echo "ncu launches rc=$?"
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -o $OUT/prof_kstep_$TAG -f \
    python bench.py --steps 40 --warmup 3 --no-cpu --e2e-steps 3 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
