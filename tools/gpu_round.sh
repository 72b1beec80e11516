#!/bin/bash
# One gpurun call: GPU parity tests, smoke, both bench arms, the ncu launch list of a short bench run, one full capture of k_step
# (and of k_reset_fast).  Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh [tag]
# Variants: python tools/build_variants.py name="-D..." (here), then `python tools/build_variants.py --bench` on the box;
# phase walls: python tools/phase_wall.py --build [name -D...] (here), then `python tools/phase_wall.py name` on the box.
TAG=${1:-r03}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
tail -c 3000 $OUT/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "bench ref rc=$?"
# launch list of a short bench run: per-launch device time, cold-cache and serialised
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu --no-configs --e2e-steps 3 > $OUT/bench_under_ncu_$TAG.log 2>&1; echo "ncu launches rc=$?"
# full captures
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -o $OUT/prof_kstep_$TAG -f \
    python bench.py --steps 40 --warmup 3 --no-cpu --no-configs --e2e-steps 3 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:k_reset_fast -c 1 -o $OUT/prof_kreset_$TAG -f \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --e2e-steps 3 > $OUT/ncu_reset_$TAG.log 2>&1; echo "ncu reset rc=$?"
# digests kept under profiles/ (the .ncu-rep files stay in gpurun_out/)
python tools/ncu_summary.py $OUT/prof_kstep_$TAG.ncu-rep > $OUT/kstep_ncu_full_summary_$TAG.txt 2>&1
python tools/ncu_by_function.py $OUT/prof_kstep_$TAG.ncu-rep --lines 40 > $OUT/kstep_by_function_$TAG.txt 2>&1
python tools/ncu_summary.py $OUT/prof_kreset_$TAG.ncu-rep > $OUT/kreset_ncu_full_summary_$TAG.txt 2>&1
python tools/sass_evidence.py > $OUT/sass_$TAG.txt 2>&1
head -25 $OUT/kstep_ncu_full_summary_$TAG.txt
ls -la $OUT | tail -20
