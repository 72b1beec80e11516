#!/bin/bash
# scratch GPU call: full GPU test suite + smoke on the default build, bench, ncu full capture of k_step with source counters
TAG=${1:-r04n}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_$TAG.log
timeout 900 python bench.py --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; head -c 1200 $OUT/bench_$TAG.json; echo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o $OUT/prof_kstep_$TAG -f \
    python bench.py --steps 40 --warmup 3 --no-cpu --no-configs --e2e-steps 3 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_by_function.py $OUT/prof_kstep_$TAG.ncu-rep --lines 40 > $OUT/kstep_by_function_$TAG.txt 2>&1; head -60 $OUT/kstep_by_function_$TAG.txt
ls -la $OUT/prof_kstep_$TAG.ncu-rep
