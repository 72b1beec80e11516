#!/bin/bash
# compute-sanitizer over every output path and kernel instantiation (tools/sanitizer_smoke.py); run on the GPU box
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 compute-sanitizer --tool memcheck --log-file $OUT/sanitizer_smoke_r04.log python tools/sanitizer_smoke.py > $OUT/sanitizer_smoke_r04.out 2>&1; echo "memcheck rc=$?"; tail -2 $OUT/sanitizer_smoke_r04.log; tail -1 $OUT/sanitizer_smoke_r04.out
timeout 1200 compute-sanitizer --tool racecheck --log-file $OUT/racecheck_smoke_r04.log python tools/sanitizer_smoke.py > $OUT/racecheck_smoke_r04.out 2>&1; echo "racecheck rc=$?"; grep "Error: Race" $OUT/racecheck_smoke_r04.log | sort | uniq -c | head; tail -1 $OUT/racecheck_smoke_r04.log
