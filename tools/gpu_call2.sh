#!/bin/bash
TAG=${1:-r03b}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
( time timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -c 6000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
( time timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench20_$TAG.json 2>> $OUT/bench_$TAG.err ) 2>&1 | tail -3
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err ) 2>&1 | tail -3; cat $OUT/bench_ref_$TAG.json
for e in 4736 9472; do timeout 200 python tools/phase_wall.py base --envs $e --warm 150 --ticks 200; done > $OUT/phase_wall_epb_$TAG.txt 2>&1
cat $OUT/phase_wall_epb_$TAG.txt
