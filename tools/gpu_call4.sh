#!/bin/bash
TAG=${1:-r03d}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -30 $OUT/pytest_gpu_$TAG.log
( time timeout 600 python bench.py --steps 100 --e2e-steps 100 --no-configs > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2>&1 | tail -3
python -c "
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().split('\n')[-1])
print({k: j[k] for k in ('value','ms_per_step','reset_ms')}, j['roofline']['kernel_ms'])"
