#!/bin/bash
TAG=${1:-r03j}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu_$TAG.log
for i in 1 2; do timeout 900 python tools/build_variants.py --bench > $OUT/variants${i}_$TAG.txt 2>&1; cat $OUT/variants${i}_$TAG.txt | cut -c1-150; done
( time timeout 600 python bench.py --steps 60 --e2e-steps 60 --no-configs --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2>&1 | tail -3
python -c "
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().split('\n')[-1])
print({k: j[k] for k in ('value','ms_per_step','reset_ms','reset_ms_samples','value_incl_reset')}, j['roofline']['kernel_ms'])"
