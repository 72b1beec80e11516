#!/bin/bash
TAG=${1:-r03e}
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -30 $OUT/pytest_gpu_$TAG.log
timeout 900 python tools/build_variants.py --bench > $OUT/variants_$TAG.txt 2>&1
cat $OUT/variants_$TAG.txt
timeout 900 python tools/build_variants.py --bench > $OUT/variants2_$TAG.txt 2>&1
cat $OUT/variants2_$TAG.txt
( time timeout 600 python bench.py --steps 100 --e2e-steps 100 --no-configs > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2>&1 | tail -3
python -c "
import json
j=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().split('\n')[-1])
print({k: j[k] for k in ('value','ms_per_step','reset_ms','reset_ms_samples','value_incl_reset')}, j['roofline']['kernel_ms'])"
CC4_RESET_SERIAL=1 timeout 600 python bench.py --steps 20 --e2e-steps 5 --no-configs --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('serial reset', j['reset_ms_samples'])"
