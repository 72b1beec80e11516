#!/bin/bash
TAG=${1:-r03c}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu_$TAG.log
timeout 900 python tools/build_variants.py --bench > $OUT/variants_$TAG.txt 2>&1
cat $OUT/variants_$TAG.txt
( time timeout 600 python bench.py --steps 100 --e2e-steps 100 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2>&1 | tail -3
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_%s.json" % "r03c").read().strip().split("\n")[-1])
print({k: j[k] for k in ("value","ms_per_step","reset_ms","reset_ms_samples")})
print(json.dumps(j["e2e"])[:600]); print(json.dumps(j["e2e_bits"])[:200]); print(json.dumps(j["configs"]["cfg5"])[:400])
PY
tail -3 $OUT/bench_$TAG.err
