#!/bin/bash
# N-GPU call: the bench line at N GPUs (weak headline + configs.cfg4 strong scaling + cfg5 + both host-buffer forms) and the reference arm.  bash tools/gpu_multi.sh tag [N]
TAG=${1:-r04x}
OUT=gpurun_out; mkdir -p $OUT
N=${2:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --e2e-steps 60 > $OUT/bench_${N}gpu_$TAG.json 2> $OUT/bench_${N}gpu_$TAG.err ) 2>&1 | tail -3
tail -c 3000 $OUT/bench_${N}gpu_$TAG.json; tail -5 $OUT/bench_${N}gpu_$TAG.err
python - <<PY
import json
for f in ("bench_${N}gpu_$TAG",):
    try:
        j = json.loads(open("gpurun_out/%s.json" % f).read().strip().split("\n")[-1])
        print(f, "value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], "host_write_gbs", j["e2e"].get("host_write_gbs"), "frac", j["e2e"].get("host_write_frac"), "bits", j["e2e_bits"]["value"], "cfg4", j["configs"]["cfg4"]["value"], "cfg5 wall", j["configs"]["cfg5"]["wall_s"])
    except Exception as e:
        print(f, "ERR", e)
PY
