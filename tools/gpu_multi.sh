#!/bin/bash
# 8-GPU call: the bench line at N = 8 (weak headline + configs.cfg4 strong scaling + cfg5), and the host-buffer path with / without pinned pool threads
TAG=${1:-r03f}
OUT=gpurun_out; mkdir -p $OUT
N=${2:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --e2e-steps 60 > $OUT/bench_${N}gpu_$TAG.json 2> $OUT/bench_${N}gpu_$TAG.err ) 2>&1 | tail -3
tail -c 4000 $OUT/bench_${N}gpu_$TAG.json; tail -5 $OUT/bench_${N}gpu_$TAG.err
( time CC4_HOST_PIN=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --e2e-steps 60 --no-configs > $OUT/bench_${N}gpu_nopin_$TAG.json 2>> $OUT/bench_${N}gpu_$TAG.err ) 2>&1 | tail -3
python - <<PY
import json
for f in ("bench_${N}gpu_$TAG", "bench_${N}gpu_nopin_$TAG"):
    try:
        j = json.loads(open("gpurun_out/%s.json" % f).read().strip().split("\n")[-1])
        print(f, "value", j["value"], "e2e", j["e2e"]["value"], "host_write_gbs", j["e2e"]["host_write_gbs"], "frac", j["e2e"]["host_write_frac"], "bits", j["e2e_bits"]["value"])
    except Exception as e:
        print(f, "ERR", e)
PY
