#!/bin/bash
# e2e (host-buffer path) under different host-side settings; run on the GPU box.  bash tools/e2e_sweep.sh
run() { env "$@" timeout 300 python bench.py --steps 60 --no-cpu --e2e-steps 200 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(j['value']/1e6,1), round(j['e2e']['value']/1e6,1))"; }
run CC4_HOST_THREADS=16
run CC4_HOST_THREADS=8
run CC4_HOST_THREADS=16 CC4_HOST_NO_AVX2=1
run CC4_HOST_THREADS=16 CC4_HOST_STREAM=1
run CC4_HOST_THREADS=16 CC4_HOST_CHUNK_EPISODES=4736
run CC4_HOST_THREADS=16 CC4_HOST_CHUNK_EPISODES=18944
run CC4_HOST_THREADS=16 CC4_HOST_CHUNK_EPISODES=6336
run CC4_HOST_EXPAND=0
