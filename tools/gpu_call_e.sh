#!/bin/bash
TAG=${1:-r04o}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python tools/e2e_diag.py > $OUT/e2e_diag_$TAG.txt 2>&1; tail -5 $OUT/e2e_diag_$TAG.txt
CC4_HOST_STREAMED=0 timeout 600 python tools/e2e_diag.py > $OUT/e2e_diag_old_$TAG.txt 2>&1; tail -2 $OUT/e2e_diag_old_$TAG.txt
CC4_HOST_STREAM=1 timeout 600 python tools/e2e_diag.py > $OUT/e2e_diag_nt_$TAG.txt 2>&1; tail -1 $OUT/e2e_diag_nt_$TAG.txt
CC4_HOST_THREADS=12 timeout 600 python tools/e2e_diag.py > $OUT/e2e_diag_t12_$TAG.txt 2>&1; tail -1 $OUT/e2e_diag_t12_$TAG.txt
nproc; lscpu | grep -i -E "model name|L3|L2|Thread|Core|Socket" 
