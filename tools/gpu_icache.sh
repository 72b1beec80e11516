#!/bin/bash
# GPU: instruction-cache (SM ICC, GPC L1.5) and stall counters of k_step for the named variants.  bash tools/gpu_icache.sh tag variant ...
TAG=${1:-r04j}; shift
OUT=gpurun_out; mkdir -p $OUT
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss_tag_hit.sum,sm__icc_requests_lookup_miss_tag_miss.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction_lookup_miss.sum
M=$M,smsp__warps_issue_stalled_no_instruction.sum,smsp__warps_active.sum,smsp__inst_executed.sum,smsp__warps_issue_stalled_barrier.sum,smsp__warps_issue_stalled_long_scoreboard.sum,smsp__warps_issue_stalled_short_scoreboard.sum,smsp__warps_issue_stalled_wait.sum,smsp__warps_issue_stalled_branch_resolving.sum,gpu__time_duration.sum
for v in "$@"; do
CC4_B200_LIB=$PWD/cage_challenge_4_b200/variants/libcc4_b200_$v.so timeout 600 ncu --metrics $M --clock-control none -k regex:k_step -s 30 -c 1 --csv --log-file $OUT/icache_${v}_$TAG.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu --no-configs --e2e-steps 3 > $OUT/icache_${v}_$TAG.log 2>&1; echo "ncu $v rc=$?"
python - <<PY
import csv
rows=list(csv.DictReader(l for l in open("$OUT/icache_${v}_$TAG.csv") if l.startswith('"')))
print("$v", {r["Metric Name"].replace("smsp__warps_issue_stalled_","st_").replace(".sum",""): r["Metric Value"] for r in rows})
PY
done
