#!/bin/bash
# scratch GPU call: instruction-cache metrics of k_step (SM ICC, GPC L1.5 = GCC, L2)
TAG=${1:-r04c}
OUT=gpurun_out; mkdir -p $OUT
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum,sm__icc_requests_lookup_miss_tag_hit.sum,sm__icc_requests_lookup_miss_tag_miss.sum,sm__icc_requests_lookup_miss_tag_unavailable.sum
M=$M,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction_lookup_hit.sum,gcc__cache_requests_type_instruction_lookup_miss.sum,gcc__cache_requests_type_instruction_lookup_miss_tag_hit.sum,gcc__cache_requests_type_instruction_lookup_miss_tag_miss.sum
M=$M,gcc__gcc2xbar_requests_type_instruction.sum,gcc__xbar2gcc_sectors.sum,gcc__cycles_elapsed.sum,gcc__cycles_elapsed.avg,gcc__cycles_active.avg,sm__cycles_elapsed.avg,sm__cycles_elapsed.sum
M=$M,gcc__cache_requests_type_constant.sum,gcc__cache_requests_type_constant_lookup_miss.sum
M=$M,smsp__warps_issue_stalled_no_instruction.sum,smsp__warps_active.sum,smsp__inst_executed.sum,smsp__warps_issue_stalled_barrier.sum,smsp__warps_issue_stalled_long_scoreboard.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:k_step -s 30 -c 2 --csv --log-file $OUT/icache_$TAG.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu --no-configs --e2e-steps 3 > $OUT/icache_$TAG.log 2>&1; echo "ncu rc=$?"
cat $OUT/icache_$TAG.csv | tail -80
