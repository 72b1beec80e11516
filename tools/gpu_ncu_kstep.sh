TAG=$1; OUT=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o $OUT/prof_kstep_$TAG -f python bench.py --steps 40 --warmup 3 --no-cpu --no-configs --e2e-steps 3 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py $OUT/prof_kstep_$TAG.ncu-rep > $OUT/kstep_ncu_full_summary_$TAG.txt; python tools/ncu_by_function.py $OUT/prof_kstep_$TAG.ncu-rep --lines 120 > $OUT/kstep_by_function_$TAG.txt 2>&1; head -22 $OUT/kstep_ncu_full_summary_$TAG.txt
