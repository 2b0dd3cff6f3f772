#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py > $OUT/prof.log 2>&1; echo "ncu $?"
python tools/gemm_table.py $OUT/launches.csv $OUT/plan_gemms.json > $OUT/gemm_table.txt 2>&1; tail -n 8 $OUT/gemm_table.txt
python tools/launch_summary.py $OUT/launches.csv 0 100 > $OUT/launch_summary.txt; head -30 $OUT/launch_summary.txt
