#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/k_all.log 2>&1; echo "k_all $? $(tail -n1 $OUT/k_all.log)"; grep -E "^E |^FAILED" $OUT/k_all.log | head -20
timeout 1500 $PYT tests/test_gpu_unet.py -m gpu > $OUT/u_tc.log 2>&1; echo "u_tc $? $(tail -n1 $OUT/u_tc.log)"; grep -E "^E |^FAILED" $OUT/u_tc.log | head -30
timeout 1200 python bench.py > $OUT/bench_full.log 2>&1; echo "bench_full $?"; grep '^{' $OUT/bench_full.log | cut -c1-260
timeout 600 python tools/op_times.py > $OUT/op_times.txt 2>&1; head -3 $OUT/op_times.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py > $OUT/prof.log 2>&1; echo "ncu launches $?"
python tools/gemm_table.py $OUT/launches.csv $OUT/plan_gemms.json > $OUT/gemm_table.txt 2>&1; tail -n 7 $OUT/gemm_table.txt
python tools/launch_summary.py $OUT/launches.csv 0 100 > $OUT/launch_summary.txt 2>&1; head -8 $OUT/launch_summary.txt
