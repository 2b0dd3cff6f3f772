#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu -x > $OUT/k_all.log 2>&1; echo "k_all $? $(tail -n1 $OUT/k_all.log)"; grep -E "^E " $OUT/k_all.log | head
timeout 500 python tools/row_bench.py > $OUT/row_bench.log 2>&1; cat $OUT/row_bench.log
timeout 1500 $PYT tests/test_gpu_unet.py -m gpu -x > $OUT/u_tc.log 2>&1; echo "u_tc $? $(tail -n1 $OUT/u_tc.log)"; grep -E "^E " $OUT/u_tc.log | head
timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_100.log 2>&1; echo "bench100 $?"; grep '^{' $OUT/bench_100.log | cut -c1-200
B200_IMAGEN_GEMM_PAIR=0 timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_100_nopair.log 2>&1; echo "bench100 nopair $?"; grep '^{' $OUT/bench_100_nopair.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py > $OUT/prof.log 2>&1; echo "ncu $?"
python tools/gemm_table.py $OUT/launches.csv $OUT/plan_gemms.json > $OUT/gemm_table.txt 2>&1; tail -n 12 $OUT/gemm_table.txt
B200_IMAGEN_GEMM_PAIR=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_nopair.csv python tools/profile_step.py > $OUT/prof2.log 2>&1; echo "ncu $?"
python tools/gemm_table.py $OUT/launches_nopair.csv $OUT/plan_gemms.json > $OUT/gemm_table_nopair.txt 2>&1; tail -n 12 $OUT/gemm_table_nopair.txt
