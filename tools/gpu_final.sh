#!/bin/bash
# Round-end evidence run: full bench (both arms), launch list, warm per-call times, ncu --set full captures of the two dominant kernels.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
timeout 1200 python bench.py > $OUT/bench_full.log 2>&1; echo "bench_full $?"; grep '^{' $OUT/bench_full.log | cut -c1-300
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.log 2>&1; echo "bench_ref $?"; grep '^{' $OUT/bench_ref.log | cut -c1-300
timeout 900 python bench.py --impl reference --eager-gpu --steps 2 --warmup 1 > $OUT/bench_ref_gpu.log 2>&1; echo "bench_ref_gpu $?"; grep '^{' $OUT/bench_ref_gpu.log | cut -c1-300
timeout 600 python tools/op_times.py > $OUT/op_times.txt 2>&1; head -3 $OUT/op_times.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_step.py > $OUT/prof.log 2>&1; echo "ncu launches $?"
python tools/gemm_table.py $OUT/launches.csv $OUT/plan_gemms.json > $OUT/gemm_table.txt 2>&1; tail -n 8 $OUT/gemm_table.txt
python tools/launch_summary.py $OUT/launches.csv 0 100 > $OUT/launch_summary.txt 2>&1; head -12 $OUT/launch_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn_pp -s 2 -c 1 -o $OUT/prof_attn -f python tools/profile_step.py 1 16 > $OUT/prof_attn.log 2>&1; echo "ncu attn $?"
GEMM_BENCH_ONLY="conv3x3 64x64" timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_tc -s 4 -c 1 -o $OUT/prof_gemm -f python tools/gemm_bench.py child > $OUT/prof_gemm.log 2>&1; echo "ncu gemm $?"
ls -la $OUT/*.ncu-rep
