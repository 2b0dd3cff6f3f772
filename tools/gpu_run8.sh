#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
timeout 1200 python bench.py > $OUT/bench_full.log 2>&1; echo "bench_full $?"; tail -n 1 $OUT/bench_full.log | cut -c1-250
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.log 2>&1; echo "bench_ref $?"; tail -n 1 $OUT/bench_ref.log | cut -c1-250
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn_pp -s 1 -c 1 -o $OUT/prof_attn_pp -f python tools/profile_step.py 1 16 > $OUT/prof_attn.log 2>&1; echo "ncu attn $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc -s 6 -c 6 -o $OUT/prof_gemm3 -f python tools/profile_step.py 1 16 > $OUT/prof_gemm.log 2>&1; echo "ncu gemm $?"
