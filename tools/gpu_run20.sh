#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
for c in plain; do
GEMM_BENCH_ONLY=$c timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_tc -s 4 -c 1 -o $OUT/gemm_$c -f python tools/gemm_bench.py child > $OUT/ncu_$c.log 2>&1; echo "ncu $c $?"
done
ls -la $OUT/*.ncu-rep
