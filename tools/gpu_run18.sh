#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
timeout 900 python tools/gemm_bench.py > $OUT/gemm_bench.log 2>&1; cat $OUT/gemm_bench.log
B200_IMAGEN_ROW_VPT=2 timeout 300 python tools/row_bench.py child > $OUT/row_bench2.log 2>&1; cat $OUT/row_bench2.log
