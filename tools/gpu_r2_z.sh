#!/bin/bash
# round 2, GPU call Z: the other BASELINE configurations with the final kernels, CTA-pair re-test, time line of the three-issuer attention kernel
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
bash tools/gpu_r2_configs.sh
{ for pr in 0 1; do B200_IMAGEN_GEMM_PAIR=$pr timeout 300 python tools/gemm_bench.py child 2>&1; done; } | tee $OUT/z_gemm_pair.txt
B200_IMAGEN_FA_VARIANT=115 timeout 300 python tools/attn_trace.py $OUT/attn_trace_split.bin 2>&1 | tee $OUT/z_attn_trace_split.txt | tail -8
