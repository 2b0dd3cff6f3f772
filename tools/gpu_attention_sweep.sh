#!/bin/bash
# attention variant sweep: attention barrier-chain experiments (one arrival per warp, P handed over in 32-key chunks, score preload)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
SWEEP_VARIANTS="${SWEEP_VARIANTS:-41:100,54:100,60:100,61:100,58:100,59:100,62:100,63:100,47:100,56:100,53:100,55:100,52:100,57:100,60:0,60:400}" timeout 1200 python tools/sweep_attention.py 2>&1 | tee $OUT/l_attn_sweep.txt
