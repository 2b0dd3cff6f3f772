#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for d in 0 1 2 3; do echo "rms debug $d"; B200_IMAGEN_RMS_DEBUG=$d B200_IMAGEN_ROW_VPT=2 timeout 300 python tools/row_bench.py child 2>&1 | head -2; done
