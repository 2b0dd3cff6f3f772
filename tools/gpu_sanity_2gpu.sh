#!/bin/bash
# 2-GPU sanity: sanity of the final (cleaned) build on 2 GPUs: kernel tests, smoke, 1- and 2-GPU bench lines with per-rank own sampling times
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 600 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/f_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/f_kernels.log)"; grep -E "^E  |^FAILED" $OUT/f_kernels.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/f_smoke.log 2>&1; echo "smoke $? $(tail -n2 $OUT/f_smoke.log | tr '\n' ' ' | cut -c1-300)"
SCALE_NS="1 2" SCALE_TIMESTEPS=100 bash tools/gpu_scale.sh 2>&1 | grep -v "NCCL INFO\|^NIC\|^GPU[0-9]\|Legend\|^  [A-Z]* *=" | tail -12
