#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
./tools/micro/mufu_bench > $OUT/mufu_bench.log 2>&1; cat $OUT/mufu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn_pp -s 3 -c 1 -o $OUT/prof_attn_pp3 -f python tools/attn_only.py > $OUT/prof_attn.log 2>&1; echo "ncu attn $?"
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 1500 $PYT tests/test_gpu_unet.py -m gpu > $OUT/u_tc.log 2>&1; echo "u_tc $? $(tail -n1 $OUT/u_tc.log)"; grep -E "^E " $OUT/u_tc.log | head
