#!/bin/bash
# round 2, GPU call N: attention per-CTA fixed cost vs per-tile cost, staged im2col test + timing, kernel tests with the new attention default
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
for v in 65 69; do echo "== variant $v"; B200_IMAGEN_FA_VARIANT=$v timeout 300 python tools/attn_keys_scan.py 2>&1 | tail -9; done | tee $OUT/n_attn_keys_scan.txt
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/n_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/n_kernels.log)"; grep -E "^E  |^FAILED" $OUT/n_kernels.log | head -30
for x in 0 1; do B200_IMAGEN_IM2COL_STAGED=$x timeout 600 python tools/op_times.py 2>&1 | grep -E "total|im2col|b200_attention" | sed "s/^/im2col_staged=$x /"; done | tee $OUT/n_im2col_ab.txt
timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/n_bench.log 2>&1; echo "bench $? $(grep '^{' $OUT/n_bench.log | cut -c1-140)"
