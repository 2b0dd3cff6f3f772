#!/bin/bash
# round 2, GPU call R: persistent self-attention kernel -- sweep, parity tests under the variant, per-CTA fixed cost, step A/B
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
SWEEP_VARIANTS="65:100,90:100,91:100,92:100,93:100,94:100,90:0,90:400" timeout 900 python tools/sweep_attention.py 2>&1 | tee $OUT/r_attn_sweep.txt
B200_IMAGEN_FA_VARIANT=90 timeout 600 $PYT tests/test_gpu_kernels.py -m gpu -k "attention" > $OUT/r_kernels_attn.log 2>&1; echo "attention tests (variant 90) $? $(tail -n1 $OUT/r_kernels_attn.log)"; grep -E "^E  |^FAILED" $OUT/r_kernels_attn.log | head -20
B200_IMAGEN_FA_VARIANT=90 timeout 300 python tools/attn_keys_scan.py 2>&1 | tail -9 | tee $OUT/r_attn_keys_scan.txt
for v in 65 90 65 90; do
  B200_IMAGEN_FA_VARIANT=$v timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/r_bench_v$v.log 2>&1
  echo "bench variant=$v $? $(grep '^{' $OUT/r_bench_v$v.log | cut -c1-140)"
done | tee $OUT/r_persistent_ab.txt
