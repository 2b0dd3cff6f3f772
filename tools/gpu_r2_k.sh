#!/bin/bash
# round 2, GPU call K: persistent tcgen05 cross-attention kernel -- parity tests, A/B against the CUDA-core few-keys kernel, network tests
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 300 $PYT tests/test_gpu_kernels.py -m gpu -k "cross_attention" > $OUT/k_xattn.log 2>&1; echo "xattn tests $? $(tail -n1 $OUT/k_xattn.log)"; grep -E "^E  |^FAILED" $OUT/k_xattn.log | head -30
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/k_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/k_kernels.log)"; grep -E "^E  |^FAILED" $OUT/k_kernels.log | head -30
for x in 0 1 0 1; do
  B200_IMAGEN_XATTN_TC=$x timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/k_bench_x$x.log 2>&1
  echo "bench xattn_tc=$x $? $(grep '^{' $OUT/k_bench_x$x.log | cut -c1-140)"
done | tee $OUT/k_xattn_ab.txt
for x in 0 1; do B200_IMAGEN_XATTN_TC=$x timeout 600 python tools/op_times.py 2>&1 | grep -E "total|b200_attention" | sed "s/^/xattn_tc=$x /"; done | tee -a $OUT/k_xattn_ab.txt
timeout 1200 $PYT tests/test_gpu_unet.py tests/test_gpu_baseline_shapes.py -m gpu > $OUT/k_unet.log 2>&1; echo "unet $? $(tail -n1 $OUT/k_unet.log)"; grep -E "^E  |^FAILED" $OUT/k_unet.log | head -30
