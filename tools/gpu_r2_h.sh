#!/bin/bash
# round 2, GPU call H: GEMM regression check (plain barrier polling restored), kernel tests, bench, the other BASELINE configurations
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
for t in 0 1; do GEMM_BENCH_ONLY="conv3x3" B200_IMAGEN_GEMM_T=$t timeout 300 python tools/gemm_bench.py child 2>&1 | sed "s/^/T=$t /" | tail -n 4; done | tee $OUT/h_gemm_conv.txt
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/h_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/h_kernels.log)"; grep -E "^E  |^FAILED" $OUT/h_kernels.log | head -40
timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/h_bench_100.log 2>&1; echo "bench100 $? $(grep '^{' $OUT/h_bench_100.log | cut -c1-160)"; tail -n 3 $OUT/h_bench_100.log | grep -v '^{' | cut -c1-300
timeout 600 python tools/op_times.py > $OUT/h_op_times.txt 2>&1; head -24 $OUT/h_op_times.txt
bash tools/gpu_r2_configs.sh
