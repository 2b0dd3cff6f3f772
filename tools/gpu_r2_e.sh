#!/bin/bash
# round 2, GPU call E: transposed GEMM for <= 128 output channels (tests + A/B), attention poly / barrier-wait sweep
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/e_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/e_kernels.log)"; grep -E "^E  |^FAILED" $OUT/e_kernels.log | head -40
for t in 0 1; do GEMM_BENCH_ONLY="conv3x3 64x64" B200_IMAGEN_GEMM_T=$t timeout 300 python tools/gemm_bench.py child 2>&1 | sed "s/^/T=$t /" | tail -n 2; done | tee $OUT/e_gemm_t_ab.txt
for t in 0 1; do GEMM_BENCH_ONLY="to_out" B200_IMAGEN_GEMM_T=$t timeout 300 python tools/gemm_bench.py child 2>&1 | sed "s/^/T=$t /" | tail -n 2; done | tee -a $OUT/e_gemm_t_ab.txt
SWEEP_VARIANTS=40,40:200,40:1000,47,47:200,47:1000,39,48,49,44,45,46,46:1000 timeout 1200 python tools/sweep_attention.py > $OUT/e_sweep.txt 2>&1; cat $OUT/e_sweep.txt
timeout 1500 $PYT tests/test_gpu_unet.py tests/test_gpu_baseline_shapes.py -m gpu -x > $OUT/e_unet.log 2>&1; echo "unet $? $(tail -n1 $OUT/e_unet.log)"; grep -E "^E  |^FAILED" $OUT/e_unet.log | head -40
B200_IMAGEN_GEMM_T=0 timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/e_bench_100_noT.log 2>&1; echo "bench100 noT $?"; grep '^{' $OUT/e_bench_100_noT.log | cut -c1-200; tail -n 3 $OUT/e_bench_100_noT.log | grep -v '^{' | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/e_bench_100.log 2>&1; echo "bench100 $?"; grep '^{' $OUT/e_bench_100.log | cut -c1-200; tail -n 3 $OUT/e_bench_100.log | grep -v '^{' | cut -c1-300
B200_IMAGEN_FA_VARIANT=47 timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/e_bench_100_fa47.log 2>&1; echo "bench100 fa47 $?"; grep '^{' $OUT/e_bench_100_fa47.log | cut -c1-200
timeout 600 python tools/op_times.py > $OUT/e_op_times.txt 2>&1; head -30 $OUT/e_op_times.txt
