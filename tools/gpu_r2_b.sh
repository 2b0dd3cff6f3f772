#!/bin/bash
# round 2, GPU call B: register-resident fused-norm epilogue (kernel tests + A/B bench), attention MUFU-token variants
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 600 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/b_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/b_kernels.log)"; grep -E "^E  |^FAILED" $OUT/b_kernels.log | head -40
SWEEP_VARIANTS=12,20,30,31,32,33,34,36,37 timeout 900 python tools/sweep_attention.py > $OUT/b_sweep.txt 2>&1; cat $OUT/b_sweep.txt
timeout 1500 $PYT tests/test_gpu_unet.py tests/test_gpu_baseline_shapes.py -m gpu -x > $OUT/b_unet.log 2>&1; echo "unet $? $(tail -n1 $OUT/b_unet.log)"; grep -E "^E  |^FAILED" $OUT/b_unet.log | head -40
B200_IMAGEN_FUSE_NORM=0 timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/b_bench_100_nofuse.log 2>&1; echo "bench100 nofuse $?"; grep '^{' $OUT/b_bench_100_nofuse.log | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/b_bench_100.log 2>&1; echo "bench100 $?"; grep '^{' $OUT/b_bench_100.log | cut -c1-200; tail -n 3 $OUT/b_bench_100.log | grep -v '^{' | cut -c1-300
for v in 30 33; do B200_IMAGEN_FA_VARIANT=$v timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/b_bench_100_fa$v.log 2>&1; echo "bench100 fa$v $?"; grep '^{' $OUT/b_bench_100_fa$v.log | cut -c1-200; done
timeout 600 python tools/op_times.py > $OUT/b_op_times.txt 2>&1; head -48 $OUT/b_op_times.txt
