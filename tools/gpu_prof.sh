#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 600 $PYT tests/test_gpu_kernels.py -m gpu -k "ddpm or quantile" > $OUT/k_fix.log 2>&1; echo "k_fix $? $(tail -n1 $OUT/k_fix.log)"
timeout 900 $PYT tests/test_gpu_unet.py -m gpu -k "equals or replay or full_size" > $OUT/u_fix.log 2>&1; echo "u_fix $? $(tail -n1 $OUT/u_fix.log)"
timeout 600 python bench.py --steps 1 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_100.log 2>&1; echo "bench100 $?"; tail -n 3 $OUT/bench_100.log
timeout 1200 python bench.py > $OUT/bench_full.log 2>&1; echo "bench_full $?"; tail -n 3 $OUT/bench_full.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.log 2>&1; echo "bench_ref $?"; tail -n 2 $OUT/bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv python tools/profile_step.py 2 16 > $OUT/prof_launch.log 2>&1; echo "ncu launches $? $(tail -n1 $OUT/prof_launch.log)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn_kernel -s 4 -c 2 -o $OUT/prof_attn -f python tools/profile_step.py 1 16 > $OUT/prof_attn.log 2>&1; echo "ncu attn $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc -s 8 -c 6 -o $OUT/prof_gemm -f python tools/profile_step.py 1 16 > $OUT/prof_gemm.log 2>&1; echo "ncu gemm $?"
grep -hE "^(FAILED|ERROR)" $OUT/k_fix.log $OUT/u_fix.log | head
ls -la $OUT
