#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
B200_IMAGEN_GEMM_SPLITK=1 timeout 900 $PYT tests/test_gpu_kernels.py -m gpu -k "conv or gemm or linear" > $OUT/k_sk.log 2>&1; echo "k_splitk $? $(tail -n1 $OUT/k_sk.log)"; grep -E "^E |^FAILED" $OUT/k_sk.log | head
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu -k "split_k" > $OUT/k_nosk.log 2>&1; echo "k_nosplit $? $(tail -n1 $OUT/k_nosk.log)"
B200_IMAGEN_GEMM_SPLITK=1 timeout 1500 $PYT tests/test_gpu_unet.py -m gpu -k "golden or properties or graph or cfg5" > $OUT/u_sk.log 2>&1; echo "u_splitk $? $(tail -n1 $OUT/u_sk.log)"; grep -E "^E |^FAILED" $OUT/u_sk.log | head
B200_IMAGEN_GEMM_SPLITK=1 timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_sk.log 2>&1; echo "bench splitk=1 $?"; grep '^{' $OUT/bench_sk.log | cut -c1-180
timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_nosk.log 2>&1; echo "bench splitk=0 $?"; grep '^{' $OUT/bench_nosk.log | cut -c1-180
B200_IMAGEN_GEMM_SPLITK=1 timeout 600 python tools/op_times.py > $OUT/op_times_sk.txt 2>&1; grep -E "M=2048 N=1024|M=2048 N=512 K=4608|by entry|conv_gemm" $OUT/op_times_sk.txt | head
timeout 600 python tools/op_times.py > $OUT/op_times_nosk.txt 2>&1; grep -E "M=2048 N=1024|M=2048 N=512 K=4608|by entry|conv_gemm" $OUT/op_times_nosk.txt | head
