#!/bin/bash
# round 2, GPU call F: cluster K-split GEMM, fused GlobalContext kernels, barrier-wait hints (attention + GEMM), A/B benches
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/f_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/f_kernels.log)"; grep -E "^E  |^FAILED" $OUT/f_kernels.log | head -40
for k in 0 1; do GEMM_BENCH_ONLY="conv3x3 8x8" B200_IMAGEN_GEMM_CLUSTER_K=$k timeout 300 python tools/gemm_bench.py child 2>&1 | sed "s/^/clusterK=$k /" | tail -n 2; done | tee $OUT/f_gemm_clusterk_ab.txt
SWEEP_VARIANTS=40:50,40:100,40:200,40:400,42:200,41:200,20:200,46:200 timeout 1200 python tools/sweep_attention.py > $OUT/f_sweep.txt 2>&1; cat $OUT/f_sweep.txt
timeout 1500 $PYT tests/test_gpu_unet.py tests/test_gpu_baseline_shapes.py -m gpu -x > $OUT/f_unet.log 2>&1; echo "unet $? $(tail -n1 $OUT/f_unet.log)"; grep -E "^E  |^FAILED" $OUT/f_unet.log | head -40
run() { name=$1; shift; env "$@" timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/f_bench_$name.log 2>&1; echo "bench100 $name $? $(grep '^{' $OUT/f_bench_$name.log | cut -c1-120)"; tail -n 3 $OUT/f_bench_$name.log | grep -v '^{' | cut -c1-300; }
run base B200_IMAGEN_GEMM_CLUSTER_K=0 B200_IMAGEN_GCA_FUSED=0
run clusterk B200_IMAGEN_GCA_FUSED=0
run gca B200_IMAGEN_GEMM_CLUSTER_K=0
run all X=1
run all_gemmwait200 B200_IMAGEN_GEMM_WAIT_NS=200
run all_noT B200_IMAGEN_GEMM_T=0
timeout 600 python tools/op_times.py > $OUT/f_op_times.txt 2>&1; head -34 $OUT/f_op_times.txt
