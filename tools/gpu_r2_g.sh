#!/bin/bash
# round 2, GPU call G: chained row kernels (LayerNorm / gate-residual + following norm), restored barrier polling in the GEMMs,
# attention default (variant 41, 100 ns waits), GlobalContext variants
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/g_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/g_kernels.log)"; grep -E "^E  |^FAILED" $OUT/g_kernels.log | head -40
for g in 1 2 3; do B200_IMAGEN_GCA_FUSED=$g timeout 300 $PYT tests/test_gpu_kernels.py -m gpu -k global_context > $OUT/g_gca$g.log 2>&1; echo "gca bits=$g $? $(tail -n1 $OUT/g_gca$g.log)"; done
SWEEP_VARIANTS=41:100,41:200,35:100,38:100,40:100 timeout 900 python tools/sweep_attention.py > $OUT/g_sweep.txt 2>&1; cat $OUT/g_sweep.txt
timeout 1500 $PYT tests/test_gpu_unet.py tests/test_gpu_baseline_shapes.py -m gpu -x > $OUT/g_unet.log 2>&1; echo "unet $? $(tail -n1 $OUT/g_unet.log)"; grep -E "^E  |^FAILED" $OUT/g_unet.log | head -40
run() { name=$1; shift; env "$@" timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/g_bench_$name.log 2>&1; echo "bench100 $name $? $(grep '^{' $OUT/g_bench_$name.log | cut -c1-120)"; tail -n 3 $OUT/g_bench_$name.log | grep -v '^{' | cut -c1-300; }
run default X=1
run gca1 B200_IMAGEN_GCA_FUSED=1
run gca2 B200_IMAGEN_GCA_FUSED=2
run gca3 B200_IMAGEN_GCA_FUSED=3
run noT B200_IMAGEN_GEMM_T=0
run default_again X=1
timeout 600 python tools/op_times.py > $OUT/g_op_times.txt 2>&1; head -30 $OUT/g_op_times.txt
