#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/k_all.log 2>&1; echo "k_all $? $(tail -n1 $OUT/k_all.log)"; grep -E "^E |^FAILED" $OUT/k_all.log | head -20
timeout 1500 $PYT tests/test_gpu_unet.py -m gpu > $OUT/u_tc.log 2>&1; echo "u_tc $? $(tail -n1 $OUT/u_tc.log)"; grep -E "^E |^FAILED" $OUT/u_tc.log | head -30
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke $? $(tail -n1 $OUT/smoke.log)"
timeout 900 python tools/gemm_bench.py > $OUT/gemm_bench.log 2>&1; grep -E "to_q|plain|ff1|to_out|pixshuf" $OUT/gemm_bench.log
B200_IMAGEN_GEMM_EPI16=1 timeout 900 $PYT tests/test_gpu_kernels.py -m gpu -k 'conv or gemm or linear or attention' > $OUT/k_epi16.log 2>&1; echo "k_epi16 $? $(tail -n1 $OUT/k_epi16.log)"
B200_IMAGEN_GEMM_EPI16=1 timeout 1500 $PYT tests/test_gpu_unet.py -m gpu > $OUT/u_epi16.log 2>&1; echo "u_epi16 $? $(tail -n1 $OUT/u_epi16.log)"; grep -E "^E |^FAILED" $OUT/u_epi16.log | head
B200_IMAGEN_GEMM_EPI16=1 timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_100.log 2>&1; echo "bench100 epi16=1 $?"; grep '^{' $OUT/bench_100.log | cut -c1-200
B200_IMAGEN_GEMM_EPI16=0 timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_100_b.log 2>&1; echo "bench100 epi16=0 $?"; grep '^{' $OUT/bench_100_b.log | cut -c1-200
