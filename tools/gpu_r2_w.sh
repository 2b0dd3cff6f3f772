#!/bin/bash
# round 2, GPU call W: elect.sync-issued tcgen05 / TMA (no per-instruction ELECT loops) and the split S issuer, against the previous build
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PREV=$PWD/imagen_pytorch_b200/csrc/libb200imagen_prev.so
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/w_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/w_kernels.log)"; grep -E "^E  |^FAILED" $OUT/w_kernels.log | head -30
{ echo "== previous build"; B200_IMAGEN_LIB=$PREV SWEEP_VARIANTS="-1:100" timeout 300 python tools/sweep_attention.py
  echo "== elect.sync build"; SWEEP_VARIANTS="-1:100,110:100,111:100,112:100,113:100,114:100,116:100,69:100,100:100,90:100" timeout 900 python tools/sweep_attention.py; } 2>&1 | tee $OUT/w_attn_sweep.txt
B200_IMAGEN_FA_VARIANT=110 timeout 600 $PYT tests/test_gpu_kernels.py -m gpu -k attention 2>&1 | tail -2
{ echo "== previous build"; B200_IMAGEN_LIB=$PREV timeout 300 python tools/gemm_bench.py child 2>&1 | tail -n 12
  echo "== elect.sync build"; timeout 300 python tools/gemm_bench.py child 2>&1 | tail -n 12; } | tee $OUT/w_gemm_ab.txt
for lib in prev new prev new; do
  if [ $lib = prev ]; then export B200_IMAGEN_LIB=$PREV; else unset B200_IMAGEN_LIB; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/w_bench_$lib.log 2>&1
  echo "bench build=$lib $? $(grep '^{' $OUT/w_bench_$lib.log | cut -c1-140)"
done | tee $OUT/w_step_ab.txt
unset B200_IMAGEN_LIB
B200_IMAGEN_FA_VARIANT=110 timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 2>&1 | grep '^{' | cut -c1-140 | sed 's/^/bench variant=110 /' | tee -a $OUT/w_step_ab.txt
