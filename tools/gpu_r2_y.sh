#!/bin/bash
# round 2, GPU call Y: GEMM tile-width model and transposed kernel after the elect.sync issue fix; MMA-only rates
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
{ for t in 1 0; do B200_IMAGEN_GEMM_T=$t GEMM_BENCH_ONLY="conv3x3" timeout 300 python tools/gemm_bench.py child 2>&1 | sed "s/^/T=$t /"; done
  for t in 1 0; do B200_IMAGEN_GEMM_T=$t B200_IMAGEN_GEMM_DEBUG=5 GEMM_BENCH_ONLY="conv3x3" timeout 300 python tools/gemm_bench.py child 2>&1 | sed "s/^/T=$t MMA-only /"; done; } | tee $OUT/y_gemm.txt
for cfg in "T=1 R=878" "T=0 R=878" "T=1 R=1150" "T=1 R=1400" "T=1 R=878"; do
  set -- $cfg; t=${1#T=}; r=${2#R=}
  B200_IMAGEN_GEMM_T=$t B200_IMAGEN_GEMM_RATE128=$r timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 2>&1 | grep '^{' | cut -c1-140 | sed "s/^/bench $cfg /"
done | tee $OUT/y_step_ab.txt
timeout 600 python tools/op_times.py > $OUT/y_op_times.txt 2>&1; head -30 $OUT/y_op_times.txt
