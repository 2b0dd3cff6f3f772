#!/bin/bash
# round 2, GPU call J: full-length bench (1000 DDPM steps, both arms), ncu --set full of the transposed conv kernel and of the default attention kernel
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
timeout 1500 python bench.py > $OUT/j_bench_full.log 2>&1; echo "bench_full $?"; grep '^{' $OUT/j_bench_full.log | tee $OUT/j_bench_full.json | cut -c1-400; tail -n 3 $OUT/j_bench_full.log | grep -v '^{' | cut -c1-300
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/j_bench_ref.log 2>&1; echo "bench_ref $?"; grep '^{' $OUT/j_bench_ref.log | tee $OUT/j_bench_ref.json | cut -c1-300
GEMM_BENCH_ONLY="conv3x3 64x64" timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_tcT -s 4 -c 1 -o $OUT/j_prof_gemm_t -f python tools/gemm_bench.py child > $OUT/j_prof_gemm_t.log 2>&1; echo "ncu gemm T $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn_pt -s 0 -c 1 -o $OUT/j_prof_attn -f python tools/profile_step.py 1 16 > $OUT/j_prof_attn.log 2>&1; echo "ncu attn $?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file $OUT/j_launches.csv python tools/profile_step.py 2 16 > $OUT/j_prof.log 2>&1; echo "ncu launches $? $(wc -l < $OUT/j_launches.csv)"
ls -la $OUT/j_*.ncu-rep
