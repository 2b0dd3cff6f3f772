#!/bin/bash
# round 2, GPU call C: PDL (programmatic dependent launch) A/B, 16-warp fused-norm epilogue, ncu evidence
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 600 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/c_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/c_kernels.log)"; grep -E "^E  |^FAILED" $OUT/c_kernels.log | head -40
timeout 1500 $PYT tests/test_gpu_unet.py tests/test_gpu_baseline_shapes.py -m gpu -x > $OUT/c_unet.log 2>&1; echo "unet $? $(tail -n1 $OUT/c_unet.log)"; grep -E "^E  |^FAILED" $OUT/c_unet.log | head -40
B200_IMAGEN_PDL=0 timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/c_bench_100_nopdl.log 2>&1; echo "bench100 nopdl $?"; grep '^{' $OUT/c_bench_100_nopdl.log | cut -c1-200; tail -n 3 $OUT/c_bench_100_nopdl.log | grep -v '^{' | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/c_bench_100.log 2>&1; echo "bench100 $?"; grep '^{' $OUT/c_bench_100.log | cut -c1-200; tail -n 3 $OUT/c_bench_100.log | grep -v '^{' | cut -c1-300
B200_IMAGEN_FUSE_NORM=0 timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/c_bench_100_nofuse.log 2>&1; echo "bench100 nofuse $?"; grep '^{' $OUT/c_bench_100_nofuse.log | cut -c1-200
timeout 600 python tools/op_times.py > $OUT/c_op_times.txt 2>&1; head -36 $OUT/c_op_times.txt
# every launch of one denoising step with DRAM bytes / time / pipe utilisation (few replay passes)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file $OUT/c_launches.csv python tools/profile_step.py 2 16 > $OUT/c_prof.log 2>&1; echo "ncu launches $? $(wc -l < $OUT/c_launches.csv)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn_pt -s 2 -c 1 -o $OUT/c_prof_attn -f python tools/profile_step.py 1 16 > $OUT/c_prof_attn.log 2>&1; echo "ncu attn $?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_gemm_tc_kernel<128, 5, 1, 16, 1>' -c 8 -o $OUT/c_prof_gemm_norm -f python tools/profile_step.py 1 16 > $OUT/c_prof_gemm_norm.log 2>&1; echo "ncu gemm norm $?"
timeout 900 ncu --set full --clock-control none -k 'regex:layernorm_kernel|rmsnorm_film_silu_kernel|cross_attn_fewkeys_kernel|gca_pool_kernel|gate_residual_kernel|ddpm_step_kernel|im2col_init_kernel' -c 14 -o $OUT/c_prof_hbm -f python tools/profile_step.py 1 16 > $OUT/c_prof_hbm.log 2>&1; echo "ncu hbm $?"
ls -la $OUT/*.ncu-rep
