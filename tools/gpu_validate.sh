#!/bin/bash
# full validation: full GPU suite with the new attention defaults, full-length bench (both arms), ncu --set full of the attention kernels, launch list
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/t_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/t_kernels.log)"; grep -E "^E  |^FAILED" $OUT/t_kernels.log | head -30
timeout 1500 $PYT tests/test_gpu_unet.py tests/test_gpu_baseline_shapes.py -m gpu > $OUT/t_unet.log 2>&1; echo "unet $? $(tail -n1 $OUT/t_unet.log)"; grep -E "^E  |^FAILED" $OUT/t_unet.log | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/t_smoke.log 2>&1; echo "smoke $? $(tail -n2 $OUT/t_smoke.log | tr '\n' ' ' | cut -c1-300)"
timeout 1500 python bench.py > $OUT/t_bench_full.log 2>&1; echo "bench_full $?"; grep '^{' $OUT/t_bench_full.log | tee $OUT/t_bench_full.json | cut -c1-300
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/t_bench_ref.log 2>&1; echo "bench_ref $?"; grep '^{' $OUT/t_bench_ref.log | tee $OUT/t_bench_ref.json | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn_pt_kernel -s 0 -c 1 -o $OUT/t_prof_attn -f python tools/profile_step.py 1 16 > $OUT/t_prof_attn.log 2>&1; echo "ncu attn $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn_ptp_kernel -s 0 -c 1 -o $OUT/t_prof_attn_p -f python tools/profile_step.py 1 16 > $OUT/t_prof_attn_p.log 2>&1; echo "ncu attn persistent $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cross_attn_tc_kernel -s 0 -c 1 -o $OUT/t_prof_xattn -f python tools/profile_step.py 1 16 > $OUT/t_prof_xattn.log 2>&1; echo "ncu cross-attn $?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file $OUT/t_launches.csv python tools/profile_step.py 2 16 > $OUT/t_prof.log 2>&1; echo "ncu launches $? $(wc -l < $OUT/t_launches.csv)"
timeout 600 python tools/op_times.py > $OUT/t_op_times.txt 2>&1; head -12 $OUT/t_op_times.txt
ls -la $OUT/t_*.ncu-rep
