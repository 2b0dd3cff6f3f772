#!/bin/bash
# round 2, GPU call A: full GPU suite (incl. BASELINE-shape parity + fused-norm epilogue), attention variant sweep (P-in-TMEM kernels),
# short bench with / without norm fusion
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 600 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/a_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/a_kernels.log)"; grep -E "^E  |^FAILED" $OUT/a_kernels.log | head -40
B200_IMAGEN_FUSE_NORM=0 timeout 1500 $PYT tests/test_gpu_unet.py tests/test_gpu_baseline_shapes.py -m gpu > $OUT/a_unet_nofuse.log 2>&1; echo "unet(nofuse) $? $(tail -n1 $OUT/a_unet_nofuse.log)"; grep -E "^E  |^FAILED" $OUT/a_unet_nofuse.log | head -40
cp $OUT/parity_report_baseline.json $OUT/parity_report_baseline_nofuse.json 2>/dev/null; cp $OUT/parity_report.json $OUT/parity_report_nofuse.json 2>/dev/null
timeout 1500 $PYT tests/test_gpu_unet.py tests/test_gpu_baseline_shapes.py -m gpu > $OUT/a_unet_fuse.log 2>&1; echo "unet(fuse) $? $(tail -n1 $OUT/a_unet_fuse.log)"; grep -E "^E  |^FAILED" $OUT/a_unet_fuse.log | head -40
echo "--- parity nofuse"; cat $OUT/parity_report_baseline_nofuse.json 2>/dev/null | tr -d '\n ' | cut -c1-1600; echo
echo "--- parity fuse"; cat $OUT/parity_report_baseline.json 2>/dev/null | tr -d '\n ' | cut -c1-1600; echo
SWEEP_VARIANTS=12,20,21,22,23,24,25,26,27,28 timeout 900 python tools/sweep_attention.py > $OUT/a_sweep.txt 2>&1; cat $OUT/a_sweep.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/a_smoke.log 2>&1; echo "smoke $? $(tail -n2 $OUT/a_smoke.log | tr '\n' ' ')"
B200_IMAGEN_FUSE_NORM=0 timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/a_bench_100_nofuse.log 2>&1; echo "bench100 nofuse $?"; grep '^{' $OUT/a_bench_100_nofuse.log | cut -c1-400; tail -n 3 $OUT/a_bench_100_nofuse.log | grep -v '^{' | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/a_bench_100.log 2>&1; echo "bench100 $?"; grep '^{' $OUT/a_bench_100.log | cut -c1-2500; tail -n 3 $OUT/a_bench_100.log | grep -v '^{' | cut -c1-300
timeout 600 python tools/op_times.py > $OUT/a_op_times.txt 2>&1; head -40 $OUT/a_op_times.txt
