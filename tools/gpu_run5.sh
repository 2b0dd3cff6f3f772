#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/k_all.log 2>&1; echo "k_all $? $(tail -n1 $OUT/k_all.log)"
timeout 1200 $PYT tests/test_gpu_unet.py -m gpu > $OUT/u_tc.log 2>&1; echo "u_tc $? $(tail -n1 $OUT/u_tc.log)"
timeout 600 python bench.py --steps 2 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_100.log 2>&1; echo "bench100 $?"; tail -n 1 $OUT/bench_100.log | cut -c1-200; grep -o '"roofline".*' $OUT/bench_100.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv python tools/profile_step.py 2 16 > $OUT/prof_launch.log 2>&1; echo "ncu launches $? $(tail -n1 $OUT/prof_launch.log)"
grep -hE "^(FAILED|ERROR)" $OUT/k_all.log $OUT/u_tc.log | head -40
for f in k_all u_tc; do echo "--- $f"; grep -E "^E " $OUT/$f.log | head -12; done
