#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
cat /sys/fs/cgroup/cpu.max > $OUT/cpu.txt 2>&1; python -c "import os; print(len(os.sched_getaffinity(0)), os.cpu_count())" >> $OUT/cpu.txt
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu -k "attention and mma" > $OUT/k_att_mma.log 2>&1; echo "k_att_mma $? $(tail -n1 $OUT/k_att_mma.log)"
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu -k "attention and tc" > $OUT/k_att_tc.log 2>&1; echo "k_att_tc $? $(tail -n1 $OUT/k_att_tc.log)"
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu -k "not attention" > $OUT/k_rest.log 2>&1; echo "k_rest $? $(tail -n1 $OUT/k_rest.log)"
B200_IMAGEN_ATTN=mma timeout 1200 $PYT tests/test_gpu_unet.py -m gpu > $OUT/u_mma.log 2>&1; echo "u_mma $? $(tail -n1 $OUT/u_mma.log)"; cp $OUT/parity_report.json $OUT/parity_mma.json
timeout 1200 $PYT tests/test_gpu_unet.py -m gpu > $OUT/u_tc.log 2>&1; echo "u_tc $? $(tail -n1 $OUT/u_tc.log)"
timeout 600 python bench.py --steps 1 --warmup 3 --timesteps 100 > $OUT/bench_100.log 2>&1; echo "bench100 $?"; tail -n 2 $OUT/bench_100.log
B200_IMAGEN_ATTN=mma timeout 600 python bench.py --steps 1 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_100_mma.log 2>&1; echo "bench100 mma $?"; tail -n 1 $OUT/bench_100_mma.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv python tools/profile_step.py 2 16 > $OUT/prof_launch.log 2>&1; echo "ncu launches $? $(tail -n1 $OUT/prof_launch.log)"
grep -hE "^(FAILED|ERROR)" $OUT/k_att_mma.log $OUT/k_att_tc.log $OUT/k_rest.log $OUT/u_mma.log $OUT/u_tc.log | head -40
for f in k_att_tc k_rest u_tc; do echo "--- $f"; grep -E "^E " $OUT/$f.log | head -12; done
cat $OUT/cpu.txt
