#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
timeout 1500 python tools/sweep_attention.py > $OUT/sweep_attn.log 2>&1; cat $OUT/sweep_attn.log
timeout 600 python bench.py --steps 2 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_100.log 2>&1; echo "bench100 $?"; tail -n 1 $OUT/bench_100.log | cut -c1-200
python tools/launch_summary.py $OUT/launches.csv 0 100 2>/dev/null | head -5
