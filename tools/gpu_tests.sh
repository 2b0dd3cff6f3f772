#!/bin/bash
# full GPU test-suite + short bench (one gpurun call)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 900 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/k_all.log 2>&1; echo "k_all $? $(tail -n1 $OUT/k_all.log)"; grep -E "^E |^FAILED" $OUT/k_all.log | head -20
timeout 1500 $PYT tests/test_gpu_unet.py -m gpu > $OUT/u_tc.log 2>&1; echo "u_tc $? $(tail -n1 $OUT/u_tc.log)"; grep -E "^E |^FAILED" $OUT/u_tc.log | head -30
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke $? $(tail -n1 $OUT/smoke.log)"
timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_100.log 2>&1; echo "bench100 $?"; grep '^{' $OUT/bench_100.log | cut -c1-200
