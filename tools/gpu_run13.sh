#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 600 $PYT tests/test_gpu_kernels.py -m gpu -k "attention" > $OUT/k_att.log 2>&1; echo "k_att $? $(tail -n1 $OUT/k_att.log)"; grep -E "^E " $OUT/k_att.log | head
timeout 900 python - > $OUT/sweep_attn.log 2>&1 <<'PY'
src = open('tools/sweep_attention.py').read().replace("for var in sorted(NAMES):", "for var in (9, 10):")
exec(compile(src, 'sweep', 'exec'))
PY
cat $OUT/sweep_attn.log
timeout 1500 $PYT tests/test_gpu_unet.py -m gpu > $OUT/u_tc.log 2>&1; echo "u_tc $? $(tail -n1 $OUT/u_tc.log)"; grep -E "^E " $OUT/u_tc.log | head
timeout 600 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_100.log 2>&1; echo "bench100 $?"; grep '^{' $OUT/bench_100.log | cut -c1-200
