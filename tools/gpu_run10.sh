#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 600 $PYT tests/test_gpu_kernels.py -m gpu -k "attention" > $OUT/k_att.log 2>&1; echo "k_att $? $(tail -n1 $OUT/k_att.log)"; grep -E "^E " $OUT/k_att.log | head
timeout 900 python - > $OUT/sweep_attn.log 2>&1 <<'PY'
src = open('tools/sweep_attention.py').read().replace("for var in sorted(NAMES):", "for var in (9,):")
exec(compile(src, 'sweep', 'exec'))
PY
cat $OUT/sweep_attn.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_2gpu.log 2>&1; echo "bench 2gpu $?"; grep '^{' $OUT/bench_2gpu.log | cut -c1-300; tail -n 3 $OUT/bench_2gpu.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 --timesteps 100 --no-cpu-baseline > $OUT/bench_1gpu.log 2>&1; echo "bench 1gpu $?"; grep '^{' $OUT/bench_1gpu.log | cut -c1-200
