#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
timeout 900 python - > $OUT/sweep_attn.log 2>&1 <<'PY'
src = open('tools/sweep_attention.py').read().replace("for var in sorted(NAMES):", "for var in (10, 12, 13):")
exec(compile(src, 'sweep', 'exec'))
PY
cat $OUT/sweep_attn.log
B200_IMAGEN_FA_VARIANT=13 timeout 600 $PYT tests/test_gpu_kernels.py -m gpu -k "attention" > $OUT/k_att.log 2>&1; echo "k_att(13) $? $(tail -n1 $OUT/k_att.log)"; grep -E "^E " $OUT/k_att.log | head
