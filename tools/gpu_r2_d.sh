#!/bin/bash
# round 2, GPU call D: attention no-subtract / non-ragged specialisation + bottleneck ablations, ncu of the 64x64 attention launch, new API tests
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
SWEEP_VARIANTS=20,40,41,42,43,46,47,50,51,52 timeout 900 python tools/sweep_attention.py > $OUT/d_sweep.txt 2>&1; cat $OUT/d_sweep.txt
timeout 600 $PYT tests/test_gpu_kernels.py -m gpu > $OUT/d_kernels.log 2>&1; echo "kernels $? $(tail -n1 $OUT/d_kernels.log)"; grep -E "^E  |^FAILED" $OUT/d_kernels.log | head -40
timeout 1500 $PYT tests/test_gpu_unet.py tests/test_gpu_baseline_shapes.py -m gpu > $OUT/d_unet.log 2>&1; echo "unet $? $(tail -n1 $OUT/d_unet.log)"; grep -E "^E  |^FAILED" $OUT/d_unet.log | head -40
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_report.json'))
print({k: {a: round(b,4) for a,b in v.items()} for k,v in d.items() if k.startswith('edm')})
PY
timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/d_bench_100.log 2>&1; echo "bench100 $?"; grep '^{' $OUT/d_bench_100.log | cut -c1-200; tail -n 3 $OUT/d_bench_100.log | grep -v '^{' | cut -c1-300
B200_IMAGEN_FA_VARIANT=40 timeout 900 python bench.py --steps 3 --warmup 3 --timesteps 100 --no-cpu-baseline --eager-steps 0 > $OUT/d_bench_100_fa40.log 2>&1; echo "bench100 fa40 $?"; grep '^{' $OUT/d_bench_100_fa40.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn_pt -s 0 -c 1 -o $OUT/d_prof_attn -f python tools/profile_step.py 1 16 > $OUT/d_prof_attn.log 2>&1; echo "ncu attn $?"
B200_IMAGEN_FA_VARIANT=40 timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn_pt -s 0 -c 1 -o $OUT/d_prof_attn40 -f python tools/profile_step.py 1 16 > $OUT/d_prof_attn40.log 2>&1; echo "ncu attn40 $?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_gemm_tc_kernel<128, 5, 1, 16, 1>' -c 6 -o $OUT/d_prof_gemm_norm -f python tools/profile_step.py 1 16 > $OUT/d_prof_gemm_norm.log 2>&1; echo "ncu gemm norm $?"
ls -la $OUT/d_*.ncu-rep
