#!/bin/bash
# One gpurun call = many isolated checks (each in its own process so a sticky CUDA error in one
# stage cannot poison the next); logs land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
run() { # name, timeout, cmd...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $to "$@" > $OUT/$name.log 2>&1
  echo "exit $? : $(tail -n 1 $OUT/$name.log)" | tee -a $OUT/summary.txt
}
: > $OUT/summary.txt
PYT="python -m pytest -q --tb=short -p no:cacheprovider"
run k_other     600 $PYT tests/test_gpu_kernels.py -m gpu -k "not tcgen05 and not simt"
run k_simt      600 $PYT tests/test_gpu_kernels.py -m gpu -k "simt"
run k_tcgen05   600 $PYT tests/test_gpu_kernels.py -m gpu -k "tcgen05"
run u_simt      900 $PYT tests/test_gpu_unet.py -m gpu -k "simt and not equals"
run u_main     1500 $PYT tests/test_gpu_unet.py -m gpu -k "not simt or equals"
run smoke       600 python -c "import __graft_entry__ as g; g.smoke()"
if [ "$1" == "bench" ]; then
  run bench_short 900 python bench.py --steps 1 --warmup 3 --timesteps 50
fi
echo "---- failures"; grep -hE "^(FAILED|ERROR)" $OUT/*.log | head -60
for f in k_other k_simt k_tcgen05 u_simt u_main smoke bench_short; do [ -f $OUT/$f.log ] && { echo "---- tail $f"; tail -n 25 $OUT/$f.log; }; done
cat $OUT/parity_report.json 2>/dev/null
