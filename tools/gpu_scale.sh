#!/bin/bash
# round 2: weak scaling on ONE 8-GPU box (gpurun --gpus 8): N = 1, 8, 2, 4 back to back, per-rank times and clocks in every JSON line
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
TS=${SCALE_TIMESTEPS:-250}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader | head -8
nvidia-smi topo -m 2>/dev/null | head -12
for N in ${SCALE_NS:-1 8 2 4}; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --timesteps $TS --no-cpu-baseline --eager-steps 0 > $OUT/scale_$N.log 2>&1
  else
    NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 3 --warmup 3 --timesteps $TS --no-cpu-baseline --eager-steps 0 > $OUT/scale_$N.log 2>&1
  fi
  echo "N=$N rc=$?"; grep '^{' $OUT/scale_$N.log | tee $OUT/scale_$N.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); pr=d['per_rank']
    print('  value %.3f images/s  ms/step %.1f  e2e %.3f  own sampling ms per rank %s  sm_mhz %s  reasons %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], ['%.1f' % x for x in pr.get('own_sampling_ms_per_step', [])], pr['sm_mhz'], pr['reasons']))
"
  grep -E "NCCL INFO.*(comm 0x.* rank 0 nranks|Init COMPLETE|NVLS|Connected all)" $OUT/scale_$N.log | head -4
  tail -n 2 $OUT/scale_$N.log | grep -v '^{' | cut -c1-200
done
