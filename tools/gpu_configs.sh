#!/bin/bash
# round 2: one bench line for each of the other BASELINE.json configurations (1 GPU)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out
run() { name=$1; shift; timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eager-steps 0 "$@" > $OUT/cfg_$name.log 2>&1; echo "$name rc=$? $(grep '^{' $OUT/cfg_$name.log | tee $OUT/cfg_$name.json | cut -c1-330)"; tail -n 3 $OUT/cfg_$name.log | grep -v '^{' | cut -c1-300; }
run config2_cs1 --config 2 --cond-scale 1
run config2_cs3 --config 2 --cond-scale 3
run config3 --config 3
run config4_t250 --config 4 --timesteps 250 --steps 2
