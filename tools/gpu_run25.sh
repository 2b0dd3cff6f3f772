#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python tools/op_times.py > gpurun_out/op_times.txt 2>&1; cat gpurun_out/op_times.txt
