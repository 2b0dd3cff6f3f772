#!/bin/bash
# Builds libb200imagen.so for sm_100a (cross-compiles without a GPU). Called by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall"
OBJS=""
pids=()
for f in abi gemm attention attention_tc elementwise cond_f32 sampler; do
  if [ ! -f $f.o ] || [ $f.cu -nt $f.o ] || [ common.cuh -nt $f.o ] || [ ptx.cuh -nt $f.o ] || [ ../../include/b200_imagen.h -nt $f.o ]; then
    $NVCC $FLAGS -c $f.cu -o $f.o &
    pids+=($!)
  fi
  OBJS="$OBJS $f.o"
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libb200imagen.so $OBJS
echo "built $(pwd)/libb200imagen.so"
