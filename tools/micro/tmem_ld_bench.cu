// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM vs number of reading warps and loads in flight.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bench tmem_ld_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
template <int INFLIGHT>
__global__ void k(uint32_t* out, int iters, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[INFLIGHT][32];
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int f = 0; f < INFLIGHT; ++f) ld32(base + (uint32_t)(((i * INFLIGHT + f) * 32) & 511 & ~31u) % 480, r[f]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int f = 0; f < INFLIGHT; ++f) acc += r[f][lane & 31 ? 0 : 1] ^ r[f][31];
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
template <int INFLIGHT>
void run(int warps, uint32_t* o, long long* c) {
  const int iters = 2000;
  k<INFLIGHT><<<148, warps * 32>>>(o, iters, c);
  cudaDeviceSynchronize();
  long long cyc; cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
  double bytes = (double)warps * iters * INFLIGHT * 32 * 32 * 4;
  printf("warps %2d  loads in flight %d : %8lld cycles  %.1f B/clk/SM  (%.0f cycles per x32 load per warp)  %s\n", warps, INFLIGHT, cyc, bytes / cyc,
         (double)cyc / (iters * INFLIGHT), cudaGetErrorString(cudaGetLastError()));
}
int main() {
  uint32_t* o; long long* c;
  cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 8);
  for (int w : {4, 8, 16}) { run<1>(w, o, c); run<2>(w, o, c); run<4>(w, o, c); }
  return 0;
}
