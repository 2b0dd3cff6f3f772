// Microbenchmark: MUFU.EX2 throughput, fp32 vs packed bf16x2, on sm_100a.  nvcc -arch=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__global__ void k_f32(float* out, int iters) {
  float a = threadIdx.x * 1e-3f - 3.f, b = a - 1.f, c = a - 2.f, d = a - 3.f;
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
    a -= 1.5f; b -= 1.5f; c -= 1.5f; d -= 1.5f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
__global__ void k_bf16x2(unsigned* out, int iters) {
  unsigned a = 0xc000c040u + threadIdx.x, b = a + 7, c = a + 13, d = a + 29;
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a));
    asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b));
    asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(c));
    asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(d));
    a ^= 0x80008000u; b ^= 0x80008000u; c ^= 0x80008000u; d ^= 0x80008000u;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
int main() {
  float* o; cudaMalloc(&o, 148 * 8 * 1024 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_f32<<<148 * 4, 512>>>(o, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 4 * 512 * iters * 4;
    printf("f32    : %.3f ms  %.1f Gexp/s  (%.2f exp/clk/SM at 1.965 GHz)\n", ms, ops / ms * 1e-6, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaEventRecord(e0); k_bf16x2<<<148 * 4, 512>>>((unsigned*)o, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("bf16x2 : %.3f ms  %.1f Gexp/s  (%.2f exp/clk/SM)  [2 results per instruction]\n", ms, 2 * ops / ms * 1e-6, 2 * ops / (ms * 1e-3) / 148 / 1.965e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
