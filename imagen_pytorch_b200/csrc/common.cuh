// Shared device/host helpers for libb200imagen (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/b200_imagen.h"

// ---------------------------------------------------------------- error plumbing (abi.cu)
void b200_set_error(const char* fmt, ...);
#define B200_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) {                                                \
      b200_set_error(__VA_ARGS__);                                \
      return B200_ERR_INVALID;                                    \
    }                                                             \
  } while (0)
#define B200_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      b200_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return B200_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)
#define B200_LAUNCH_OK() B200_CUDA_OK(cudaGetLastError())

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE function attribute: opt in once per (kernel, device), so a process
// that samples on cuda:0 and then on cuda:1 does not launch with the 48 KB default on the second device.
#define B200_SMEM_OPT_IN(func, bytes)                                                                              \
  do {                                                                                                             \
    static unsigned long long done__[2] = {0ull, 0ull};                                                            \
    int dev__ = 0;                                                                                                 \
    B200_CUDA_OK(cudaGetDevice(&dev__));                                                                           \
    if (dev__ < 0 || dev__ >= 128 || !((done__[dev__ >> 6] >> (dev__ & 63)) & 1ull)) {                             \
      B200_CUDA_OK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes)));              \
      if (dev__ >= 0 && dev__ < 128) done__[dev__ >> 6] |= 1ull << (dev__ & 63);                                   \
    }                                                                                                              \
  } while (0)

// SM count of the CURRENT device (cached per device)
static inline int b200_sm_count() {
  static int n[128] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return 148;
  if (n[dev] == 0 && (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0)) n[dev] = 148;
  return n[dev];
}

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// A denoising step is ~250 dependent kernels of ~30 us each: with plain stream order every kernel pays its launch latency, block
// scheduling ramp and prologue (barrier init, TMEM allocation, descriptor prefetch) after the previous grid has drained.  Every
// per-step kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization and executes
//   pdl_trigger();   at its very top   (lets the NEXT grid be scheduled as soon as all CTAs of this one are running)
//   ... local set-up that touches no global memory ...
//   pdl_wait();      before its first global read / write   (returns when the previous grid has completed and flushed)
// so set-up and launch overlap the predecessor's tail.  The edges are captured into the step CUDA graph as programmatic
// dependencies.  Measured on B200 (profiles/r02_pdl_ab.txt): 8.05 ms / step with, 8.00 ms without -- inside a CUDA graph the launch
// gaps are already small -- so the attribute is OFF by default; B200_IMAGEN_PDL=1 enables it (griddepcontrol.wait is a no-op without).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool b200_pdl_enabled();   // abi.cu

template <class... KArgs, class... Args>
static inline cudaError_t b200_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = b200_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- small device helpers
// MUFU.EX2 + MUFU.RCP (2 ulp) instead of the ~10-instruction IEEE division: every consumer rounds the result to bf16
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }
// exact-erf GELU (nn.GELU() default, imagen_pytorch.py:977).  erf via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far
// below the bf16 output rounding) instead of erff(): ~12 FMA-pipe instructions + 2 MUFU instead of ~40 instructions --
// the GEMM epilogue that applies it is issue-bound.
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.f - poly * t * __expf(-z * z);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
