// HBM-bound row-wise kernels on NHWC bf16 pixel rows: ChanRMSNorm+FiLM+SiLU, LayerNorm(+residual),
// GlobalContext gate, gate*x+residual, layout gathers, per-step time-conditioning plumbing.
// Reference arithmetic replaced: see include/b200_imagen.h next to each entry point.
#include "ptx.cuh"
#include <stdlib.h>

namespace {

constexpr int ROW_THREADS = 256;
constexpr int MAX_VPT = 16;  // 16-byte vectors cached per thread -> C <= 32*16*8 = 4096

// threads per row: aim at ROW_TARGET_VPT vectors (16 B each) per thread so every thread has that many independent loads
// in flight, while small rows still fill the warp (env B200_IMAGEN_ROW_VPT overrides the target for tuning sweeps)
int row_target_vpt() {
  static const int v = [] {
    const char* e = getenv("B200_IMAGEN_ROW_VPT");
    const int x = e ? atoi(e) : 0;
    return (x == 1 || x == 2 || x == 4 || x == 8) ? x : 2;
  }();
  return v;
}
int pick_tpr(int vecs) {
  const int target = row_target_vpt();
  int t = 1;
  while (t * 2 * target <= vecs && t * 2 <= 32) t *= 2;
  return t;
}
int pick_vpt(int vecs, int tpr) {
  const int need = (vecs + tpr - 1) / tpr;
  int v = 1;
  while (v < need) v *= 2;
  return v;
}

// 8 consecutive fp32 (32-byte aligned: every caller indexes at multiples of 8 from a cudaMalloc'd / 32 B-aligned base)
__device__ __forceinline__ void load8f(const float* __restrict__ p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

template <int TPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct RmsParams {
  const __nv_bfloat16* src0;
  const __nv_bfloat16* src1;
  int C0, C1, ld0, ld1;
  float scale1;
  const float* gamma;  // [Ctot], already multiplied by sqrt(Ctot)
  const float* film;   // [B, film_ld] or null
  int film_ld, rows_per_sample;
  __nv_bfloat16* out;
  int ldo;
  long long M;
};

template <int TPR, int VPT>
__global__ void __launch_bounds__(ROW_THREADS) rmsnorm_film_silu_kernel(RmsParams p) {
  pdl_trigger();
  pdl_wait();
  const int rows_per_block = ROW_THREADS / TPR;
  const long long row = (long long)blockIdx.x * rows_per_block + threadIdx.x / TPR;
  const int t = threadIdx.x % TPR;
  const int Ctot = p.C0 + p.C1;
  const int vecs = Ctot >> 3;
  const bool active = row < p.M;
  uint4 buf[VPT];
  float ss = 0.f;
  if (active) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int v = t + i * TPR;
      if (v < vecs) {
        const int c = v << 3;
        float f[8];
        if (c < p.C0) {
          buf[i] = __ldg(reinterpret_cast<const uint4*>(p.src0 + row * p.ld0 + c));
          unpack8(buf[i], f);
        } else {
          buf[i] = __ldg(reinterpret_cast<const uint4*>(p.src1 + row * p.ld1 + (c - p.C0)));
          unpack8(buf[i], f);
        }
        const float sc = c < p.C0 ? 1.f : p.scale1;
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float x = f[j] * sc; ss += x * x; }
      }
    }
  }
  ss = group_sum<TPR>(ss);
  if (!active) return;
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  const float* film = p.film != nullptr ? p.film + (long long)((unsigned)row / (unsigned)p.rows_per_sample) * p.film_ld   /* M < 2^31 checked by the caller: 32-bit divide */ : nullptr;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int v = t + i * TPR;
    if (v < vecs) {
      const int c = v << 3;
      float f[8];
      unpack8(buf[i], f);
      const float sc = (c < p.C0 ? 1.f : p.scale1) * inv;
      float g[8], fs[8], fb[8];
      load8f(p.gamma + c, g);
      if (film != nullptr) { load8f(film + c, fs); load8f(film + Ctot + c, fb); }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = f[j] * sc * g[j];
        if (film != nullptr) y = y * (fs[j] + 1.f) + fb[j];
        f[j] = silu_f(y);
      }
      *reinterpret_cast<uint4*>(p.out + row * p.ldo + c) = pack8(f);
    }
  }
}

struct LnParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* residual;
  const float* g;
  const float* beta;
  __nv_bfloat16* out;
  int ldx, ldr, ldo, C;
  float eps;
  long long M;
};

template <int TPR, int VPT>
__global__ void __launch_bounds__(ROW_THREADS) layernorm_kernel(LnParams p) {
  pdl_trigger();
  pdl_wait();
  const int rows_per_block = ROW_THREADS / TPR;
  const long long row = (long long)blockIdx.x * rows_per_block + threadIdx.x / TPR;
  const int t = threadIdx.x % TPR;
  const int vecs = p.C >> 3;
  const bool active = row < p.M;
  uint4 buf[VPT];
  float s = 0.f;
  if (active) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int v = t + i * TPR;
      if (v < vecs) {
        buf[i] = __ldg(reinterpret_cast<const uint4*>(p.x + row * p.ldx + (v << 3)));
        float f[8];
        unpack8(buf[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += f[j];
      }
    }
  }
  s = group_sum<TPR>(s);
  const float mean = s / (float)p.C;
  float vs = 0.f;
  if (active) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int v = t + i * TPR;
      if (v < vecs) {
        float f[8];
        unpack8(buf[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; vs += d * d; }
      }
    }
  }
  vs = group_sum<TPR>(vs);
  if (!active) return;
  const float rstd = rsqrtf(vs / (float)p.C + p.eps);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int v = t + i * TPR;
    if (v < vecs) {
      const int c = v << 3;
      float f[8];
      unpack8(buf[i], f);
      float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (p.residual != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(p.residual + row * p.ldr + c)), r);
      float g[8], be[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      load8f(p.g + c, g);
      if (p.beta != nullptr) load8f(p.beta + c, be);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd * g[j] + be[j] + r[j];
      *reinterpret_cast<uint4*>(p.out + row * p.ldo + c) = pack8(f);
    }
  }
}

// ---------------------------------------------------------------- chained row kernel (b200_row_chain)
template <int TPR, int VPT>
__global__ void __launch_bounds__(ROW_THREADS) row_chain_kernel(b200_rowchain p) {
  pdl_trigger();
  pdl_wait();
  const int rows_per_block = ROW_THREADS / TPR;
  const long long row = (long long)blockIdx.x * rows_per_block + threadIdx.x / TPR;
  const int t = threadIdx.x % TPR;
  const int vecs = p.C >> 3;
  const bool active = row < p.M;
  const float invC = 1.f / (float)p.C;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x);
  const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(p.residual);
  const int sample = active ? (int)((unsigned)row / (unsigned)(p.rows_per_sample > 0 ? p.rows_per_sample : 1)) : 0;
  float v[VPT][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vv = t + i * TPR;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    if (active && vv < vecs) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * p.ldx + (vv << 3))), v[i]);
      if (p.gate != nullptr) {
        float g[8];
        load8f(p.gate + (long long)sample * p.C + (vv << 3), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] *= g[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  }
  if (p.norm1) {                                   // two-pass LayerNorm in fp32
    const float mean = group_sum<TPR>(s) * invC;
    float vs = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
      if (t + i * TPR < vecs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; vs += d * d; }
      }
    const float rstd = rsqrtf(group_sum<TPR>(vs) * invC + 1e-5f);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vv = t + i * TPR;
      if (vv < vecs) {
        float g[8];
        load8f(p.norm1_g + (vv << 3), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = (v[i][j] - mean) * rstd * g[j];
      }
    }
  }
  float s2 = 0.f, ss2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vv = t + i * TPR;
    if (active && vv < vecs) {
      if (res != nullptr) {
        float r[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(res + row * p.ldr + (vv << 3))), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] += r[j];
      }
      if (p.out != nullptr) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + (vv << 3)) = pack8(v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s2 += v[i][j]; ss2 += v[i][j] * v[i][j]; }
    }
  }
  if (!p.norm2) return;
  s2 = group_sum<TPR>(s2);
  ss2 = group_sum<TPR>(ss2);
  if (!active) return;
  float m2 = 0.f, k2;
  if (p.norm2 == 1) {
    m2 = s2 * invC;
    float vs = 0.f;                                // second pass for the variance: the values are in registers anyway
#pragma unroll
    for (int i = 0; i < VPT; ++i)
      if (t + i * TPR < vecs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - m2; vs += d * d; }
      }
    k2 = rsqrtf(group_sum<TPR>(vs) * invC + 1e-5f);
  } else {
    k2 = 1.f / fmaxf(sqrtf(ss2), 1e-12f);
  }
  const float* film = (p.norm2 == 2 && p.film != nullptr) ? p.film + (long long)sample * p.film_ld : nullptr;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vv = t + i * TPR;
    if (vv < vecs) {
      const int c = vv << 3;
      float g[8], f[8];
      load8f(p.norm2_g + c, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (v[i][j] - m2) * k2 * g[j];
      if (p.norm2 == 2) {
        if (film != nullptr) {
          float fs[8], fb[8];
          load8f(film + c, fs);
          load8f(film + p.C + c, fb);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = f[j] * (fs[j] + 1.f) + fb[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out_norm) + row * p.ld_norm + c) = pack8(f);
    }
  }
}

// ---------------------------------------------------------------- GlobalContext

// logit[row] = x[row, :] . wk + bk   (GlobalContext.to_k, a 1x1 conv to one channel)
template <int TPR, int VPT>
__global__ void __launch_bounds__(ROW_THREADS) gca_logits_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int C, const float* __restrict__ wk,
                                                                 float bk, float* __restrict__ logits, long long M) {
  pdl_trigger();
  pdl_wait();
  const int rows_per_block = ROW_THREADS / TPR;
  const long long row = (long long)blockIdx.x * rows_per_block + threadIdx.x / TPR;
  const int t = threadIdx.x % TPR;
  const int vecs = C >> 3;
  float dot = 0.f;
  if (row < M) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int v = t + i * TPR;
      if (v < vecs) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * ldx + (v << 3))), f);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wk + (v << 3)));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wk + (v << 3) + 4));
        dot += f[0] * w0.x + f[1] * w0.y + f[2] * w0.z + f[3] * w0.w + f[4] * w1.x + f[5] * w1.y + f[6] * w1.z + f[7] * w1.w;
      }
    }
  }
  dot = group_sum<TPR>(dot);
  if (row < M && t == 0) logits[row] = dot + bk;
}

constexpr int GCA_THREADS = 256;
constexpr int GCA_MAX_CHUNK = 1024;   // pixels per chunk held in shared memory

// softmax-weighted channel sums of one pixel chunk of one sample -> (max, sum, acc[C]) partial
__global__ void __launch_bounds__(GCA_THREADS) gca_pool_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int rows_per_sample, int C,
                                                               const float* __restrict__ logits, int nchunk, float* __restrict__ scratch) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sw[GCA_MAX_CHUNK];
  __shared__ float red[GCA_THREADS * 8];
  __shared__ float sred[16];
  const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int ppc = (rows_per_sample + nchunk - 1) / nchunk;
  const int p0 = chunk * ppc, np = max(0, min(rows_per_sample, p0 + ppc) - p0);
  const float* lg = logits + (long long)b * rows_per_sample + p0;
  float* out = scratch + ((long long)b * nchunk + chunk) * (C + 2);
  // chunk max
  float m = -INFINITY;
  for (int p = tid; p < np; p += GCA_THREADS) m = fmaxf(m, lg[p]);
  m = warp_max(m);
  if ((tid & 31) == 0) sred[tid >> 5] = m;
  __syncthreads();
  m = sred[0];
#pragma unroll
  for (int w = 1; w < GCA_THREADS / 32; ++w) m = fmaxf(m, sred[w]);
  __syncthreads();
  // weights + their sum
  float l = 0.f;
  for (int p = tid; p < np; p += GCA_THREADS) {
    const float e = __expf(lg[p] - m);
    sw[p] = e;
    l += e;
  }
  l = warp_sum(l);
  if ((tid & 31) == 0) sred[8 + (tid >> 5)] = l;
  __syncthreads();
  // weighted channel sums: thread = (8-channel vector cv, pixel lane pl)
  const int vecs = C >> 3;
  const int npl = GCA_THREADS / vecs > 0 ? GCA_THREADS / vecs : 1;
  const int cv = tid % vecs, pl = tid / vecs;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (pl < npl && tid < npl * vecs) {
    const __nv_bfloat16* xb = x + ((long long)b * rows_per_sample + p0) * ldx + (cv << 3);
#pragma unroll 4
    for (int p = pl; p < np; p += npl) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(xb + (long long)p * ldx)), f);
      const float w = sw[p];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += w * f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[tid * 8 + j] = acc[j];
  __syncthreads();
  for (int c = tid; c < C; c += GCA_THREADS) {
    const int v = c >> 3, j = c & 7;
    float s = 0.f;
    for (int q = 0; q < npl; ++q) s += red[(q * vecs + v) * 8 + j];
    out[2 + c] = s;
  }
  if (tid == 0) {
    float L = 0.f;
#pragma unroll
    for (int w = 0; w < GCA_THREADS / 32; ++w) L += sred[8 + w];
    out[0] = np > 0 ? m : -INFINITY;
    out[1] = L;
  }
}

// logits + pooling fused: the pixel chunk is staged ONCE in shared memory (the separate logits kernel read the whole tensor a second time):
//   x chunk -> smem;  logit[p] = x[p, :] . wk + bk (warp per pixel);  chunk max / exp weights / sum;  weighted channel sums -> partial
__global__ void __launch_bounds__(GCA_THREADS) gca_pool_fused_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int rows_per_sample, int C,
                                                                     const float* __restrict__ wk, float bk, int nchunk, float* __restrict__ scratch) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint8_t gca_smem[];
  __shared__ float red[GCA_THREADS * 8];
  __shared__ float sred[16];
  const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ppc = (rows_per_sample + nchunk - 1) / nchunk;
  const int p0 = chunk * ppc, np = max(0, min(rows_per_sample, p0 + ppc) - p0);
  const int vecs = C >> 3;
  uint4* xs = reinterpret_cast<uint4*>(gca_smem);                               // [ppc][vecs]
  float* sw = reinterpret_cast<float*>(gca_smem + (size_t)ppc * C * 2);         // [ppc] logits, then weights
  float* out = scratch + ((long long)b * nchunk + chunk) * (C + 2);
  const __nv_bfloat16* xb = x + ((long long)b * rows_per_sample + p0) * ldx;
  for (int i = tid; i < np * vecs; i += GCA_THREADS) {
    const int p = i / vecs, v = i - p * vecs;
    xs[i] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)p * ldx + (v << 3)));
  }
  __syncthreads();
  for (int p = warp; p < np; p += GCA_THREADS / 32) {                           // logits: one warp per pixel
    float dot = 0.f;
    for (int v = lane; v < vecs; v += 32) {
      float f[8];
      unpack8(xs[p * vecs + v], f);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wk + (v << 3)));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wk + (v << 3) + 4));
      dot += f[0] * w0.x + f[1] * w0.y + f[2] * w0.z + f[3] * w0.w + f[4] * w1.x + f[5] * w1.y + f[6] * w1.z + f[7] * w1.w;
    }
    dot = warp_sum(dot);
    if (lane == 0) sw[p] = dot + bk;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int p = tid; p < np; p += GCA_THREADS) m = fmaxf(m, sw[p]);
  m = warp_max(m);
  if (lane == 0) sred[warp] = m;
  __syncthreads();
  m = sred[0];
#pragma unroll
  for (int w = 1; w < GCA_THREADS / 32; ++w) m = fmaxf(m, sred[w]);
  float l = 0.f;
  for (int p = tid; p < np; p += GCA_THREADS) {
    const float e = __expf(sw[p] - m);
    sw[p] = e;
    l += e;
  }
  l = warp_sum(l);
  if (lane == 0) sred[8 + warp] = l;
  __syncthreads();
  const int npl = GCA_THREADS / vecs > 0 ? GCA_THREADS / vecs : 1;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (vecs <= GCA_THREADS) {
    const int cv = tid % vecs, pl = tid / vecs;
    if (pl < npl) {
#pragma unroll 4
      for (int p = pl; p < np; p += npl) {
        float f[8];
        unpack8(xs[p * vecs + cv], f);
        const float w = sw[p];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += w * f[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[tid * 8 + j] = acc[j];
    __syncthreads();
    for (int c = tid; c < C; c += GCA_THREADS) {
      const int v = c >> 3, j = c & 7;
      float s = 0.f;
      for (int q = 0; q < npl; ++q) s += red[(q * vecs + v) * 8 + j];
      out[2 + c] = s;
    }
  } else {                                                                      // C > 2048: every thread walks several channel vectors
    for (int cv = tid; cv < vecs; cv += GCA_THREADS) {
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int p = 0; p < np; ++p) {
        float f[8];
        unpack8(xs[p * vecs + cv], f);
        const float w = sw[p];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += w * f[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) out[2 + (cv << 3) + j] = a[j];
    }
  }
  if (tid == 0) {
    float L = 0.f;
#pragma unroll
    for (int w = 0; w < GCA_THREADS / 32; ++w) L += sred[8 + w];
    out[0] = np > 0 ? m : -INFINITY;
    out[1] = L;
  }
}

// combine + both layers of the gate MLP in ONE launch: a cluster of 8 CTAs serves up to 8 samples.  Every CTA rebuilds the pooled vectors of
// its samples from the chunk partials (tiny), computes its 1/8 slice of the hidden layer for all samples at once (each weight row is
// streamed once per cluster), publishes it in shared memory, and after one cluster barrier reads the other seven slices through
// distributed shared memory to compute its 1/8 slice of the sigmoid gate.  Replaces gca_combine + 2 x gca_mlp (3 launches of ~8 us each).
constexpr int GCA_CL = 8;
constexpr int GCA_TB = 8;    // samples per cluster
__global__ void __cluster_dims__(GCA_CL, 1, 1) __launch_bounds__(256)
gca_tail_kernel(const float* __restrict__ scratch, int nchunk, int C, int hidden, const float* __restrict__ w1, const float* __restrict__ b1,
                const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ gate, int B) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float tail_smem[];
  const int hs = (hidden + GCA_CL - 1) / GCA_CL, cs = (C + GCA_CL - 1) / GCA_CL;
  float* pooled = tail_smem;                       // [GCA_TB][C]
  float* hid_all = pooled + GCA_TB * C;            // [GCA_TB][hs * GCA_CL]
  float* hid_loc = hid_all + GCA_TB * hs * GCA_CL; // [GCA_TB][hs]   (read by the peers)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int b0 = blockIdx.y * GCA_TB;
  const int nb = B - b0 < GCA_TB ? B - b0 : GCA_TB;
  // ---- combine the online-softmax partials of every sample of this cluster (redundantly in each CTA)
  for (int i = tid; i < nb * C; i += 256) {
    const int b = i / C, c = i - b * C;
    const float* sc = scratch + (long long)(b0 + b) * nchunk * (C + 2);
    float M = -INFINITY;
    for (int k = 0; k < nchunk; ++k) M = fmaxf(M, sc[(long long)k * (C + 2)]);
    float L = 0.f, a = 0.f;
    for (int k = 0; k < nchunk; ++k) {
      const float mk = sc[(long long)k * (C + 2)];
      if (mk == -INFINITY) continue;
      const float e = __expf(mk - M);
      L += sc[(long long)k * (C + 2) + 1] * e;
      a += sc[(long long)k * (C + 2) + 2 + c] * e;
    }
    pooled[b * C + c] = a / L;
  }
  __syncthreads();
  // ---- hidden slice: one warp per output row, all samples at once
  for (int n = (int)rank * hs + warp; n < (int)(rank + 1) * hs && n < hidden; n += 8) {
    float acc[GCA_TB];
#pragma unroll
    for (int b = 0; b < GCA_TB; ++b) acc[b] = 0.f;
    const float* wr = w1 + (long long)n * C;
    for (int k = lane * 4; k < C; k += 128) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
      for (int b = 0; b < GCA_TB; ++b)
        if (b < nb) {
          const float4 xv = *reinterpret_cast<const float4*>(pooled + b * C + k);
          acc[b] += w.x * xv.x + w.y * xv.y + w.z * xv.z + w.w * xv.w;
        }
    }
#pragma unroll
    for (int b = 0; b < GCA_TB; ++b) {
      const float s = warp_sum(acc[b]);
      if (lane == 0 && b < nb) hid_loc[b * hs + (n - (int)rank * hs)] = silu_f(s + b1[n]);
    }
  }
  cluster_sync_all();
  // ---- gather the eight hidden slices through distributed shared memory
  for (int i = tid; i < GCA_CL * nb * hs; i += 256) {
    const int rr = i / (nb * hs), rem = i - rr * (nb * hs);
    const int b = rem / hs, j = rem - b * hs;
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(mapa_u32(smem_u32(hid_loc + b * hs + j), (uint32_t)rr)) : "memory");
    hid_all[b * (hs * GCA_CL) + rr * hs + j] = v;
  }
  __syncthreads();
  // ---- gate slice
  for (int c = (int)rank * cs + warp; c < (int)(rank + 1) * cs && c < C; c += 8) {
    float acc[GCA_TB];
#pragma unroll
    for (int b = 0; b < GCA_TB; ++b) acc[b] = 0.f;
    const float* wr = w2 + (long long)c * hidden;
    for (int k = lane * 4; k < hidden; k += 128) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
      for (int b = 0; b < GCA_TB; ++b)
        if (b < nb) {
          const float* xv = hid_all + b * (hs * GCA_CL) + k;
          acc[b] += w.x * xv[0] + w.y * xv[1] + w.z * xv[2] + w.w * xv[3];
        }
    }
#pragma unroll
    for (int b = 0; b < GCA_TB; ++b) {
      const float s = warp_sum(acc[b]);
      if (lane == 0 && b < nb) gate[(long long)(b0 + b) * C + c] = sigmoid_f(s + b2[c]);
    }
  }
  cluster_sync_all();          // no CTA leaves while a peer may still read its hidden slice
}

// combine the per-chunk online-softmax partials -> pooled[b, c]
__global__ void __launch_bounds__(256) gca_combine_kernel(const float* __restrict__ scratch, int nchunk, int C, float* __restrict__ pooled) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const float* sc = scratch + (long long)b * nchunk * (C + 2);
  float M = -INFINITY;
  for (int k = 0; k < nchunk; ++k) M = fmaxf(M, sc[(long long)k * (C + 2)]);
  float L = 0.f, a = 0.f;
  for (int k = 0; k < nchunk; ++k) {
    const float mk = sc[(long long)k * (C + 2)];
    if (mk == -INFINITY) continue;
    const float e = __expf(mk - M);
    L += sc[(long long)k * (C + 2) + 1] * e;
    if (c < C) a += sc[(long long)k * (C + 2) + 2 + c] * e;
  }
  if (c < C) pooled[(long long)b * C + c] = a / L;
}

// y[b, n0..n0+NW) = act(x[b, :] . W[n, :] + bias[n]) for ALL B rows at once: one warp per group of 2 output columns,
// so each weight row is streamed exactly once (the net of GlobalContext is weight-read bound: B rows << C)
template <int MAXB>
__global__ void __launch_bounds__(256) gca_mlp_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                                                      float* __restrict__ y, int B, int N, int K, int act) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  const int b0 = blockIdx.y * MAXB;
  x += (long long)b0 * K;
  y += (long long)b0 * N;
  B = B - b0 < MAXB ? B - b0 : MAXB;
  float acc[MAXB];
#pragma unroll
  for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
  const float* wr = W + (long long)n * K;
#pragma unroll 2
  for (int k = lane * 4; k < K; k += 128) {      // K % 4 == 0 (checked by the caller)
    const float4 w = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
    for (int b = 0; b < MAXB; ++b)
      if (b < B) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)b * K + k));
        acc[b] += w.x * xv.x + w.y * xv.y + w.z * xv.z + w.w * xv.w;
      }
  }
#pragma unroll
  for (int b = 0; b < MAXB; ++b) {
    const float s = warp_sum(acc[b]);
    if (lane == 0 && b < B) {
      const float v = s + bias[n];
      y[(long long)b * N + n] = act == 1 ? silu_f(v) : sigmoid_f(v);
    }
  }
}

__global__ void gate_residual_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ gate,
                                     const __nv_bfloat16* __restrict__ res, int ldr, __nv_bfloat16* __restrict__ out, int ldo,
                                     long long M, int C, int rows_per_sample) {
  pdl_trigger();
  pdl_wait();
  const int vecs = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * vecs) return;
  const long long row = idx / vecs;
  const int c = (int)(idx % vecs) << 3;
  const int b = (int)((unsigned)row / (unsigned)rows_per_sample);   // M < 2^31 (checked by the caller)
  float f[8], r[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * ldx + c)), f);
  unpack8(__ldg(reinterpret_cast<const uint4*>(res + row * ldr + c)), r);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + (long long)b * C + c));
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + (long long)b * C + c + 4));
  f[0] = f[0] * g0.x + r[0]; f[1] = f[1] * g0.y + r[1]; f[2] = f[2] * g0.z + r[2]; f[3] = f[3] * g0.w + r[3];
  f[4] = f[4] * g1.x + r[4]; f[5] = f[5] * g1.y + r[5]; f[6] = f[6] * g1.z + r[6]; f[7] = f[7] * g1.w + r[7];
  *reinterpret_cast<uint4*>(out + row * ldo + c) = pack8(f);
}

// ---------------------------------------------------------------- layout gathers

constexpr int IM2COL_ROWS = 32;

__global__ void __launch_bounds__(256) im2col_init_kernel(const float* __restrict__ img0, int C0, const float* __restrict__ img1, int C1,
                                                          const float* __restrict__ img2, int C2, const float* __restrict__ img3, int C3, int B, int H,
                                                          int W, int ks, __nv_bfloat16* __restrict__ out, int Kpad) {
  pdl_trigger();
  pdl_wait();
  // k -> (dy, dx, channel) decode table, built once per block instead of a div/mod chain per element
  extern __shared__ int lut[];
  const int Cin = C0 + C1 + C2 + C3, pad = ks / 2, K = ks * ks * Cin;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    int v = -1;
    if (k < K) {
      const int tap = k / Cin, c = k - tap * Cin;
      v = ((tap / ks - pad + 64) << 16) | ((tap % ks - pad + 64) << 8) | c;
    }
    lut[k] = v;
  }
  __syncthreads();
  const int vecs = Kpad >> 3;
  const long long M = (long long)B * H * W;
  // IM2COL_ROWS patch rows per block, so the table above is built once per IM2COL_ROWS * Kpad outputs (it used to be rebuilt for
  // every 256 vectors = 2.9 rows, which cost more than the gather).  Consecutive threads = consecutive 16-byte vectors of a patch
  // row: the 92 MB of writes are fully coalesced; the reads gather from a < 1 MB image that lives in L1/L2.
  const long long row_begin = (long long)blockIdx.x * IM2COL_ROWS;
  const int nvec = (int)((M - row_begin < IM2COL_ROWS ? M - row_begin : IM2COL_ROWS) * vecs);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const int kv = i % vecs;
    const long long row = row_begin + i / vecs;
    const int k0 = kv << 3;
    const int w = (int)(row % W), h = (int)((row / W) % H), b = (int)(row / ((long long)W * H));
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = lut[k0 + j];
      float v = 0.f;
      if (e >= 0) {
        const int hh = h + ((e >> 16) & 255) - 64, ww = w + ((e >> 8) & 255) - 64, c = e & 255;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
          v = c < C0 ? __ldg(img0 + (((long long)b * C0 + c) * H + hh) * W + ww)
              : c < C0 + C1 ? __ldg(img1 + (((long long)b * C1 + (c - C0)) * H + hh) * W + ww)
              : c < C0 + C1 + C2 ? __ldg(img2 + (((long long)b * C2 + (c - C0 - C1)) * H + hh) * W + ww)
                                 : __ldg(img3 + (((long long)b * C3 + (c - C0 - C1 - C2)) * H + hh) * W + ww);
        }
      }
      f[j] = v;
    }
    *reinterpret_cast<uint4*>(out + row * Kpad + k0) = pack8(f);
  }
}

// Same gather with the source window staged in shared memory (image width a multiple of 32): a block's 32 patch rows are 32 consecutive
// pixels of ONE image row, so everything they read is a ks x (32 + ks - 1) x Cin window.  It is staged channel-innermost, which makes
// the (dx, c) run of a patch row contiguous; lane i reads 8 consecutive k, i.e. a stride-8 pattern, made conflict-free by the
// a + (a >> 5) skew.  The direct kernel above needs one L1 wavefront per gathered element (each lane hits a different channel plane
// or row): 46 M wavefronts = 112 us for the 15x15x3 stem.
__global__ void __launch_bounds__(256) im2col_init_staged_kernel(const float* __restrict__ img0, int C0, const float* __restrict__ img1, int C1,
                                                                 const float* __restrict__ img2, int C2, const float* __restrict__ img3, int C3, int B,
                                                                 int H, int W, int ks, __nv_bfloat16* __restrict__ out, int Kpad) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ int lut[];                             // [Kpad] window offset of k (or -1), then the skewed window
  const int Cin = C0 + C1 + C2 + C3, pad = ks / 2, K = ks * ks * Cin;
  const int RW = IM2COL_ROWS + ks - 1;                     // window width in pixels
  float* win = reinterpret_cast<float*>(lut + Kpad);
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    int v = -1;
    if (k < K) {
      const int tap = k / Cin, c = k - tap * Cin;
      v = ((tap / ks) * RW + tap % ks) * Cin + c;
    }
    lut[k] = v;
  }
  const long long row_begin = (long long)blockIdx.x * IM2COL_ROWS;
  const int w0 = (int)(row_begin % W), h = (int)((row_begin / W) % H), b = (int)(row_begin / ((long long)W * H));
  const int nwin = ks * RW * Cin;
  for (int i = threadIdx.x; i < nwin; i += blockDim.x) {   // i = (c, dy, col): consecutive threads read consecutive image columns
    const int col = i % RW, dy = (i / RW) % ks, c = i / (RW * ks);
    const int hh = h + dy - pad, ww = w0 + col - pad;
    float v = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
      v = c < C0 ? __ldg(img0 + (((long long)b * C0 + c) * H + hh) * W + ww)
          : c < C0 + C1 ? __ldg(img1 + (((long long)b * C1 + (c - C0)) * H + hh) * W + ww)
          : c < C0 + C1 + C2 ? __ldg(img2 + (((long long)b * C2 + (c - C0 - C1)) * H + hh) * W + ww)
                             : __ldg(img3 + (((long long)b * C3 + (c - C0 - C1 - C2)) * H + hh) * W + ww);
    }
    const int a = (dy * RW + col) * Cin + c;
    win[a + (a >> 5)] = v;
  }
  __syncthreads();
  const int vecs = Kpad >> 3;
  const int nvec = IM2COL_ROWS * vecs;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const int kv = i % vecs, pix = i / vecs;
    const int k0 = kv << 3, base = pix * Cin;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = lut[k0 + j];
      const int a = base + e;
      f[j] = e >= 0 ? win[a + (a >> 5)] : 0.f;
    }
    *reinterpret_cast<uint4*>(out + (row_begin + pix) * Kpad + k0) = pack8(f);
  }
}

__global__ void pixel_unshuffle_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int B, int H, int W, int C,
                                       __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int H2 = H / 2, W2 = W / 2, vecs = (4 * C) >> 3;
  const long long M2 = (long long)B * H2 * W2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M2 * vecs) return;
  const long long orow = idx / vecs;
  const int k = (int)(idx % vecs) << 3;
  const int s = k / C, c = k - s * C;  // s = s1*2 + s2
  const int w2 = (int)(orow % W2), h2 = (int)((orow / W2) % H2), b = (int)(orow / ((long long)W2 * H2));
  const long long irow = ((long long)b * H + 2 * h2 + (s >> 1)) * W + 2 * w2 + (s & 1);
  *reinterpret_cast<uint4*>(out + orow * (4 * C) + k) = __ldg(reinterpret_cast<const uint4*>(x + irow * ldx + c));
}

__global__ void nchw_to_rows_kernel(const float* __restrict__ img, int B, int C, int H, int W, __nv_bfloat16* __restrict__ out, int Cpad) {
  const long long M = (long long)B * H * W;
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  const int w = (int)(row % W), h = (int)((row / W) % H), b = (int)(row / ((long long)W * H));
  for (int c = 0; c < Cpad; ++c) {
    const float v = c < C ? __ldg(img + (((long long)b * C + c) * H + h) * W + w) : 0.f;
    out[row * Cpad + c] = __float2bfloat16(v);
  }
}

// ---------------------------------------------------------------- per-step conditioning plumbing

__global__ void make_time_cond_kernel(const float* __restrict__ table, const float* __restrict__ th, const int* __restrict__ slots, int R,
                                      int D, __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)R * D) return;
  const int r = (int)(idx / D), d = (int)(idx % D);
  const float v = table[(long long)slots[r] * D + d] + th[idx];
  out[idx] = __float2bfloat16(silu_f(v));
}

__global__ void update_time_rows_kernel(const b200_timerow_job* __restrict__ jobs, const int* __restrict__ slots) {
  pdl_trigger();
  pdl_wait();
  const b200_timerow_job j = jobs[blockIdx.x];
  const int b = blockIdx.y;
  const int n = j.rows * j.width;
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(j.table) + (long long)slots[b] * n;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(j.dst) + (long long)b * j.sample_stride;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace

#define LAUNCH_ROW(T, V, KERNEL, grid_rows, ...) \
  B200_CUDA_OK(b200_launch(KERNEL<T, V>, dim3((unsigned)ceil_div64(grid_rows, ROW_THREADS / T)), dim3(ROW_THREADS), 0, st, __VA_ARGS__))
#define DISPATCH_VPT(T, vpt, KERNEL, grid_rows, ...)                               \
  switch (vpt) {                                                                   \
    case 1: LAUNCH_ROW(T, 1, KERNEL, grid_rows, __VA_ARGS__); break;               \
    case 2: LAUNCH_ROW(T, 2, KERNEL, grid_rows, __VA_ARGS__); break;               \
    case 4: LAUNCH_ROW(T, 4, KERNEL, grid_rows, __VA_ARGS__); break;               \
    case 8: LAUNCH_ROW(T, 8, KERNEL, grid_rows, __VA_ARGS__); break;               \
    default: LAUNCH_ROW(T, 16, KERNEL, grid_rows, __VA_ARGS__); break;             \
  }
#define DISPATCH_TPR(tpr, vpt, KERNEL, grid_rows, ...)                             \
  switch (tpr) {                                                                   \
    case 1: DISPATCH_VPT(1, vpt, KERNEL, grid_rows, __VA_ARGS__); break;           \
    case 2: DISPATCH_VPT(2, vpt, KERNEL, grid_rows, __VA_ARGS__); break;           \
    case 4: DISPATCH_VPT(4, vpt, KERNEL, grid_rows, __VA_ARGS__); break;           \
    case 8: DISPATCH_VPT(8, vpt, KERNEL, grid_rows, __VA_ARGS__); break;           \
    case 16: DISPATCH_VPT(16, vpt, KERNEL, grid_rows, __VA_ARGS__); break;         \
    default: DISPATCH_VPT(32, vpt, KERNEL, grid_rows, __VA_ARGS__); break;         \
  }

extern "C" int b200_rmsnorm_film_silu(const b200_src* srcs, int nsrc, float src1_scale, const float* gamma_sqrtC, const float* film,
                                      int32_t film_ld, int32_t rows_per_sample, void* out, int32_t ldo, int64_t M, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(nsrc == 1 || nsrc == 2, "rmsnorm: nsrc must be 1 or 2");
  B200_REQUIRE(M > 0 && M < (1ll << 31) && out && gamma_sqrtC, "rmsnorm: bad args");
  RmsParams p;
  p.src0 = reinterpret_cast<const __nv_bfloat16*>(srcs[0].ptr); p.C0 = srcs[0].C; p.ld0 = srcs[0].ld;
  p.src1 = nsrc == 2 ? reinterpret_cast<const __nv_bfloat16*>(srcs[1].ptr) : nullptr;
  p.C1 = nsrc == 2 ? srcs[1].C : 0; p.ld1 = nsrc == 2 ? srcs[1].ld : 0;
  p.scale1 = src1_scale; p.gamma = gamma_sqrtC; p.film = film; p.film_ld = film_ld;
  p.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : 1;
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ldo = ldo; p.M = M;
  const int Ctot = p.C0 + p.C1;
  B200_REQUIRE((p.C0 & 7) == 0 && (p.C1 & 7) == 0 && (p.ld0 & 7) == 0 && (p.ld1 & 7) == 0 && (ldo & 7) == 0,
               "rmsnorm: channel counts and strides must be multiples of 8 (C0=%d C1=%d)", p.C0, p.C1);
  const int vecs = Ctot >> 3;
  const int tpr = pick_tpr(vecs);
  const int vpt = pick_vpt(vecs, tpr);
  B200_REQUIRE(vpt <= MAX_VPT, "rmsnorm: C=%d too large", Ctot);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(gamma_sqrtC) & 15) == 0 && (film == nullptr || ((reinterpret_cast<uintptr_t>(film) & 15) == 0 && (film_ld & 3) == 0)),
               "rmsnorm: gamma / film must be 16-byte aligned (film_ld %% 4 == 0)");
  DISPATCH_TPR(tpr, vpt, rmsnorm_film_silu_kernel, M, p);
  B200_LAUNCH_OK();
  return B200_OK;
}

extern "C" int b200_layernorm(const void* x, int32_t ldx, const float* g, const float* beta, float eps, const void* residual, int32_t ldr,
                              void* out, int32_t ldo, int64_t M, int32_t C, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(x && g && out && M > 0 && C > 0, "layernorm: bad args");
  B200_REQUIRE((C & 7) == 0 && (ldx & 7) == 0 && (ldo & 7) == 0 && (residual == nullptr || (ldr & 7) == 0), "layernorm: C=%d / strides must be multiples of 8", C);
  LnParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x); p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.g = g; p.beta = beta; p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldx = ldx; p.ldr = ldr; p.ldo = ldo; p.C = C; p.eps = eps; p.M = M;
  const int vecs = C >> 3;
  const int tpr = pick_tpr(vecs);
  const int vpt = pick_vpt(vecs, tpr);
  B200_REQUIRE(vpt <= MAX_VPT, "layernorm: C=%d too large", C);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0 && (beta == nullptr || (reinterpret_cast<uintptr_t>(beta) & 15) == 0), "layernorm: g / beta must be 16-byte aligned");
  DISPATCH_TPR(tpr, vpt, layernorm_kernel, M, p);
  B200_LAUNCH_OK();
  return B200_OK;
}

extern "C" int b200_row_chain(const b200_rowchain* pp, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(pp != nullptr, "row_chain: null descriptor");
  const b200_rowchain p = *pp;
  const auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  B200_REQUIRE(p.x && p.M > 0 && p.M < (1ll << 31) && p.C > 0 && (p.C & 7) == 0 && (p.ldx & 7) == 0 && al16(p.x), "row_chain: bad input (C=%d)", p.C);
  B200_REQUIRE(p.out != nullptr || p.norm2 != 0, "row_chain: no output");
  B200_REQUIRE(p.out == nullptr || ((p.ldo & 7) == 0 && al16(p.out)), "row_chain: bad out");
  B200_REQUIRE(p.residual == nullptr || ((p.ldr & 7) == 0 && al16(p.residual)), "row_chain: bad residual");
  B200_REQUIRE(p.gate == nullptr || (p.rows_per_sample > 0 && al16(p.gate)), "row_chain: gate needs rows_per_sample and a 16-byte aligned table");
  B200_REQUIRE((p.norm1 == 0 || p.norm1 == 1) && p.norm2 >= 0 && p.norm2 <= 2, "row_chain: bad norm kinds %d / %d", p.norm1, p.norm2);
  B200_REQUIRE(p.norm1 == 0 || (p.norm1_g && al16(p.norm1_g)), "row_chain: norm1 gain");
  B200_REQUIRE(p.norm2 == 0 || (p.norm2_g && al16(p.norm2_g) && p.out_norm && al16(p.out_norm) && (p.ld_norm & 7) == 0), "row_chain: norm2 operands");
  B200_REQUIRE(p.film == nullptr || (p.norm2 == 2 && al16(p.film) && (p.film_ld & 3) == 0 && p.rows_per_sample > 0), "row_chain: FiLM operands");
  const int vecs = p.C >> 3;
  const int tpr = pick_tpr(vecs);
  const int vpt = pick_vpt(vecs, tpr);
  B200_REQUIRE(vpt <= MAX_VPT, "row_chain: C=%d too large", p.C);
  DISPATCH_TPR(tpr, vpt, row_chain_kernel, p.M, p);
  return B200_OK;
}

extern "C" int b200_gca_nchunk(int32_t rows_per_sample) {
  int n = rows_per_sample / 256;
  if (n < 1) n = 1;
  if (n > 4096) n = 4096;
  return n;
}

static int gca_fused_bits() {
  static const int fused = [] { const char* ev = getenv("B200_IMAGEN_GCA_FUSED"); return ev ? atoi(ev) : 0; }();
  return fused;
}

extern "C" int b200_gca_chunks(int32_t rows_per_sample, int32_t C) {
  if (!(gca_fused_bits() & 1)) return b200_gca_nchunk(rows_per_sample);   // stand-alone logits + pooling kernels: 256-pixel chunks
  int ppc = 256;
  while (ppc > 8 && (long long)ppc * C * 2 > 32 * 1024) ppc >>= 1;   // the fused pooling kernel stages its pixel chunk in shared memory (several CTAs per SM)
  int n = (rows_per_sample + ppc - 1) / ppc;
  return n < 1 ? 1 : n;
}

extern "C" int b200_gca_gate(const void* x, int32_t ldx, int B, int32_t rows_per_sample, int32_t C, const float* wk, float bk,
                             const float* w1, const float* b1, int32_t hidden, const float* w2, const float* b2, float* scratch,
                             int32_t nchunk, float* gate, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(x && wk && w1 && b1 && w2 && b2 && scratch && gate, "gca: null pointer");
  B200_REQUIRE((C & 7) == 0 && C <= 2048 && (ldx & 7) == 0, "gca: C=%d unsupported", C);
  B200_REQUIRE(nchunk >= 1 && B >= 1 && B <= 65535, "gca: bad nchunk/B");
  B200_REQUIRE((rows_per_sample + nchunk - 1) / nchunk <= GCA_MAX_CHUNK, "gca: chunk of %d pixels too large", (rows_per_sample + nchunk - 1) / nchunk);
  B200_REQUIRE((hidden & 3) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0 && (reinterpret_cast<uintptr_t>(w1) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(w2) & 15) == 0, "gca: hidden %% 4 and 16-byte aligned scratch / weights required");
  // scratch layout: pooled [B*C] | hidden [B*hidden] | partials [B*nchunk*(C+2)] | logits [B*rows_per_sample]
  float* pooled = scratch;
  float* hid = pooled + (long long)B * C;
  float* partials = hid + (long long)B * hidden;
  float* logits = partials + (long long)B * nchunk * (C + 2);
  const long long M = (long long)B * rows_per_sample;
  // B200_IMAGEN_GCA_FUSED bits: 1 = fused logits + pooling kernel (pixel chunk staged once in shared memory), 2 = cluster kernel for
  // combine + MLP + gate.  Default 0: the first version (128 KB chunks, one CTA per SM) measured 68 us per call vs 39 us for the five
  // small kernels (profiles/r02_gca_fused_ab.txt).
  const int fused = gca_fused_bits();
  const int ppc = (rows_per_sample + nchunk - 1) / nchunk;
  const size_t pool_smem = (size_t)ppc * C * 2 + (size_t)ppc * 4;
  const int hs = (hidden + GCA_CL - 1) / GCA_CL;
  const size_t tail_smem = ((size_t)GCA_TB * C + (size_t)GCA_TB * hs * GCA_CL + (size_t)GCA_TB * hs) * 4;
  const bool pool_fused = (fused & 1) && pool_smem <= 64 * 1024;
  const bool tail_fused = (fused & 2) && tail_smem <= 160 * 1024 && (C & 3) == 0;
  const int vecs = C >> 3;
  const int tpr = pick_tpr(vecs), vpt = pick_vpt(vecs, tpr);
  if (pool_fused) {
    B200_SMEM_OPT_IN(gca_pool_fused_kernel, 64 * 1024);
    B200_CUDA_OK(b200_launch(gca_pool_fused_kernel, dim3(nchunk, B), dim3(GCA_THREADS), pool_smem, st, reinterpret_cast<const __nv_bfloat16*>(x), ldx,
                             rows_per_sample, C, wk, bk, nchunk, partials));
  } else {
    DISPATCH_TPR(tpr, vpt, gca_logits_kernel, M, reinterpret_cast<const __nv_bfloat16*>(x), ldx, C, wk, bk, logits, M);
    B200_CUDA_OK(b200_launch(gca_pool_kernel, dim3(nchunk, B), dim3(GCA_THREADS), 0, st, reinterpret_cast<const __nv_bfloat16*>(x), ldx, rows_per_sample, C, logits, nchunk, partials));
  }
  if (tail_fused) {
    B200_SMEM_OPT_IN(gca_tail_kernel, 200 * 1024);
    gca_tail_kernel<<<dim3(GCA_CL, (B + GCA_TB - 1) / GCA_TB), 256, tail_smem, st>>>(partials, nchunk, C, hidden, w1, b1, w2, b2, gate, B);
    B200_LAUNCH_OK();
    return B200_OK;
  }
  B200_CUDA_OK(b200_launch(gca_combine_kernel, dim3((C + 255) / 256, B), dim3(256), 0, st, partials, nchunk, C, pooled));
  const int bg = (B + 7) / 8;
  B200_CUDA_OK(b200_launch(gca_mlp_kernel<8>, dim3((hidden * 32 + 255) / 256, bg), dim3(256), 0, st, pooled, w1, b1, hid, B, hidden, C, 1));
  B200_CUDA_OK(b200_launch(gca_mlp_kernel<8>, dim3((C * 32 + 255) / 256, bg), dim3(256), 0, st, hid, w2, b2, gate, B, C, hidden, 2));
  return B200_OK;
}

extern "C" int b200_gate_residual(const void* x, int32_t ldx, const float* gate, const void* residual, int32_t ldr, void* out, int32_t ldo,
                                  int64_t M, int32_t C, int32_t rows_per_sample, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(x && gate && residual && out && M > 0 && M < (1ll << 31), "gate_residual: null pointer / bad M");
  B200_REQUIRE((C & 7) == 0 && (ldx & 7) == 0 && (ldr & 7) == 0 && (ldo & 7) == 0, "gate_residual: C/strides must be multiples of 8");
  const long long tot = M * (C >> 3);
  B200_CUDA_OK(b200_launch(gate_residual_kernel, dim3((unsigned)ceil_div64(tot, 256)), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(x), ldx, gate,
                                                                        reinterpret_cast<const __nv_bfloat16*>(residual), ldr,
                                                                        reinterpret_cast<__nv_bfloat16*>(out), ldo, M, C, rows_per_sample));
  return B200_OK;
}

extern "C" int b200_im2col_init4(const float* img0, int C0, const float* img1, int C1, const float* img2, int C2, const float* img3, int C3, int B,
                                 int H, int W, int ksize, void* out, int32_t Kpad, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(img0 && out && C0 > 0 && (C1 == 0 || img1) && (C2 == 0 || img2) && (C3 == 0 || img3), "im2col: null pointer");
  B200_REQUIRE((Kpad & 63) == 0 && Kpad >= ksize * ksize * (C0 + C1 + C2 + C3), "im2col: Kpad=%d too small or not a multiple of 64", Kpad);
  B200_REQUIRE(Kpad * 4 <= 48 * 1024 && C0 + C1 + C2 + C3 < 256, "im2col: patch too large");
  const int Cin = C0 + C1 + C2 + C3;
  const int nwin = ksize * (IM2COL_ROWS + ksize - 1) * Cin;
  const size_t staged_smem = (size_t)Kpad * sizeof(int) + (size_t)(nwin + (nwin >> 5) + 1) * sizeof(float);
  static const bool staged_on = [] { const char* e = getenv("B200_IMAGEN_IM2COL_STAGED"); return !e || atoi(e) != 0; }();
  if (staged_on && W % IM2COL_ROWS == 0 && staged_smem <= 160 * 1024) {
    if (staged_smem > 48 * 1024) B200_SMEM_OPT_IN(im2col_init_staged_kernel, 160 * 1024);
    B200_CUDA_OK(b200_launch(im2col_init_staged_kernel, dim3((unsigned)((long long)B * H * W / IM2COL_ROWS)), dim3(256), staged_smem, st, img0, C0, img1, C1, img2, C2, img3, C3, B, H, W, ksize, reinterpret_cast<__nv_bfloat16*>(out), Kpad));
    return B200_OK;
  }
  B200_CUDA_OK(b200_launch(im2col_init_kernel, dim3((unsigned)ceil_div64((long long)B * H * W, IM2COL_ROWS)), dim3(256), Kpad * sizeof(int), st, img0, C0, img1, C1, img2, C2, img3, C3, B, H, W, ksize, reinterpret_cast<__nv_bfloat16*>(out), Kpad));
  return B200_OK;
}

extern "C" int b200_im2col_init3(const float* img0, int C0, const float* img1, int C1, const float* img2, int C2, int B, int H, int W, int ksize,
                                 void* out, int32_t Kpad, void* stream) {
  return b200_im2col_init4(img0, C0, img1, C1, img2, C2, nullptr, 0, B, H, W, ksize, out, Kpad, stream);
}

extern "C" int b200_im2col_init(const float* img0, int C0, const float* img1, int C1, int B, int H, int W, int ksize, void* out, int32_t Kpad,
                                void* stream) {
  return b200_im2col_init3(img0, C0, img1, C1, nullptr, 0, B, H, W, ksize, out, Kpad, stream);
}

extern "C" int b200_pixel_unshuffle(const void* x, int32_t ldx, int B, int H, int W, int C, void* out, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(x && out && (C & 7) == 0 && (ldx & 7) == 0 && (H & 1) == 0 && (W & 1) == 0, "pixel_unshuffle: bad args (C=%d H=%d W=%d)", C, H, W);
  const long long tot = (long long)B * (H / 2) * (W / 2) * ((4 * C) >> 3);
  B200_CUDA_OK(b200_launch(pixel_unshuffle_kernel, dim3((unsigned)ceil_div64(tot, 256)), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(x), ldx, B, H, W, C,
                                                                          reinterpret_cast<__nv_bfloat16*>(out)));
  return B200_OK;
}

extern "C" int b200_nchw_to_rows(const float* img, int B, int C, int H, int W, void* out, int32_t Cpad, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(img && out && Cpad >= C, "nchw_to_rows: bad args");
  const long long M = (long long)B * H * W;
  nchw_to_rows_kernel<<<(unsigned)ceil_div64(M, 256), 256, 0, st>>>(img, B, C, H, W, reinterpret_cast<__nv_bfloat16*>(out), Cpad);
  B200_LAUNCH_OK();
  return B200_OK;
}

extern "C" int b200_make_time_cond(const float* table, const float* text_hiddens, const int32_t* slots, int R, int32_t D, void* out, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(table && text_hiddens && slots && out && R > 0 && D > 0, "make_time_cond: bad args");
  const long long tot = (long long)R * D;
  B200_CUDA_OK(b200_launch(make_time_cond_kernel, dim3((unsigned)ceil_div64(tot, 256)), dim3(256), 0, st, table, text_hiddens, slots, R, D, reinterpret_cast<__nv_bfloat16*>(out)));
  return B200_OK;
}

extern "C" int b200_update_time_rows(const b200_timerow_job* jobs_dev, int njobs, const int32_t* slots, int R, int32_t max_elems, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (njobs == 0) return B200_OK;
  B200_REQUIRE(jobs_dev && slots && R > 0 && R <= 65535, "update_time_rows: bad args");
  int threads = max_elems >= 256 ? 256 : (max_elems >= 128 ? 128 : 64);
  B200_CUDA_OK(b200_launch(update_time_rows_kernel, dim3(dim3(njobs, R)), dim3(threads), 0, st, jobs_dev, slots));
  return B200_OK;
}
