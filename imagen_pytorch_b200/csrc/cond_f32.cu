// fp32 conditioning head: runs ONCE per sample() call (time MLP tables for every schedule slot,
// text_to_cond, PerceiverResampler, to_text_non_attn_cond, norm_cond, per-layer context K/V).
// All of it is step-invariant and <0.3 % of one forward's FLOPs, so it is kept in
// fp32 SIMT for parity (sin/cos of hundreds of radians, appendix A.4) instead of the bf16 tensor path.
// Reference arithmetic replaced: Unet.forward imagen_pytorch.py:1573-1660, PerceiverAttention :408-445,
// PerceiverResampler :481-498, LearnedSinusoidalPosEmb :664-669, nn.LayerNorm / LayerNorm :331-349.
#include "common.cuh"

namespace {

__device__ __forceinline__ float act_f32(float x, int act) {
  if (act == B200_ACT_SILU) return x / (1.f + expf(-x));
  if (act == B200_ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
  return x;
}

// one warp per output element block: warp computes y[m, n0..n0+3] (4 columns) for one row
__global__ void __launch_bounds__(256) linear_f32_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W,
                                                         const float* __restrict__ bias, int in_act, int out_act,
                                                         const float* __restrict__ res, int ldr, float* __restrict__ y, int ldy,
                                                         long long M, int N, int K) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int ngrp = (N + 3) >> 2;
  if (gw >= M * ngrp) return;
  const long long m = gw / ngrp;
  const int n0 = (int)(gw % ngrp) << 2;
  const float* xr = x + m * ldx;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < K; k += 32) {
    const float xv = act_f32(xr[k], in_act);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + j < N) acc[j] += xv * __ldg(W + (long long)(n0 + j) * K + k);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = warp_sum(acc[j]);
  if (lane < 4 && n0 + lane < N) {
    float v = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
    if (bias != nullptr) v += bias[n0 + lane];
    v = act_f32(v, out_act);
    if (res != nullptr) v += res[m * ldr + n0 + lane];
    y[m * ldy + n0 + lane] = v;
  }
}

__global__ void __launch_bounds__(256) layernorm_f32_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ g,
                                                            const float* __restrict__ beta, float eps, float* __restrict__ y, int ldy,
                                                            long long M, int C) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M) return;
  const float* xr = x + row * ldx;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum(s) / (float)C;
  float vs = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; vs += d * d; }
  const float rstd = rsqrtf(warp_sum(vs) / (float)C + eps);
  for (int c = lane; c < C; c += 32) {
    float v = (xr[c] - mean) * rstd;
    if (g != nullptr) v *= g[c];
    if (beta != nullptr) v += beta[c];
    y[row * ldy + c] = v;
  }
}

__global__ void sinu_pos_emb_kernel(const float* __restrict__ x, const float* __restrict__ w, int M, int half, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int D = 2 * half + 1;
  if (idx >= M * D) return;
  const int m = idx / D, j = idx % D;
  const float xv = x[m];
  float v;
  if (j == 0) v = xv;
  else {
    const int i = (j - 1) % half;
    // freqs = x * w * 2 * pi, evaluated left to right in fp32 like the reference (imagen_pytorch.py:666);
    // full-range sinf/cosf (arguments reach hundreds of radians: no fast-math intrinsics here)
    const float fr = __fmul_rn(__fmul_rn(__fmul_rn(xv, w[i]), 2.f), 3.14159265358979323846f);
    v = (j - 1) < half ? sinf(fr) : cosf(fr);
  }
  out[idx] = v;
}

// one warp per (b, h, query): cosine-sim attention with online softmax over nk keys, head dim 64
__global__ void __launch_bounds__(128) attn_f32_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                                       const float* __restrict__ v, int ldkv, const float* __restrict__ q_scale,
                                                       const float* __restrict__ k_scale, float* __restrict__ o, int ldo, int B, int H,
                                                       int nq, int nk) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= (long long)B * H * nq) return;
  const int i = (int)(gw % nq), h = (int)((gw / nq) % H), b = (int)(gw / ((long long)nq * H));
  const float* qr = q + ((long long)b * nq + i) * ldq + h * 64;
  float q0 = qr[lane], q1 = qr[lane + 32];
  const float qn = fmaxf(sqrtf(warp_sum(q0 * q0 + q1 * q1)), 1e-12f);
  q0 = q0 / qn * q_scale[lane];
  q1 = q1 / qn * q_scale[lane + 32];
  float m = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f;
  for (int j = 0; j < nk; ++j) {
    const float* kr = k + ((long long)b * nk + j) * ldkv + h * 64;
    const float* vr = v + ((long long)b * nk + j) * ldkv + h * 64;
    float k0 = kr[lane], k1 = kr[lane + 32];
    const float kn = fmaxf(sqrtf(warp_sum(k0 * k0 + k1 * k1)), 1e-12f);
    k0 = k0 / kn * k_scale[lane];
    k1 = k1 / kn * k_scale[lane + 32];
    const float s = warp_sum(q0 * k0 + q1 * k1) * 8.f;
    const float m_new = fmaxf(m, s);
    const float sc = expf(m - m_new), p = expf(s - m_new);
    m = m_new;
    l = l * sc + p;
    a0 = a0 * sc + p * vr[lane];
    a1 = a1 * sc + p * vr[lane + 32];
  }
  float* orow = o + ((long long)b * nq + i) * ldo + h * 64;
  orow[lane] = a0 / l;
  orow[lane + 32] = a1 / l;
}

__global__ void __launch_bounds__(256) headnorm_store_kernel(const float* __restrict__ x, int ldx, int col0, int ngroups, int normalize,
                                                             const float* __restrict__ scale, __nv_bfloat16* __restrict__ dst, int rpg,
                                                             long long s_grp, long long s_row, long long s_head, long long M) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= M * ngroups) return;
  const long long row = gw / ngroups;
  const int g = (int)(gw % ngroups);
  const float* xr = x + row * ldx + col0 + g * 64;
  float v0 = xr[lane], v1 = xr[lane + 32];
  if (normalize) {
    const float n = fmaxf(sqrtf(warp_sum(v0 * v0 + v1 * v1)), 1e-12f);
    v0 /= n; v1 /= n;
  }
  if (scale != nullptr) { v0 *= scale[lane]; v1 *= scale[lane + 32]; }
  __nv_bfloat16* d = dst + (row / rpg) * s_grp + (row % rpg) * s_row + (long long)g * s_head;
  d[lane] = __float2bfloat16(v0);
  d[lane + 32] = __float2bfloat16(v1);
}

}  // namespace

extern "C" int b200_linear_f32(const float* x, int32_t ldx, const float* W, const float* b, int in_act, int out_act, const float* residual,
                               int32_t ldr, float* y, int32_t ldy, int64_t M, int32_t N, int32_t K, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(x && W && y && M > 0 && N > 0 && K > 0, "linear_f32: bad args");
  const long long warps = M * ((N + 3) / 4);
  linear_f32_kernel<<<(unsigned)ceil_div64(warps * 32, 256), 256, 0, st>>>(x, ldx, W, b, in_act, out_act, residual, ldr, y, ldy, M, N, K);
  B200_LAUNCH_OK();
  return B200_OK;
}

extern "C" int b200_layernorm_f32(const float* x, int32_t ldx, const float* g, const float* beta, float eps, float* y, int32_t ldy, int64_t M,
                                  int32_t C, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(x && y && M > 0 && C > 0, "layernorm_f32: bad args");
  layernorm_f32_kernel<<<(unsigned)ceil_div64(M * 32, 256), 256, 0, st>>>(x, ldx, g, beta, eps, y, ldy, M, C);
  B200_LAUNCH_OK();
  return B200_OK;
}

extern "C" int b200_sinu_pos_emb(const float* x, const float* w, int M, int half, float* out, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(x && w && out && M > 0 && half > 0, "sinu_pos_emb: bad args");
  const int tot = M * (2 * half + 1);
  sinu_pos_emb_kernel<<<(tot + 127) / 128, 128, 0, st>>>(x, w, M, half, out);
  B200_LAUNCH_OK();
  return B200_OK;
}

extern "C" int b200_attn_f32(const float* q, int32_t ldq, const float* k, const float* v, int32_t ldkv, const float* q_scale,
                             const float* k_scale, float* o, int32_t ldo, int B, int H, int nq, int nk, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(q && k && v && q_scale && k_scale && o && B > 0 && H > 0 && nq > 0 && nk > 0, "attn_f32: bad args");
  const long long warps = (long long)B * H * nq;
  attn_f32_kernel<<<(unsigned)ceil_div64(warps * 32, 128), 128, 0, st>>>(q, ldq, k, v, ldkv, q_scale, k_scale, o, ldo, B, H, nq, nk);
  B200_LAUNCH_OK();
  return B200_OK;
}

extern "C" int b200_headnorm_store(const float* x, int32_t ldx, int32_t col0, int32_t ngroups, int normalize, const float* scale, void* dst,
                                   int32_t rpg, int64_t s_grp, int64_t s_row, int64_t s_head, int64_t M, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(x && dst && ngroups > 0 && rpg > 0 && M > 0, "headnorm_store: bad args");
  const long long warps = M * ngroups;
  headnorm_store_kernel<<<(unsigned)ceil_div64(warps * 32, 256), 256, 0, st>>>(x, ldx, col0, ngroups, normalize, scale,
                                                                                reinterpret_cast<__nv_bfloat16*>(dst), rpg, s_grp, s_row, s_head, M);
  B200_LAUNCH_OK();
  return B200_OK;
}
