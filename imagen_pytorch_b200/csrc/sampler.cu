// Fused sampler steps on the fp32 NCHW state: classifier-free-guidance combine, x0 prediction, exact
// torch.quantile dynamic thresholding (radix select + lerp), posterior / Heun update, noise injection.
// One CTA per sample; every formula is evaluated op-by-op with round-to-nearest intrinsics (no FMA
// contraction) in the order the reference evaluates it, so that the only difference to the eager
// reference is the U-Net prediction itself.
//
// Reference arithmetic replaced:
//   DDPM : Imagen.p_sample imagen_pytorch.py:2112-2165, p_mean_variance :2085-2110,
//          predict_start_from_noise/_v :308-318, q_posterior :252-270, forward_with_cond_scale :1522
//   EDM  : ElucidatedImagen.one_unet_sample elucidated_imagen.py:481-531,
//          preconditioned_network_forward :340-369, threshold_x_start :309-321
#include "common.cuh"

namespace {

constexpr int SMP_THREADS = 1024;
constexpr int SMP_CACHE_FLOATS = 49152;  // |x0| cached in smem up to 3 x 128 x 128

struct Quant {
  int q_lo, q_hi;
  float q_w;
  int dynamic;  // 1 dynamic thresholding, 0 static clamp(-1, 1)
};

__device__ __forceinline__ float block_reduce_minu(unsigned v, unsigned* sh) {  // returns as float bits
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 16));
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 8));
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 1));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned w = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0xffffffffu;
    w = min(w, __shfl_xor_sync(0xffffffffu, w, 16));
    w = min(w, __shfl_xor_sync(0xffffffffu, w, 8));
    w = min(w, __shfl_xor_sync(0xffffffffu, w, 4));
    w = min(w, __shfl_xor_sync(0xffffffffu, w, 2));
    w = min(w, __shfl_xor_sync(0xffffffffu, w, 1));
    if (threadIdx.x == 0) sh[32] = w;
  }
  __syncthreads();
  const unsigned r = sh[32];
  __syncthreads();
  return __uint_as_float(r);
}

// Exact torch.quantile(|x|, q, dim=-1) for one sample held by one CTA, then .clamp_(min=1).
// F(i) returns |x_i| (non-negative, so the IEEE bit pattern orders like the value).
// sh_hist: 256 + 40 unsigned of shared scratch.
template <class F>
__device__ float block_threshold(F absval, long long n, const Quant& q, unsigned* sh_hist) {
  unsigned* hist = sh_hist;        // [256]
  unsigned* misc = sh_hist + 256;  // [40]
  unsigned prefix = 0, mask = 0;
  unsigned k = (unsigned)q.q_lo;
  for (int pass = 3; pass >= 0; --pass) {
    const int shift = pass * 8;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned u = __float_as_uint(absval(i));
      if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned cum = 0;
      int sel = 255;
      for (int bin = 0; bin < 256; ++bin) {
        const unsigned hcount = hist[bin];
        if (cum + hcount > k) { sel = bin; break; }
        cum += hcount;
      }
      misc[33] = (unsigned)sel;
      misc[34] = k - cum;
    }
    __syncthreads();
    prefix |= misc[33] << shift;
    mask |= 255u << shift;
    k = misc[34];
    __syncthreads();
  }
  const float v_lo = __uint_as_float(prefix);
  float v_hi = v_lo;
  if (q.q_hi != q.q_lo) {
    // next order statistic: same value if duplicates cover rank q_lo+1, else the smallest value above
    if (threadIdx.x == 0) misc[35] = 0;
    __syncthreads();
    unsigned cnt = 0, mn = 0xffffffffu;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned u = __float_as_uint(absval(i));
      if (u <= prefix) ++cnt;
      else mn = min(mn, u);
    }
    atomicAdd(&misc[35], cnt);
    const float mn_f = block_reduce_minu(mn, misc);   // contains __syncthreads
    const unsigned count_le = misc[35];
    v_hi = (count_le >= (unsigned)q.q_lo + 2u) ? v_lo : mn_f;
  }
  // torch.lerp(a, b, w): w < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)
  // (ATen's CUDA lerp kernel is compiled with FMA contraction, so the reference-on-GPU value is the fused one)
  const float diff = __fsub_rn(v_hi, v_lo);
  float s = q.q_w < 0.5f ? __fmaf_rn(q.q_w, diff, v_lo) : __fmaf_rn(-diff, __fsub_rn(1.f, q.q_w), v_hi);
  return fmaxf(s, 1.f);
}

__device__ __forceinline__ float clamp_div(float x0, float s, int dynamic) {
  if (dynamic) return __fdiv_rn(fminf(fmaxf(x0, -s), s), s);
  return fminf(fmaxf(x0, -1.f), 1.f);
}

__device__ __forceinline__ float cfg_combine(const float* pred, long long idx_c, long long idx_n, float cond_scale, bool has_null) {
  const float pc = pred[idx_c];
  if (!has_null) return pc;
  const float pn = pred[idx_n];
  return __fadd_rn(pn, __fmul_rn(__fsub_rn(pc, pn), cond_scale));   // null + (cond - null) * s
}

// ------------------------------------------------------------------------------------------ DDPM

__global__ void __launch_bounds__(SMP_THREADS) ddpm_step_kernel(float* __restrict__ x, const float* __restrict__ pred,
                                                                const float* __restrict__ noise, const b200_ddpm_coef* __restrict__ coefs,
                                                                int* __restrict__ slots, int R, int B, long long chw, float cond_scale,
                                                                int objective, Quant q, float* __restrict__ x_start_out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float smp_cache[];
  __shared__ unsigned sh_hist[256 + 40];
  const int b = blockIdx.x;
  const bool has_null = R > B;
  const b200_ddpm_coef cf = coefs[slots[b]];
  const bool cached = chw <= SMP_CACHE_FLOATS;
  float* xb = x + (long long)b * chw;
  const long long oc = (long long)b * chw, on = (long long)(b + B) * chw;

  auto x0_of = [&](long long i) -> float {
    const float e = cfg_combine(pred, oc + i, on + i, cond_scale, has_null);
    const float xt = xb[i];
    if (objective == 0) return __fdiv_rn(__fsub_rn(xt, __fmul_rn(cf.sigma, e)), fmaxf(cf.alpha, 1e-8f));
    if (objective == 1) return e;
    return __fsub_rn(__fmul_rn(cf.alpha, xt), __fmul_rn(cf.sigma, e));
  };
  if (cached)
    for (long long i = threadIdx.x; i < chw; i += blockDim.x) smp_cache[i] = x0_of(i);
  __syncthreads();
  float s = 1.f;
  if (q.dynamic) {
    if (cached) s = block_threshold([&](long long i) { return fabsf(smp_cache[i]); }, chw, q, sh_hist);
    else s = block_threshold([&](long long i) { return fabsf(x0_of(i)); }, chw, q, sh_hist);
  }
  const float one_minus_c = __fsub_rn(1.f, cf.c);
  for (long long i = threadIdx.x; i < chw; i += blockDim.x) {
    const float x0 = clamp_div(cached ? smp_cache[i] : x0_of(i), s, q.dynamic);
    if (x_start_out != nullptr) x_start_out[oc + i] = x0;      // the next step's self-conditioning input (imagen_pytorch.py:2252)
    const float xt = xb[i];
    // alpha_next * (x_t * (1 - c) / alpha + c * x_start)
    const float mean = __fmul_rn(cf.alpha_next, __fadd_rn(__fdiv_rn(__fmul_rn(xt, one_minus_c), cf.alpha), __fmul_rn(cf.c, x0)));
    xb[i] = __fadd_rn(mean, __fmul_rn(cf.noise_std, noise[oc + i]));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    slots[b] += 1;
    if (has_null) slots[b + B] += 1;
  }
}

// ------------------------------------------------------------------------------------------ EDM

__global__ void __launch_bounds__(SMP_THREADS) edm_phase_kernel(int phase, float* __restrict__ x, float* __restrict__ x_hat,
                                                                float* __restrict__ x1, float* __restrict__ d, float* __restrict__ net_in,
                                                                const float* __restrict__ pred, const float* __restrict__ eps,
                                                                const b200_edm_coef* __restrict__ coefs, int* __restrict__ step_ctr,
                                                                int* __restrict__ slots, int R, int B, long long chw, float cond_scale, Quant q,
                                                                float* __restrict__ denoised_out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float smp_cache[];
  __shared__ unsigned sh_hist[256 + 40];
  const int b = blockIdx.x;
  const bool has_null = R > B;
  const b200_edm_coef cf = coefs[step_ctr[0]];
  const long long oc = (long long)b * chw, on = (long long)(b + B) * chw;
  const bool cached = chw <= SMP_CACHE_FLOATS;

  if (phase == 0) {
    for (long long i = threadIdx.x; i < chw; i += blockDim.x) {
      const float xh = __fadd_rn(x[oc + i], __fmul_rn(cf.noise_coef, __fmul_rn(cf.s_noise, eps[oc + i])));
      x_hat[oc + i] = xh;
      net_in[oc + i] = __fmul_rn(cf.c_in_hat, xh);
    }
    return;
  }
  const float* xin = phase == 1 ? x_hat : x1;
  const float c_skip = phase == 1 ? cf.c_skip_hat : cf.c_skip_next;
  const float c_out = phase == 1 ? cf.c_out_hat : cf.c_out_next;
  auto den_of = [&](long long i) -> float {
    const float f = cfg_combine(pred, oc + i, on + i, cond_scale, has_null);
    return __fadd_rn(__fmul_rn(c_skip, xin[oc + i]), __fmul_rn(c_out, f));
  };
  if (cached)
    for (long long i = threadIdx.x; i < chw; i += blockDim.x) smp_cache[i] = den_of(i);
  __syncthreads();
  float s = 1.f;
  if (q.dynamic) {
    if (cached) s = block_threshold([&](long long i) { return fabsf(smp_cache[i]); }, chw, q, sh_hist);
    else s = block_threshold([&](long long i) { return fabsf(den_of(i)); }, chw, q, sh_hist);
  }
  const float dt = cf.dt;   // fp32(sigma_next - sigma_hat), subtraction done in double on the host like the reference's python floats
  if (phase == 1) {
    for (long long i = threadIdx.x; i < chw; i += blockDim.x) {
      const float D = clamp_div(cached ? smp_cache[i] : den_of(i), s, q.dynamic);
      if (denoised_out != nullptr) denoised_out[oc + i] = D;   // self-conditioning input of the next evaluation (elucidated_imagen.py:518, :538)
      const float xh = x_hat[oc + i];
      const float dd = __fdiv_rn(__fsub_rn(xh, D), cf.sigma_hat);
      const float xn = __fadd_rn(xh, __fmul_rn(dt, dd));
      d[oc + i] = dd;
      if (cf.has_second != 0.f) {
        x1[oc + i] = xn;
        net_in[oc + i] = __fmul_rn(cf.c_in_next, xn);
      } else {
        x[oc + i] = xn;
      }
    }
  } else {
    const float half_dt = cf.half_dt;
    for (long long i = threadIdx.x; i < chw; i += blockDim.x) {
      const float D = clamp_div(cached ? smp_cache[i] : den_of(i), s, q.dynamic);
      if (denoised_out != nullptr) denoised_out[oc + i] = D;
      const float xn1 = x1[oc + i];
      const float dp = __fdiv_rn(__fsub_rn(xn1, D), cf.sigma_next);
      x[oc + i] = __fadd_rn(x_hat[oc + i], __fmul_rn(half_dt, __fadd_rn(d[oc + i], dp)));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    slots[b] += 1;
    if (has_null) slots[b + B] += 1;
    const bool last_phase = (phase == 2) || (phase == 1 && cf.has_second == 0.f);
    if (b == 0 && last_phase) {
      // every CTA read step_ctr at entry; CTA 0 publishes the increment for the NEXT launch
      __threadfence();
      step_ctr[1] = step_ctr[0] + 1;   // staged; committed by the next phase-0 launch (see b200_edm_phase)
    }
  }
}

__global__ void edm_commit_step_kernel(int* step_ctr) { step_ctr[0] = step_ctr[1]; }

__global__ void finalize_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, int flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = (flags & 2) ? x[i] : fminf(fmaxf(x[i], -1.f), 1.f);
  out[i] = (flags & 1) ? __fmul_rn(__fadd_rn(v, 1.f), 0.5f) : v;
}

// RePaint conditioning: where the mask is set, x becomes q_sample(known) = alpha * known + sigma * noise (torch evaluates the two
// products and the sum separately: no FMA contraction here either); elsewhere x is kept.  mask: uint8 [B, HW], shared by the C channels.
__global__ void inpaint_mix_kernel(float* __restrict__ x, const float* __restrict__ known, const uint8_t* __restrict__ mask,
                                   const float* __restrict__ noise, float alpha, float sigma, int C, long long hw, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long b = i / (C * hw), p = i % hw;
  if (!mask[b * hw + p]) return;
  const float nz = noise != nullptr ? __fmul_rn(sigma, noise[i]) : 0.f;
  x[i] = __fadd_rn(__fmul_rn(alpha, known[i]), nz);
}

// q_sample_from_to: x = x * c1 + (noise * c2) / alpha with c1 = alpha_to / alpha, c2 = sigma_to * alpha - sigma * alpha_to
__global__ void renoise_kernel(float* __restrict__ x, const float* __restrict__ noise, float c1, float c2, float alpha, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  x[i] = __fadd_rn(__fmul_rn(x[i], c1), __fdiv_rn(__fmul_rn(noise[i], c2), alpha));
}

int smp_smem(long long chw) { return chw <= SMP_CACHE_FLOATS ? (int)(chw * sizeof(float)) : 0; }

}  // namespace

extern "C" int b200_ddpm_step_sc(float* x, const float* pred, const float* noise, const b200_ddpm_coef* coefs, int32_t* slots, int R, int B,
                                 int64_t chw, float cond_scale, int objective, int thresholding, int32_t q_lo, int32_t q_hi, float q_w,
                                 float* x_start_out, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(x && pred && noise && coefs && slots, "ddpm_step: null pointer");
  B200_REQUIRE(B > 0 && (R == B || R == 2 * B) && chw > 0, "ddpm_step: bad R=%d B=%d", R, B);
  B200_REQUIRE(objective >= 0 && objective <= 2, "ddpm_step: bad objective");
  B200_REQUIRE(!thresholding || (q_lo >= 0 && q_hi >= q_lo && q_hi <= q_lo + 1 && q_hi < chw), "ddpm_step: bad quantile ranks");
  Quant q{q_lo, q_hi, q_w, thresholding};
  const int smem = smp_smem(chw);
  B200_SMEM_OPT_IN(ddpm_step_kernel, SMP_CACHE_FLOATS * 4);   // dynamic cache + 1.2 KB static histogram exceeds the 48 KB default already at 3x64x64
  B200_CUDA_OK(b200_launch(ddpm_step_kernel, dim3(B), dim3(SMP_THREADS), smem, st, x, pred, noise, coefs, slots, R, B, chw, cond_scale, objective, q, x_start_out));
  B200_LAUNCH_OK();
  return B200_OK;
}

extern "C" int b200_ddpm_step(float* x, const float* pred, const float* noise, const b200_ddpm_coef* coefs, int32_t* slots, int R, int B,
                              int64_t chw, float cond_scale, int objective, int thresholding, int32_t q_lo, int32_t q_hi, float q_w,
                              void* stream) {
  return b200_ddpm_step_sc(x, pred, noise, coefs, slots, R, B, chw, cond_scale, objective, thresholding, q_lo, q_hi, q_w, nullptr, stream);
}

extern "C" int b200_edm_phase_sc(int phase, float* x, float* x_hat, float* x1, float* d, float* net_in, const float* pred, const float* eps,
                                 const b200_edm_coef* coefs, int32_t* step_ctr, int32_t* slots, int R, int B, int64_t chw, float cond_scale,
                                 int thresholding, int32_t q_lo, int32_t q_hi, float q_w, float* denoised_out, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(phase >= 0 && phase <= 2, "edm_phase: bad phase");
  B200_REQUIRE(x && x_hat && x1 && d && net_in && coefs && step_ctr && slots, "edm_phase: null pointer");
  B200_REQUIRE(phase == 0 ? eps != nullptr : pred != nullptr, "edm_phase: missing eps/pred");
  B200_REQUIRE(B > 0 && (R == B || R == 2 * B) && chw > 0, "edm_phase: bad R=%d B=%d", R, B);
  Quant q{q_lo, q_hi, q_w, thresholding};
  const int smem = phase == 0 ? 0 : smp_smem(chw);
  B200_SMEM_OPT_IN(edm_phase_kernel, SMP_CACHE_FLOATS * 4);
  if (phase == 0) {
    // commit the step counter staged by the previous step's last phase (step_ctr[1]); step_ctr is int32[2]
    edm_commit_step_kernel<<<1, 1, 0, st>>>(step_ctr);
    B200_LAUNCH_OK();
  }
  edm_phase_kernel<<<B, SMP_THREADS, smem, st>>>(phase, x, x_hat, x1, d, net_in, pred, eps, coefs, step_ctr, slots, R, B, chw, cond_scale, q,
                                                 denoised_out);
  B200_LAUNCH_OK();
  return B200_OK;
}

extern "C" int b200_edm_phase(int phase, float* x, float* x_hat, float* x1, float* d, float* net_in, const float* pred, const float* eps,
                              const b200_edm_coef* coefs, int32_t* step_ctr, int32_t* slots, int R, int B, int64_t chw, float cond_scale,
                              int thresholding, int32_t q_lo, int32_t q_hi, float q_w, void* stream) {
  return b200_edm_phase_sc(phase, x, x_hat, x1, d, net_in, pred, eps, coefs, step_ctr, slots, R, B, chw, cond_scale, thresholding, q_lo, q_hi,
                           q_w, nullptr, stream);
}

extern "C" int b200_inpaint_mix(float* x, const float* known, const uint8_t* mask, const float* noise, float alpha, float sigma, int B, int C,
                                int64_t hw, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(x && known && mask && B > 0 && C > 0 && hw > 0, "inpaint_mix: bad args");
  const long long n = (long long)B * C * hw;
  inpaint_mix_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(x, known, mask, noise, alpha, sigma, C, hw, n);
  B200_LAUNCH_OK();
  return B200_OK;
}

extern "C" int b200_renoise(float* x, const float* noise, float c1, float c2, float alpha, int64_t n, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(x && noise && n > 0, "renoise: bad args");
  renoise_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(x, noise, c1, c2, alpha, n);
  B200_LAUNCH_OK();
  return B200_OK;
}

extern "C" int b200_finalize_images(const float* x, float* out, int64_t n, int unnormalize, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(x && out && n > 0, "finalize: bad args");
  finalize_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(x, out, n, unnormalize);
  B200_LAUNCH_OK();
  return B200_OK;
}
