// Implicit-GEMM convolution / linear layer for sm_100a.
//
//   out[m, n] = epilogue( sum_seg sum_c A_seg[pixel(m) + (dh, dw), c] * Wp[n, k(seg, c)] )
//
// Product path (impl 0): warp-specialised tcgen05 kernel
//   warp 0      : TMA producer.  A tiles are 4-D boxes {64 ch, bw, bh, bb} of the NHWC activation
//                 (bw*bh*bb = 128 output pixels); a 3x3 tap is the same box shifted by (dh, dw) --
//                 out-of-bounds rows are zero-filled by TMA, which *is* the conv zero padding.
//                 B tiles are {64 k, BN n} boxes of the packed weights.  Both land 128B-swizzled.
//   warp 1      : TMEM allocation + single-thread tcgen05.mma issue (M=128, N=BN, K=16 bf16),
//                 fp32 accumulator in TMEM; tcgen05.commit releases smem stages / signals the epilogue.
//   warps 2..5  : epilogue: tcgen05.ld (thread = output row, 64 consecutive columns in registers)
//                 -> bias / SiLU / GELU / per-head L2 norm / residual / pixel-shuffle / NCHW fp32 stores.
// Checker path (impl 1): a plain SIMT fp32 implicit GEMM + the same epilogue code, used only by tests
// to isolate tensor-core/TMA descriptor bugs from epilogue/packing bugs.
//
// Reference arithmetic replaced: see include/b200_imagen.h (b200_conv_gemm).
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int A_TILE_BYTES = BM * BK * 2;

struct SegDev {
  int8_t src, dh, dw, pad;
};

struct EpiDev {
  const float* bias;
  const __nv_bfloat16* residual;
  void* out;
  void* out2;
  const float* l2_scale;
  float out_scale;
  int act, ldr, out_mode, ldc, ldc2, split_col, rows_per_group, group_stride, row_offset, l2_cols, ps_C, dup_rows;
};

struct GemmParams {
  int B, H, W;
  int bw, bh, bb;
  int tiles_w, tiles_h;
  int N, Npad;
  int nseg, total_chunks;
  int nchunks[B200_MAX_SRC];
  SegDev seg[B200_MAX_SEG];
  EpiDev epi;
};

// ------------------------------------------------------------------------------------------ PTX wrappers

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s: a pipeline deadlock becomes an error, not a hang
  }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulation.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128B-swizzled operand tile (rows of 64 bf16 = 128 B; 8-row atoms 1024 B apart).
// Field layout per cute/arch/mma_sm100_desc.hpp (SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) with SWIZZLE_128B = 2.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;             // LBO: unused for swizzled K-major, canonical value 1
  d |= (uint64_t)(1024 >> 4) << 32;   // SBO: 8 rows * 128 B
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}

// ------------------------------------------------------------------------------------------ epilogue

struct RowInfo {
  bool valid;
  int b, h, w;
  long long row;  // (b*H + h)*W + w
};

__device__ __forceinline__ RowInfo tile_row(const GemmParams& p, int tile, int m) {
  const int wblk = tile % p.tiles_w;
  const int hblk = (tile / p.tiles_w) % p.tiles_h;
  const int bblk = tile / (p.tiles_w * p.tiles_h);
  const int lw = m % p.bw, lh = (m / p.bw) % p.bh, lb = m / (p.bw * p.bh);
  RowInfo r;
  r.w = wblk * p.bw + lw;
  r.h = hblk * p.bh + lh;
  r.b = bblk * p.bb + lb;
  r.valid = (r.w < p.W) && (r.h < p.H) && (r.b < p.B);
  r.row = ((long long)r.b * p.H + r.h) * p.W + r.w;
  return r;
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == B200_ACT_SILU) return silu_f(x);
  if (act == B200_ACT_GELU) return gelu_erf_f(x);
  return x;
}

// One thread owns NC consecutive columns [col0, col0+NC) of one output row.
template <int NC>
__device__ __forceinline__ void epi_chunk(const GemmParams& p, const RowInfo& ri, int col0, float* v) {
  const EpiDev& e = p.epi;
  const int N = p.N;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int n = col0 + j;
    float x = v[j];
    if (e.bias != nullptr && n < N) x += __ldg(e.bias + n);
    x = apply_act(x, e.act);
    v[j] = x * e.out_scale;
  }
  if (NC == 64 && col0 < e.l2_cols) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NC; ++j) ss += v[j] * v[j];
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = v[j] * inv * (e.l2_scale != nullptr ? __ldg(e.l2_scale + j) : 1.f);
  }
  if (!ri.valid) return;
  if (e.residual != nullptr) {
    const __nv_bfloat16* rp = e.residual + ri.row * (long long)e.ldr + col0;
#pragma unroll
    for (int j = 0; j < NC; j += 8) {
      if (col0 + j + 8 <= N && (e.ldr & 7) == 0) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp + j));
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int t = 0; t < 8; ++t) v[j + t] += f[t];
      } else {
        for (int t = 0; t < 8; ++t)
          if (col0 + j + t < N) v[j + t] += __bfloat162float(rp[j + t]);
      }
    }
  }
  long long orow = ri.row;
  if (e.rows_per_group > 0)
    orow = (ri.row / e.rows_per_group) * (long long)e.group_stride + e.row_offset + (ri.row % e.rows_per_group);

  if (e.out_mode == B200_OUT_BF16) {
    __nv_bfloat16* base;
    int ld, c;
    if (e.split_col > 0 && col0 >= e.split_col) {
      base = reinterpret_cast<__nv_bfloat16*>(e.out2); ld = e.ldc2; c = col0 - e.split_col;
    } else {
      base = reinterpret_cast<__nv_bfloat16*>(e.out); ld = e.ldc; c = col0;
    }
    for (int rep = 0; rep < (e.dup_rows > 0 ? 2 : 1); ++rep) {
      __nv_bfloat16* dst = base + (orow + (long long)rep * e.dup_rows) * ld + c;
      const bool vec_ok = ((ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
#pragma unroll
      for (int j = 0; j < NC; j += 8) {
        if (vec_ok && col0 + j + 8 <= N) {
          *reinterpret_cast<uint4*>(dst + j) = pack8(v + j);
        } else {
          for (int t = 0; t < 8; ++t)
            if (col0 + j + t < N) dst[j + t] = __float2bfloat16(v[j + t]);
        }
      }
    }
  } else if (e.out_mode == B200_OUT_PIXEL_SHUFFLE) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(e.out);
    const int H2 = 2 * p.H, W2 = 2 * p.W;
#pragma unroll
    for (int j = 0; j < NC; j += 8) {
      const int n = col0 + j;
      if (n >= N) break;
      const int r = n / e.ps_C, c = n - r * e.ps_C;
      const long long prow = ((long long)ri.b * H2 + 2 * ri.h + (r >> 1)) * W2 + 2 * ri.w + (r & 1);
      __nv_bfloat16* dst = out + prow * e.ldc + c;
      if ((e.ps_C & 7) == 0 && (e.ldc & 7) == 0) {
        *reinterpret_cast<uint4*>(dst) = pack8(v + j);
      } else {
        for (int t = 0; t < 8; ++t) {
          const int nn = n + t;
          if (nn < N) {
            const int rr = nn / e.ps_C, cc = nn - rr * e.ps_C;
            const long long pr = ((long long)ri.b * H2 + 2 * ri.h + (rr >> 1)) * W2 + 2 * ri.w + (rr & 1);
            out[pr * e.ldc + cc] = __float2bfloat16(v[j + t]);
          }
        }
      }
    }
  } else if (e.out_mode == B200_OUT_F32_NCHW) {
    float* out = reinterpret_cast<float*>(e.out);
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int n = col0 + j;
      if (n < N) out[(((long long)ri.b * N + n) * p.H + ri.h) * p.W + ri.w] = v[j];
    }
  } else {  // B200_OUT_F32
    float* out = reinterpret_cast<float*>(e.out);
    for (int rep = 0; rep < (e.dup_rows > 0 ? 2 : 1); ++rep) {
      float* dst = out + (orow + (long long)rep * e.dup_rows) * e.ldc + col0;
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (col0 + j < N) dst[j] = v[j];
    }
  }
}

// ------------------------------------------------------------------------------------------ tcgen05 kernel

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, (BN <= 128 ? 2 : 1))
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                    const __grid_constant__ CUtensorMap mapB, const __grid_constant__ GemmParams p) {
  constexpr int B_TILE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr int EPI_NC = BN >= 64 ? 64 : 32;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int n0 = blockIdx.y * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    if (p.nchunks[1] > 0) tma_prefetch_desc(&mapA1);
    if (p.nchunks[2] > 0) tma_prefetch_desc(&mapA2);
    if (p.nchunks[3] > 0) tma_prefetch_desc(&mapA3);
    tma_prefetch_desc(&mapB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer
      const int wblk = tile % p.tiles_w;
      const int hblk = (tile / p.tiles_w) % p.tiles_h;
      const int bblk = tile / (p.tiles_w * p.tiles_h);
      const int w0 = wblk * p.bw, h0 = hblk * p.bh, b0 = bblk * p.bb;
      int stage = 0;
      uint32_t phase = 0;
      int kc = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const int src = p.seg[s].src;
        const CUtensorMap* mA = src == 0 ? &mapA0 : (src == 1 ? &mapA1 : (src == 2 ? &mapA2 : &mapA3));
        const int dh = p.seg[s].dh, dw = p.seg[s].dw;
        const int nch = p.nchunks[src];
        for (int cc = 0; cc < nch; ++cc, ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sA = smem + stage * STAGE_BYTES;
          tma_load_4d(sA, mA, &full_bar[stage], cc * BK, w0 + dw, h0 + dh, b0);
          tma_load_2d(sA + A_TILE_BYTES, &mapB, &full_bar[stage], kc * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer (InstrDescriptor: c_format F32 [4,6)=1, a/b BF16 [7,10)=[10,13)=1,
      // K-major both, N>>3 at [17,23), M>>4 at [24,29))
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < p.total_chunks; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
        const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + A_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (>>4) start-address field
          umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tmem_full_bar);        // accumulator complete
    }
  } else {
    // ---------------- epilogue warps 2..5: TMEM lane quarter = warp % 4
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const RowInfo ri = tile_row(p, tile, m);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    float v[EPI_NC];
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += EPI_NC) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      tmem_ld32(taddr, v);
      if (EPI_NC == 64) tmem_ld32(taddr + 32, v + 32);
      if (n0 + c0 < p.N) epi_chunk<EPI_NC>(p, ri, n0 + c0, v);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ SIMT checker

struct RefSrc {
  const __nv_bfloat16* ptr;
  int C, ld;
};
struct RefParams {
  RefSrc src[B200_MAX_SRC];
};

__global__ void conv_gemm_ref_kernel(RefParams rs, GemmParams p, const __nv_bfloat16* __restrict__ Wp, int Ktot,
                                     float* __restrict__ scratch) {
  const long long M = (long long)p.B * p.H * p.W;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * p.Npad) return;
  const int n = (int)(idx % p.Npad);
  const long long m = idx / p.Npad;
  const int w = (int)(m % p.W), h = (int)((m / p.W) % p.H), b = (int)(m / ((long long)p.W * p.H));
  float acc = 0.f;
  int koff = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const int si = p.seg[s].src;
    const int hh = h + p.seg[s].dh, ww = w + p.seg[s].dw;
    const int C = rs.src[si].C;
    if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) {
      const __nv_bfloat16* a = rs.src[si].ptr + (((long long)b * p.H + hh) * p.W + ww) * rs.src[si].ld;
      const __nv_bfloat16* wr = Wp + (long long)n * Ktot + koff;
      for (int c = 0; c < C; ++c) acc += __bfloat162float(a[c]) * __bfloat162float(wr[c]);
    }
    koff += p.nchunks[si] * BK;
  }
  scratch[m * p.Npad + n] = acc;
}

template <int NC>
__global__ void conv_gemm_ref_epilogue_kernel(GemmParams p, const float* __restrict__ scratch) {
  const long long M = (long long)p.B * p.H * p.W;
  const int chunks = p.Npad / NC;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * chunks) return;
  const int ch = (int)(idx % chunks);
  const long long m = idx / chunks;
  RowInfo ri;
  ri.valid = true;
  ri.w = (int)(m % p.W); ri.h = (int)((m / p.W) % p.H); ri.b = (int)(m / ((long long)p.W * p.H));
  ri.row = m;
  float v[NC];
  for (int j = 0; j < NC; ++j) v[j] = scratch[m * p.Npad + ch * NC + j];
  if (ch * NC < p.N) epi_chunk<NC>(p, ri, ch * NC, v);
}

// ------------------------------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int next_pow2(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

template <int BN, int STAGES>
int launch_tc(const CUtensorMap* maps, const CUtensorMap& mapB, const GemmParams& p, int ntiles, cudaStream_t st) {
  constexpr int smem = STAGES * (A_TILE_BYTES + BN * BK * 2) + 1024 /*align*/ + 256 /*barriers*/;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_OK(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(ntiles, p.Npad / BN);
  conv_gemm_tc_kernel<BN, STAGES><<<grid, GEMM_THREADS, smem, st>>>(maps[0], maps[1], maps[2], maps[3], mapB, p);
  B200_LAUNCH_OK();
  return B200_OK;
}

}  // namespace

extern "C" int b200_conv_gemm_npad(int N) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  return (N + 127) / 128 * 128;
}

extern "C" int b200_conv_gemm(const b200_src* srcs, int nsrc, const b200_seg* segs, int nseg, int B, int H, int W,
                              const void* w_packed, int N, const b200_epilogue* epi, int impl, void* f32_scratch,
                              void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(nsrc >= 1 && nsrc <= B200_MAX_SRC, "conv_gemm: nsrc %d out of range", nsrc);
  B200_REQUIRE(nseg >= 1 && nseg <= B200_MAX_SEG, "conv_gemm: nseg %d out of range", nseg);
  B200_REQUIRE(B > 0 && H > 0 && W > 0 && N > 0, "conv_gemm: bad shape B=%d H=%d W=%d N=%d", B, H, W, N);
  B200_REQUIRE(epi != nullptr && epi->out != nullptr && w_packed != nullptr, "conv_gemm: null pointer");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W;
  p.N = N;
  p.Npad = b200_conv_gemm_npad(N);
  p.nseg = nseg;
  for (int i = 0; i < nsrc; ++i) {
    B200_REQUIRE(srcs[i].ptr != nullptr && srcs[i].C > 0, "conv_gemm: src %d empty", i);
    B200_REQUIRE((srcs[i].ld & 7) == 0 && srcs[i].ld >= srcs[i].C, "conv_gemm: src %d ld=%d must be a multiple of 8 and >= C=%d", i, srcs[i].ld, srcs[i].C);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(srcs[i].ptr) & 15) == 0, "conv_gemm: src %d not 16-byte aligned", i);
    p.nchunks[i] = (srcs[i].C + BK - 1) / BK;
  }
  int total = 0;
  for (int s = 0; s < nseg; ++s) {
    B200_REQUIRE(segs[s].src >= 0 && segs[s].src < nsrc, "conv_gemm: seg %d src out of range", s);
    B200_REQUIRE(segs[s].dh >= -7 && segs[s].dh <= 7 && segs[s].dw >= -7 && segs[s].dw <= 7, "conv_gemm: seg %d tap out of range", s);
    p.seg[s].src = (int8_t)segs[s].src; p.seg[s].dh = (int8_t)segs[s].dh; p.seg[s].dw = (int8_t)segs[s].dw;
    total += p.nchunks[segs[s].src];
  }
  p.total_chunks = total;
  const int Ktot = total * BK;
  // tile box
  p.bw = W >= 128 ? 128 : next_pow2(W);
  p.bh = next_pow2(H) < 128 / p.bw ? next_pow2(H) : 128 / p.bw;
  p.bb = 128 / (p.bw * p.bh);
  p.tiles_w = (W + p.bw - 1) / p.bw;
  p.tiles_h = (H + p.bh - 1) / p.bh;
  const int tiles_b = (B + p.bb - 1) / p.bb;
  const long long ntiles_ll = (long long)p.tiles_w * p.tiles_h * tiles_b;
  B200_REQUIRE(ntiles_ll < (1ll << 31), "conv_gemm: too many tiles");
  const int ntiles = (int)ntiles_ll;
  // epilogue
  EpiDev& e = p.epi;
  e.bias = epi->bias; e.residual = reinterpret_cast<const __nv_bfloat16*>(epi->residual);
  e.out = epi->out; e.out2 = epi->out2; e.l2_scale = epi->l2_scale;
  e.out_scale = epi->out_scale; e.act = epi->act; e.ldr = epi->ldr; e.out_mode = epi->out_mode; e.ldc = epi->ldc;
  e.ldc2 = epi->ldc2; e.split_col = epi->split_col; e.rows_per_group = epi->rows_per_group;
  e.group_stride = epi->group_stride; e.row_offset = epi->row_offset; e.l2_cols = epi->l2_cols; e.ps_C = epi->ps_C;
  e.dup_rows = epi->dup_rows;
  B200_REQUIRE(e.out_mode >= 0 && e.out_mode <= 3, "conv_gemm: bad out_mode %d", e.out_mode);
  B200_REQUIRE(e.split_col % 64 == 0 && e.l2_cols % 64 == 0, "conv_gemm: split_col/l2_cols must be multiples of 64");
  B200_REQUIRE(e.l2_cols == 0 || p.Npad >= 64, "conv_gemm: l2norm epilogue needs N >= 64");
  B200_REQUIRE(e.split_col == 0 || e.out2 != nullptr, "conv_gemm: split_col without out2");
  B200_REQUIRE(e.out_mode != B200_OUT_PIXEL_SHUFFLE || (e.ps_C > 0 && N == 4 * e.ps_C), "conv_gemm: pixel shuffle needs N == 4*ps_C");

  if (impl == 1) {
    B200_REQUIRE(f32_scratch != nullptr, "conv_gemm: SIMT checker needs scratch");
    RefParams rs;
    memset(&rs, 0, sizeof(rs));
    for (int i = 0; i < nsrc; ++i) {
      rs.src[i].ptr = reinterpret_cast<const __nv_bfloat16*>(srcs[i].ptr);
      rs.src[i].C = srcs[i].C; rs.src[i].ld = srcs[i].ld;
    }
    const long long M = (long long)B * H * W;
    const long long tot = M * p.Npad;
    conv_gemm_ref_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(rs, p, reinterpret_cast<const __nv_bfloat16*>(w_packed), Ktot,
                                                                          reinterpret_cast<float*>(f32_scratch));
    B200_LAUNCH_OK();
    if (p.Npad >= 64) {
      const long long t2 = M * (p.Npad / 64);
      conv_gemm_ref_epilogue_kernel<64><<<(unsigned)ceil_div64(t2, 128), 128, 0, st>>>(p, reinterpret_cast<const float*>(f32_scratch));
    } else {
      const long long t2 = M * (p.Npad / 32);
      conv_gemm_ref_epilogue_kernel<32><<<(unsigned)ceil_div64(t2, 128), 128, 0, st>>>(p, reinterpret_cast<const float*>(f32_scratch));
    }
    B200_LAUNCH_OK();
    return B200_OK;
  }

  // ---- tensor maps
  EncodeTiledFn enc = get_encode_fn();
  B200_REQUIRE(enc != nullptr, "conv_gemm: cuTensorMapEncodeTiled not available from the driver");
  int BN;
  if (p.Npad <= 64) BN = p.Npad;
  else if (p.Npad % 256 == 0 && (long long)ntiles * (p.Npad / 256) >= 120) BN = 256;
  else BN = 128;
  CUtensorMap maps[B200_MAX_SRC];
  memset(maps, 0, sizeof(maps));
  for (int i = 0; i < nsrc; ++i) {
    cuuint64_t dims[4] = {(cuuint64_t)srcs[i].C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)srcs[i].ld * 2, (cuuint64_t)srcs[i].ld * 2 * W, (cuuint64_t)srcs[i].ld * 2 * W * H};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(srcs[i].ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "conv_gemm: cuTensorMapEncodeTiled(A%d) failed with %d (C=%d ld=%d W=%d H=%d B=%d box %d,%d,%d)", i, (int)r,
                 srcs[i].C, srcs[i].ld, W, H, B, p.bw, p.bh, p.bb);
  }
  for (int i = nsrc; i < B200_MAX_SRC; ++i) maps[i] = maps[0];
  CUtensorMap mapB;
  {
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)p.Npad};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "conv_gemm: cuTensorMapEncodeTiled(B) failed with %d (Ktot=%d Npad=%d BN=%d)", (int)r, Ktot, p.Npad, BN);
  }
  switch (BN) {
    // stage counts sized so that two CTAs (BN <= 128) share one SM: one CTA's epilogue overlaps the other's mainloop
    case 32: return launch_tc<32, 4>(maps, mapB, p, ntiles, st);
    case 64: return launch_tc<64, 4>(maps, mapB, p, ntiles, st);
    case 128: return launch_tc<128, 3>(maps, mapB, p, ntiles, st);
    default: return launch_tc<256, 4>(maps, mapB, p, ntiles, st);
  }
}
