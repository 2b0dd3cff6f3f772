// Flash-style cosine-sim attention, head dim 64 (no N x M score matrix is ever materialised).
//
// Replaces the einsum -> softmax(fp32) -> einsum of Attention.forward (imagen_pytorch.py:565-588,
// multi-query: all 8 heads share one K/V, so the 8*n query rows of a sample are ONE problem against
// [context | null | self] keys) and of CrossAttention.forward (:818-833, per-head K/V, 39 keys).
// Q is L2-normalised and pre-multiplied by q_scale * 8 * log2(e) by the to_q GEMM epilogue and K by
// k_scale, so the kernel computes softmax_2(Q K^T) V with exp2.
//
// v1 data path: cp.async double-buffered K/V tiles (64 keys) in XOR-swizzled smem, ldmatrix +
// mma.sync.m16n8k16 bf16 with fp32 accumulation, online softmax in registers, 8 warps x 16 query
// rows per CTA.  (The tcgen05/TMEM version of this kernel is the next optimisation step; see DESIGN.md.)
#include "common.cuh"

namespace {

constexpr int ATT_BM = 128;
constexpr int ATT_BN = 64;
constexpr int ATT_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// byte offset of 16-byte chunk `c` of row `r` in a [rows][64] bf16 tile with XOR swizzle
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

static inline int sm_count_attn() { return b200_sm_count(); }

struct AttnParams {
  const __nv_bfloat16* q;
  __nv_bfloat16* o;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  long long q_bs, q_hs, kv_bs, kv_hs;
  int q_rs, kv_rs, rows, n_keys;
};

__device__ __forceinline__ void load_kv_tile(uint8_t* sK, uint8_t* sV, const __nv_bfloat16* kb, const __nv_bfloat16* vb, int key0,
                                             int n_keys, int kv_rs, int tid) {
#pragma unroll
  for (int i = 0; i < (ATT_BN * 8) / ATT_THREADS; ++i) {
    const int idx = tid + i * ATT_THREADS;
    const int r = idx >> 3, c = idx & 7;
    const int key = key0 + r;
    if (key < n_keys) {
      cp_async16(sK + swz(r, c), kb + (long long)key * kv_rs + c * 8);
      cp_async16(sV + swz(r, c), vb + (long long)key * kv_rs + c * 8);
    } else {
      *reinterpret_cast<uint4*>(sK + swz(r, c)) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sV + swz(r, c)) = make_uint4(0, 0, 0, 0);
    }
  }
}

__global__ void __launch_bounds__(ATT_THREADS, 2) flash_attn_kernel(AttnParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t att_smem[];
  uint8_t* sQ = att_smem;                       // 128 x 128 B
  uint8_t* sK = sQ + ATT_BM * 128;              // 2 x 64 x 128 B
  uint8_t* sV = sK + 2 * ATT_BN * 128;          // 2 x 64 x 128 B

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * ATT_BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const __nv_bfloat16* qb = p.q + (long long)b * p.q_bs + (long long)h * p.q_hs;
  __nv_bfloat16* ob = p.o + (long long)b * p.q_bs + (long long)h * p.q_hs;
  const __nv_bfloat16* kb = p.k + (long long)b * p.kv_bs + (long long)h * p.kv_hs;
  const __nv_bfloat16* vb = p.v + (long long)b * p.kv_bs + (long long)h * p.kv_hs;

  // ---- Q tile + first K/V tile
#pragma unroll
  for (int i = 0; i < (ATT_BM * 8) / ATT_THREADS; ++i) {
    const int idx = tid + i * ATT_THREADS;
    const int r = idx >> 3, c = idx & 7;
    if (row0 + r < p.rows) cp_async16(sQ + swz(r, c), qb + (long long)(row0 + r) * p.q_rs + c * 8);
    else *reinterpret_cast<uint4*>(sQ + swz(r, c)) = make_uint4(0, 0, 0, 0);
  }
  load_kv_tile(sK, sV, kb, vb, 0, p.n_keys, p.kv_rs, tid);
  cp_async_commit();

  const int ntiles = (p.n_keys + ATT_BN - 1) / ATT_BN;
  const int g = lane >> 2, tq = lane & 3;

  uint32_t qf[4][4];
  float o_acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o_acc[j][0] = o_acc[j][1] = o_acc[j][2] = o_acc[j][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      load_kv_tile(sK + (buf ^ 1) * ATT_BN * 128, sV + (buf ^ 1) * ATT_BN * 128, kb, vb, (t + 1) * ATT_BN, p.n_keys, p.kv_rs, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
      const uint32_t qbase = smem_u32(sQ);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = 2 * ks + (lane >> 4);
        ldsm_x4(qbase + swz(r, c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const uint32_t kbase = smem_u32(sK + buf * ATT_BN * 128);
    const uint32_t vbase = smem_u32(sV + buf * ATT_BN * 128);

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b0, b1, b2, b3;
        const int r = 8 * (2 * jp + (lane >> 4)) + (lane & 7);
        const int c = 2 * ks + ((lane >> 3) & 1);
        ldsm_x4(kbase + swz(r, c), b0, b1, b2, b3);
        mma_bf16_16816(s[2 * jp], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
        mma_bf16_16816(s[2 * jp + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
      }
    }
    // ---- mask the ragged last tile
    const int key0 = t * ATT_BN;
    if (key0 + ATT_BN > p.n_keys) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kidx = key0 + 8 * j + 2 * tq;
        if (kidx >= p.n_keys) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (kidx + 1 >= p.n_keys) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    // ---- online softmax (rows g and g+8 of this warp's 16)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    }
    float alpha[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float m_new = fmaxf(m_run[i], mx[i]);
      alpha[i] = exp2f(m_run[i] - m_new);   // m_run = -inf on the first tile -> 0
      m_run[i] = m_new;
      l_run[i] *= alpha[i];
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] - m_run[0]), p1 = exp2f(s[j][1] - m_run[0]);
      const float p2 = exp2f(s[j][2] - m_run[1]), p3 = exp2f(s[j][3] - m_run[1]);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      const int kk = j >> 1;
      if ((j & 1) == 0) { pf[kk][0] = pack_bf16x2(p0, p1); pf[kk][1] = pack_bf16x2(p2, p3); }
      else              { pf[kk][2] = pack_bf16x2(p0, p1); pf[kk][3] = pack_bf16x2(p2, p3); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o_acc[j][0] *= alpha[0]; o_acc[j][1] *= alpha[0];
      o_acc[j][2] *= alpha[1]; o_acc[j][3] *= alpha[1];
    }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b0, b1, b2, b3;
        const int r = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = 2 * jp + (lane >> 4);
        ldsm_x4_trans(vbase + swz(r, c), b0, b1, b2, b3);
        mma_bf16_16816(o_acc[2 * jp], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b0, b1);
        mma_bf16_16816(o_acc[2 * jp + 1], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b2, b3);
      }
    }
    __syncthreads();   // everyone done with buf before it is refilled at t+2
  }

  // ---- normalise and store
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 1);
    l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int r_lo = row0 + warp * 16 + g, r_hi = r_lo + 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = 8 * j + 2 * tq;
    if (r_lo < p.rows) *reinterpret_cast<uint32_t*>(ob + (long long)r_lo * p.q_rs + col) = pack_bf16x2(o_acc[j][0] * inv0, o_acc[j][1] * inv0);
    if (r_hi < p.rows) *reinterpret_cast<uint32_t*>(ob + (long long)r_hi * p.q_rs + col) = pack_bf16x2(o_acc[j][2] * inv1, o_acc[j][3] * inv1);
  }
}


// ---- few-keys variant (cross-attention over the 39 context rows: n_keys <= 64, one K/V tile) ---------------------------------
// The generic kernel above spends one short-lived CTA per 128 query rows (load -> compute -> 4-byte scattered stores): 126 us for
// the 268 MB of q + o at the 64x64 level.  Here a CTA keeps K/V of its (sample, head) in shared memory, walks a contiguous range
// of query tiles with a 3-deep cp.async ring (prefetch distance 2), needs no online rescaling (single tile), and stages O through
// the query buffer so every store instruction writes whole 128-byte rows.
constexpr int XA_QBUF = 3;

__global__ void __launch_bounds__(ATT_THREADS, 2) cross_attn_fewkeys_kernel(AttnParams p, int tiles_per_cta) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t att_smem[];
  uint8_t* sK = att_smem;                        // 64 x 128 B
  uint8_t* sV = sK + ATT_BN * 128;               // 64 x 128 B
  uint8_t* sQ = sV + ATT_BN * 128;               // XA_QBUF x 128 x 128 B

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const __nv_bfloat16* qb = p.q + (long long)b * p.q_bs + (long long)h * p.q_hs;
  __nv_bfloat16* ob = p.o + (long long)b * p.q_bs + (long long)h * p.q_hs;
  const __nv_bfloat16* kb = p.k + (long long)b * p.kv_bs + (long long)h * p.kv_hs;
  const __nv_bfloat16* vb = p.v + (long long)b * p.kv_bs + (long long)h * p.kv_hs;
  const int ntiles = (p.rows + ATT_BM - 1) / ATT_BM;
  const int t0 = blockIdx.x * tiles_per_cta;
  const int t1 = t0 + tiles_per_cta < ntiles ? t0 + tiles_per_cta : ntiles;
  if (t0 >= t1) return;

  auto load_q = [&](int tile, int buf) {
    uint8_t* dst = sQ + buf * ATT_BM * 128;
    const int row0 = tile * ATT_BM;
#pragma unroll
    for (int i = 0; i < (ATT_BM * 8) / ATT_THREADS; ++i) {
      const int idx = tid + i * ATT_THREADS;
      const int r = idx >> 3, c = idx & 7;
      if (row0 + r < p.rows) cp_async16(dst + swz(r, c), qb + (long long)(row0 + r) * p.q_rs + c * 8);
      else *reinterpret_cast<uint4*>(dst + swz(r, c)) = make_uint4(0, 0, 0, 0);
    }
  };

  load_kv_tile(sK, sV, kb, vb, 0, p.n_keys, p.kv_rs, tid);
  load_q(t0, 0);
  cp_async_commit();
  if (t0 + 1 < t1) load_q(t0 + 1, 1);
  cp_async_commit();

  const int g = lane >> 2, tq = lane & 3;
  const uint32_t kbase = smem_u32(sK), vbase = smem_u32(sV);
  for (int t = t0; t < t1; ++t) {
    const int buf = (t - t0) % XA_QBUF;
    cp_async_wait<1>();          // tile t (and K/V) landed; tile t+1 may still be in flight
    __syncthreads();             // ... for every thread; also: everyone is done with the buffer refilled below
    if (t + 2 < t1) load_q(t + 2, (t - t0 + 2) % XA_QBUF);
    cp_async_commit();

    uint8_t* sQb = sQ + buf * ATT_BM * 128;
    const uint32_t qbase = smem_u32(sQb);
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int c = 2 * ks + (lane >> 4);
      ldsm_x4(qbase + swz(r, c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b0, b1, b2, b3;
        const int r = 8 * (2 * jp + (lane >> 4)) + (lane & 7);
        const int c = 2 * ks + ((lane >> 3) & 1);
        ldsm_x4(kbase + swz(r, c), b0, b1, b2, b3);
        mma_bf16_16816(s[2 * jp], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
        mma_bf16_16816(s[2 * jp + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
      }
    }
    if (p.n_keys < ATT_BN) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kidx = 8 * j + 2 * tq;
        if (kidx >= p.n_keys) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (kidx + 1 >= p.n_keys) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    }
    float l[2] = {0.f, 0.f};
    uint32_t pf[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] - mx[0]), p1 = exp2f(s[j][1] - mx[0]);
      const float p2 = exp2f(s[j][2] - mx[1]), p3 = exp2f(s[j][3] - mx[1]);
      l[0] += p0 + p1;
      l[1] += p2 + p3;
      const int kk = j >> 1;
      if ((j & 1) == 0) { pf[kk][0] = pack_bf16x2(p0, p1); pf[kk][1] = pack_bf16x2(p2, p3); }
      else              { pf[kk][2] = pack_bf16x2(p0, p1); pf[kk][3] = pack_bf16x2(p2, p3); }
    }
    float o_acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o_acc[j][0] = o_acc[j][1] = o_acc[j][2] = o_acc[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t b0, b1, b2, b3;
        const int r = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = 2 * jp + (lane >> 4);
        ldsm_x4_trans(vbase + swz(r, c), b0, b1, b2, b3);
        mma_bf16_16816(o_acc[2 * jp], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b0, b1);
        mma_bf16_16816(o_acc[2 * jp + 1], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b2, b3);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      l[i] += __shfl_xor_sync(0xffffffffu, l[i], 1);
      l[i] += __shfl_xor_sync(0xffffffffu, l[i], 2);
    }
    const float inv0 = 1.f / l[0], inv1 = 1.f / l[1];
    // ---- O -> this warp's 16 rows of the query buffer (only this warp ever read them) -> coalesced 16-byte stores
    __syncwarp();
    const int rl = warp * 16 + g;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<uint32_t*>(sQb + swz(rl, j) + 4 * tq) = pack_bf16x2(o_acc[j][0] * inv0, o_acc[j][1] * inv0);
      *reinterpret_cast<uint32_t*>(sQb + swz(rl + 8, j) + 4 * tq) = pack_bf16x2(o_acc[j][2] * inv1, o_acc[j][3] * inv1);
    }
    __syncwarp();
    const int row0 = t * ATT_BM;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = warp * 16 + it * 4 + (lane >> 3), c = lane & 7;
      if (row0 + r < p.rows)
        *reinterpret_cast<uint4*>(ob + (long long)(row0 + r) * p.q_rs + c * 8) = *reinterpret_cast<const uint4*>(sQb + swz(r, c));
    }
  }
  cp_async_wait<0>();
}

}  // namespace

int b200_attention_tc(const void* q, void* o, int64_t q_bs, int64_t q_hs, int32_t q_rs, int32_t rows, const void* k, const void* v,
                      int64_t kv_bs, int64_t kv_hs, int32_t kv_rs, int32_t n_keys, int B, int n_heads, float max_logit, cudaStream_t st);

int b200_cross_attention_tc(const void* q, void* o, int64_t q_bs, int64_t q_hs, int32_t q_rs, int32_t rows, const void* k, const void* v,
                            int64_t kv_bs, int64_t kv_hs, int32_t kv_rs, int32_t n_keys, int B, int n_heads, cudaStream_t st);

extern "C" int b200_attention(const void* q, void* o, int64_t q_bs, int64_t q_hs, int32_t q_rs, int32_t rows, const void* k,
                              const void* v, int64_t kv_bs, int64_t kv_hs, int32_t kv_rs, int32_t n_keys, int B, int n_heads,
                              float max_logit, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(q && o && k && v, "attention: null pointer");
  B200_REQUIRE(rows > 0 && n_keys > 0 && B > 0 && n_heads > 0, "attention: bad shape rows=%d keys=%d B=%d heads=%d", rows, n_keys, B, n_heads);
  B200_REQUIRE((q_rs & 7) == 0 && (kv_rs & 7) == 0 && kv_rs >= 64 && (q_hs & 7) == 0 && (q_bs & 7) == 0 && (kv_bs & 7) == 0 && (kv_hs & 7) == 0,
               "attention: strides must be multiples of 8 elements (16 bytes)");
  B200_REQUIRE(n_heads <= 65535 && B <= 65535, "attention: grid too large");
  // A usable logit bound selects the tcgen05 kernel (no running max, no rescale); 2^(-2*bound) must stay far above
  // fp32 underflow.  Without one (<= 0) or with a huge one, the online-softmax mma.sync kernel below is used.
  // (single-tile problems -- cross-attention over 39 context keys -- stay on the lighter mma.sync kernel: one CTA of the
  //  tcgen05 kernel pays ~4 us of TMEM/TMA/barrier setup, measured 238 us vs 127 us at the 64x64 level)
  if (max_logit > 0.f && max_logit <= 40.f && n_keys > 256)
    return b200_attention_tc(q, o, q_bs, q_hs, q_rs, rows, k, v, kv_bs, kv_hs, kv_rs, n_keys, B, n_heads, max_logit, st);
  AttnParams p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.q_bs = q_bs; p.q_hs = q_hs; p.kv_bs = kv_bs; p.kv_hs = kv_hs;
  p.q_rs = q_rs; p.kv_rs = kv_rs; p.rows = rows; p.n_keys = n_keys;
  // <= 64 keys, heads adjacent in q / k / v (the CrossAttention layout), bounded logits: persistent tcgen05 kernel
  if (n_keys <= 64 && max_logit > 0.f && max_logit <= 40.f && n_heads >= 1 && n_heads <= 8 && q_hs == 64 && kv_hs == 64 && q_rs == n_heads * 64 &&
      kv_rs == n_heads * 64 && (long long)rows >= 128) {
    static const bool tc_on = [] { const char* e = getenv("B200_IMAGEN_XATTN_TC"); return e == nullptr || atoi(e) != 0; }();
    if (tc_on) return b200_cross_attention_tc(q, o, q_bs, q_hs, q_rs, rows, k, v, kv_bs, kv_hs, kv_rs, n_keys, B, n_heads, st);
  }
  if (n_keys <= ATT_BN) {
    static const bool on = [] { const char* e = getenv("B200_IMAGEN_XATTN_FEWKEYS"); return e == nullptr || atoi(e) != 0; }();
    if (on) {
      constexpr int xsmem = 2 * ATT_BN * 128 + XA_QBUF * ATT_BM * 128;
      B200_SMEM_OPT_IN(cross_attn_fewkeys_kernel, xsmem);
      // contiguous tile ranges per CTA; one wave of CTAs (2 resident per SM) when the problem allows
      const int ntiles = (rows + ATT_BM - 1) / ATT_BM;
      const long long pairs = (long long)B * n_heads;
      int splits = (int)(2ll * sm_count_attn() / pairs);
      if (splits < 1) splits = 1;
      if (splits > ntiles) splits = ntiles;
      const int tiles_per_cta = (ntiles + splits - 1) / splits;
      dim3 xgrid((ntiles + tiles_per_cta - 1) / tiles_per_cta, n_heads, B);
      B200_CUDA_OK(b200_launch(cross_attn_fewkeys_kernel, xgrid, dim3(ATT_THREADS), xsmem, st, p, tiles_per_cta));
      return B200_OK;
    }
  }
  constexpr int smem = ATT_BM * 128 + 4 * ATT_BN * 128;
  B200_SMEM_OPT_IN(flash_attn_kernel, smem);
  dim3 grid((rows + ATT_BM - 1) / ATT_BM, n_heads, B);
  B200_CUDA_OK(b200_launch(flash_attn_kernel, grid, dim3(ATT_THREADS), smem, st, p));
  return B200_OK;
}
