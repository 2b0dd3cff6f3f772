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
//   warps 2..5 (2..9 for tiles of >= 128 columns): epilogue: tcgen05.ld (thread = output row, 16 columns at a time)
//                 -> bias / SiLU / GELU / per-head L2 norm / residual / pixel-shuffle / NCHW fp32 stores.
// Checker path (impl 1): a plain SIMT fp32 implicit GEMM + the same epilogue code, used only by tests
// to isolate tensor-core/TMA descriptor bugs from epilogue/packing bugs.
//
// Reference arithmetic replaced: see include/b200_imagen.h (b200_conv_gemm).
#include "ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int EPI_STAGE_BYTES = 2048 + 512;   // per epilogue warp: 32 rows x 64 B staging tile + 32 x 16 B row metadata

struct SegDev {
  int8_t src, dh, dw, pad;
};

struct EpiDev {
  const float* bias;
  const __nv_bfloat16* residual;
  void* out;
  void* out2;
  const float* l2_scale;
  const float* norm1_g;
  const float* norm2_g;
  const float* film;
  void* out_norm;
  float out_scale;
  int act, ldr, out_mode, ldc, ldc2, split_col, rows_per_group, group_stride, row_offset, l2_cols, ps_C, dup_rows;
  int norm1, norm2, film_ld, rows_per_sample, ld_norm;
};

struct GemmParams {
  int B, H, W;
  int bw, bh, bb;
  int lbw, lbh;   // log2(bw), log2(bh)
  int tiles_w, tiles_h;
  int N, Npad;
  int nseg, total_chunks, ntiles_m;
  int ksplit;      // > 1: K is split over ksplit adjacent work items; each writes its raw fp32 partial tile to out + split * M * ldc
  int tma_store;   // staged epilogue hands each 32-row x 32-column group to a TMA store (plain bf16 row-major destinations)
  int debug;   // B200_IMAGEN_GEMM_DEBUG bit mask (bottleneck experiments only): 1 skip epilogue work, 2 skip MMA issue, 4 skip TMA loads, 8 no wait before restaging (WRONG results), 16 no proxy fence (WRONG results)
  int nchunks[B200_MAX_SRC];
  SegDev seg[B200_MAX_SEG];
  EpiDev epi;
};

// ------------------------------------------------------------------------------------------ epilogue

struct RowInfo {
  bool valid;
  int b, h, w;
  long long row;   // (b*H + h)*W + w
  long long orow;  // after the optional row remap
};

// position of accumulator row m inside the tile box (constant per thread): bw / bh are powers of two
struct RowInTile {
  int lw, lh, lb;
};
__device__ __forceinline__ RowInTile row_in_tile(const GemmParams& p, int m) {
  RowInTile r;
  r.lw = m & (p.bw - 1);
  r.lh = (m >> p.lbw) & (p.bh - 1);
  r.lb = m >> (p.lbw + p.lbh);
  return r;
}

__device__ __forceinline__ RowInfo tile_row(const GemmParams& p, int tile, const RowInTile& t) {
  const unsigned ut = (unsigned)tile;
  const unsigned tq = ut / (unsigned)p.tiles_w;
  const int wblk = (int)(ut - tq * (unsigned)p.tiles_w);
  const unsigned bblk = tq / (unsigned)p.tiles_h;
  const int hblk = (int)(tq - bblk * (unsigned)p.tiles_h);
  RowInfo r;
  r.w = wblk * p.bw + t.lw;
  r.h = hblk * p.bh + t.lh;
  r.b = (int)bblk * p.bb + t.lb;
  r.valid = (r.w < p.W) && (r.h < p.H) && (r.b < p.B);
  r.row = ((long long)r.b * p.H + r.h) * p.W + r.w;
  r.orow = r.row;
  if (p.epi.rows_per_group > 0)
    r.orow = (r.row / p.epi.rows_per_group) * (long long)p.epi.group_stride + p.epi.row_offset + (r.row % p.epi.rows_per_group);
  return r;
}

// bias + activation + scale on 16 consecutive columns starting at n (all branches are warp-uniform)
__device__ __forceinline__ void epi_math16(const EpiDev& e, int n, float* v) {
  if (e.bias != nullptr) {   // bias is padded to Npad by the caller: vector loads are always in bounds
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + n + j));
      v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
    }
  }
  if (e.act == B200_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
  } else if (e.act == B200_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = gelu_erf_f(v[j]);
  }
  if (e.out_scale != 1.f) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] *= e.out_scale;
  }
}

// residual + store of 16 consecutive columns [n, n+16) of one output row.
// SIMPLE = plain bf16 row-major output, N % 16 == 0, 16-byte aligned rows, no split / remap / dup: the common case gets
// its own tiny instruction stream (the epilogue is instruction-fetch sensitive, see epilogue_row).
template <bool SIMPLE>
__device__ __forceinline__ void epi_store16(const GemmParams& p, const RowInfo& ri, int n, float* v) {
  const EpiDev& e = p.epi;
  const int N = p.N;
  if (!ri.valid || n >= N) return;
  if (SIMPLE) {
    if (e.residual != nullptr) {
      const __nv_bfloat16* rp = e.residual + ri.row * (long long)e.ldr + n;
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(rp)), f);
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] += f[t];
      unpack8(__ldg(reinterpret_cast<const uint4*>(rp + 8)), f);
#pragma unroll
      for (int t = 0; t < 8; ++t) v[8 + t] += f[t];
    }
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.out) + ri.row * (long long)e.ldc + n;
    reinterpret_cast<uint4*>(dst)[0] = pack8(v);
    reinterpret_cast<uint4*>(dst)[1] = pack8(v + 8);
    return;
  }
  const bool full = n + 16 <= N;
  if (e.residual != nullptr) {
    const __nv_bfloat16* rp = e.residual + ri.row * (long long)e.ldr + n;
    if (full && (e.ldr & 7) == 0) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(rp)), f);
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] += f[t];
      unpack8(__ldg(reinterpret_cast<const uint4*>(rp + 8)), f);
#pragma unroll
      for (int t = 0; t < 8; ++t) v[8 + t] += f[t];
    } else {
#pragma unroll
      for (int t = 0; t < 16; ++t)
        if (n + t < N) v[t] += __bfloat162float(rp[t]);
    }
  }
  if (e.out_mode == B200_OUT_BF16) {
    __nv_bfloat16* base;
    int ld, c;
    if (e.split_col > 0 && n >= e.split_col) {
      base = reinterpret_cast<__nv_bfloat16*>(e.out2); ld = e.ldc2; c = n - e.split_col;
    } else {
      base = reinterpret_cast<__nv_bfloat16*>(e.out); ld = e.ldc; c = n;
    }
    const int reps = e.dup_rows > 0 ? 2 : 1;
    for (int rep = 0; rep < reps; ++rep) {
      __nv_bfloat16* dst = base + (ri.orow + (long long)rep * e.dup_rows) * ld + c;
      if (full && (ld & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        reinterpret_cast<uint4*>(dst)[0] = pack8(v);
        reinterpret_cast<uint4*>(dst)[1] = pack8(v + 8);
      } else {
#pragma unroll
        for (int t = 0; t < 16; ++t)
          if (n + t < N) dst[t] = __float2bfloat16(v[t]);
      }
    }
  } else if (e.out_mode == B200_OUT_PIXEL_SHUFFLE) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(e.out);
    const int H2 = 2 * p.H, W2 = 2 * p.W;
#pragma unroll
    for (int g8 = 0; g8 < 16; g8 += 8) {
      const int nn = n + g8;
      if (nn >= N) break;
      const int r = nn / e.ps_C, c = nn - r * e.ps_C;
      const long long prow = ((long long)ri.b * H2 + 2 * ri.h + (r >> 1)) * W2 + 2 * ri.w + (r & 1);
      __nv_bfloat16* dst = out + prow * e.ldc + c;
      if ((e.ps_C & 7) == 0 && (e.ldc & 7) == 0) {
        *reinterpret_cast<uint4*>(dst) = pack8(v + g8);
      } else {
        for (int t = 0; t < 8; ++t) {
          const int n2 = nn + t;
          if (n2 < N) {
            const int rr = n2 / e.ps_C, cc = n2 - rr * e.ps_C;
            const long long pr = ((long long)ri.b * H2 + 2 * ri.h + (rr >> 1)) * W2 + 2 * ri.w + (rr & 1);
            out[pr * e.ldc + cc] = __float2bfloat16(v[g8 + t]);
          }
        }
      }
    }
  } else if (e.out_mode == B200_OUT_F32_NCHW) {
    float* out = reinterpret_cast<float*>(e.out);
#pragma unroll
    for (int t = 0; t < 16; ++t)
      if (n + t < N) out[(((long long)ri.b * N + n + t) * p.H + ri.h) * p.W + ri.w] = v[t];
  } else {  // B200_OUT_F32
    float* out = reinterpret_cast<float*>(e.out);
    const int reps = e.dup_rows > 0 ? 2 : 1;
    for (int rep = 0; rep < reps; ++rep) {
      float* dst = out + (ri.orow + (long long)rep * e.dup_rows) * e.ldc + n;
      if (full && (e.ldc & 3) == 0) {
#pragma unroll
        for (int t = 0; t < 16; t += 4) *reinterpret_cast<float4*>(dst + t) = make_float4(v[t], v[t + 1], v[t + 2], v[t + 3]);
      } else {
#pragma unroll
        for (int t = 0; t < 16; ++t)
          if (n + t < N) dst[t] = v[t];
      }
    }
  }
}

// Epilogue of one accumulator row over columns [n0, n0 + BNW) in chunks of 16 columns.
// load_async(c, v) starts fetching 16 fp32 accumulator columns from tile column c; wait(v) completes it.
// Design notes (both learnt from ncu on B200):
//  * the loop body is deliberately small (one 16-column chunk) and NOT unrolled: only four warps run the epilogue, so it
//    is instruction-fetch bound as soon as the body outgrows the L0 instruction cache (stall_no_inst dominated both a
//    256-column and a 64-column unrolled version);
//  * the TMEM read of chunk c+1 (~230 cycles of tcgen05.ld latency) is issued before chunk c is processed.
template <int BNW, bool SIMPLE, class LoadAsync, class Wait>
__device__ __forceinline__ void epilogue_row(const GemmParams& p, const RowInfo& ri, int n0, LoadAsync load_async, Wait wait) {
  const EpiDev& e = p.epi;
  constexpr int GROUP = BNW >= 64 ? 64 : BNW;
  constexpr int NCH = GROUP / 16;
  float v[16], vn[16];
#pragma unroll 1
  for (int g0 = 0; g0 < BNW; g0 += GROUP) {
    if (n0 + g0 >= p.N) break;
    float inv = 1.f;
    const bool l2 = !SIMPLE && (GROUP == 64) && (n0 + g0 < e.l2_cols);
    if (l2) {   // F.normalize over the 64-column head: first pass = sum of squares of the activated values
      float ss = 0.f;
      load_async(g0, v);
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        wait(v);
        if (c + 1 < NCH) load_async(g0 + 16 * (c + 1), vn);
        epi_math16(e, n0 + g0 + 16 * c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) ss += v[j] * v[j];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = vn[j];
      }
      inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    }
    load_async(g0, v);
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      wait(v);
      if (c + 1 < NCH) load_async(g0 + 16 * (c + 1), vn);
      epi_math16(e, n0 + g0 + 16 * c, v);
      if (l2) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (e.l2_scale != nullptr) s4 = __ldg(reinterpret_cast<const float4*>(e.l2_scale + 16 * c + j));
          v[j] *= inv * s4.x; v[j + 1] *= inv * s4.y; v[j + 2] *= inv * s4.z; v[j + 3] *= inv * s4.w;
        }
      }
      epi_store16<SIMPLE>(p, ri, n0 + g0 + 16 * c, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = vn[j];
    }
  }
}

// ---- staged epilogue (product path for every bf16 output mode) --------------------------------------------------------
// The TMEM layout gives each thread one output ROW, so direct stores scatter every warp instruction over 32 rows (32
// half-used sectors; tools/gemm_bench.py measured the small-K linears 3-5x slower with the epilogue than without).  Each
// epilogue warp therefore owns a 32-row x 64-column (128 B) staging tile in shared memory, XOR-swizzled in 16 B chunks:
//   phase A (thread = row): tcgen05.ld 16 columns at a time -> bias/act/scale -> bf16 -> st.shared      (conflict free)
//   phase B (8 lanes = one row): ld.shared 16 B -> [L2-norm scale] [+ residual] -> st.global: every warp instruction writes
//            four complete 128 B lines.
// Per-row data (destination, validity, 1/norm, residual row) travel from the row's owner lane by shuffle.
// bf16 destinations only; a 64-column group never straddles a destination switch (split_col % 64 == 0, ps_C % 64 == 0).

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// where column n (a multiple of 64) of this row goes, and the row pitch of that destination
__device__ __forceinline__ __nv_bfloat16* bf16_group_dst(const GemmParams& p, const RowInfo& ri, int n, int& ld) {
  const EpiDev& e = p.epi;
  if (e.out_mode == B200_OUT_PIXEL_SHUFFLE) {
    const int r = n / e.ps_C, c = n - r * e.ps_C;
    const long long prow = ((long long)ri.b * (2 * p.H) + 2 * ri.h + (r >> 1)) * (2 * p.W) + 2 * ri.w + (r & 1);
    ld = e.ldc;
    return reinterpret_cast<__nv_bfloat16*>(e.out) + prow * e.ldc + c;
  }
  if (e.split_col > 0 && n >= e.split_col) {
    ld = e.ldc2;
    return reinterpret_cast<__nv_bfloat16*>(e.out2) + ri.orow * (long long)e.ldc2 + (n - e.split_col);
  }
  ld = e.ldc;
  return reinterpret_cast<__nv_bfloat16*>(e.out) + ri.orow * (long long)e.ldc + n;
}

// bias + activation + scale on 32 consecutive columns starting at n (branches are warp-uniform)
__device__ __forceinline__ void epi_math32(const EpiDev& e, int n, float* v) {
  if (e.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + n + j));
      v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
    }
  }
  if (e.act == B200_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
  } else if (e.act == B200_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
  }
  if (e.out_scale != 1.f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= e.out_scale;
  }
}

// phase A for 32 columns: the thread's row of the 32-row x 64-byte staging tile (16-byte chunks XOR-swizzled so that both the
// row-wise writes here and the 4-lanes-per-row reads of flush32 are bank-conflict free)
__device__ __forceinline__ void stage32(const float* v, uint32_t stage, int lane) {
  const uint32_t rowaddr = stage + (uint32_t)lane * 64u;
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) st_shared_v4(rowaddr + (uint32_t)((i ^ sw) << 4), pack8(v + 8 * i));
}

// phase B for 32 columns starting at ng: 4 lanes = one row (64 contiguous bytes), 8 rows per warp instruction
__device__ __forceinline__ void flush32(const GemmParams& p, const RowInfo& ri, int ng, uint32_t stage, uint32_t meta, int lane) {
  const EpiDev& e = p.epi;
  int ld;
  __nv_bfloat16* dst = bf16_group_dst(p, ri, ng, ld);
  const long long my_dst = ri.valid ? reinterpret_cast<long long>(dst) : 0ll;
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(meta + (uint32_t)lane * 16u), "r"((uint32_t)my_dst), "r"((uint32_t)(my_dst >> 32)),
               "r"((uint32_t)ri.row), "r"(0u)
               : "memory");
  const int rs = lane >> 2, ch = lane & 3;
  const int col = ng + ch * 8;
  const uint32_t rd = stage + (uint32_t)rs * 64u + (uint32_t)((ch ^ ((rs >> 1) & 3)) << 4);   // + 512 per 8 rows; (row >> 1) & 3 == (rs >> 1) & 3
  __syncwarp();
  uint4 u[4];
  uint32_t m0[4], m1[4], m2[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    u[r] = ld_shared_v4(rd + (uint32_t)r * 512u);
    asm volatile("{\n\t.reg .b32 pad;\n\tld.shared.v4.b32 {%0, %1, %2, pad}, [%3];\n\t}" : "=r"(m0[r]), "=r"(m1[r]), "=r"(m2[r]) : "r"(meta + (uint32_t)(8 * r + rs) * 16u) : "memory");
  }
  if (col < p.N) {
    if (e.residual != nullptr) {
      uint4 rr[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        rr[r] = make_uint4(0u, 0u, 0u, 0u);
        if ((m0[r] | m1[r]) != 0u) rr[r] = __ldg(reinterpret_cast<const uint4*>(e.residual + (long long)(int)m2[r] * e.ldr + col));
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float f[8], g[8];
        unpack8(u[r], f);
        unpack8(rr[r], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += g[j];
        u[r] = pack8(f);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const long long d = (long long)(((unsigned long long)m1[r] << 32) | m0[r]);
      if (d != 0) {
        __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(d) + ch * 8;
        *reinterpret_cast<uint4*>(q) = u[r];
        if (e.dup_rows > 0) *reinterpret_cast<uint4*>(q + (long long)e.dup_rows * ld) = u[r];
      }
    }
  }
  __syncwarp();
}

// Straight-line code on purpose: the epilogue runs on 8 warps (2 per scheduler), so it is bound by dependent-instruction
// latency, not issue slots (ncu on the K=128 linears: CPI 6 per warp with a rolled 16-column loop).  One tcgen05.ld.x32 per
// wait, the next one in flight while 32 columns are converted, and all row reads of a flush issued together.
// hand the staged 32 x 32 group to the TMA engine: the async proxy reads the tile after the fence; OOB rows / columns are clipped
__device__ __forceinline__ void flush32_tma(const CUtensorMap* mapO, const RowInfo& ri, int ng, uint32_t stage, int lane, int debug) {
  if (!(debug & 16)) fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {   // lane 0 owns the first row of the warp's 32: its (w, h, b) are the box origin
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(mapO), "r"(stage), "r"(ng), "r"(ri.w),
                 "r"(ri.h), "r"(ri.b)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}
// the staging tile may be rewritten once the previous store has finished READING it
__device__ __forceinline__ void tma_store_wait_read(int lane, int debug) {
  if (lane == 0 && !(debug & 8)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

template <int BNW, bool PIPE, class Load32, class Wait32>
__device__ __forceinline__ void epilogue_staged(const GemmParams& p, const CUtensorMap* mapO, const RowInfo& ri, int n0, uint32_t stage, uint32_t meta,
                                                int lane, Load32 load32, Wait32 wait32) {
  static_assert(BNW % 32 == 0, "staged epilogue works on 32-column groups");
  const EpiDev& e = p.epi;
  float va[32];
  if (e.l2_cols > 0) {
    // F.normalize over 64-column heads: pass 1 = sum of squares of the activated values, pass 2 = reload, scale, stage.
    // (the scale is applied in fp32 before the single bf16 rounding)
#pragma unroll 1
    for (int g0 = 0; g0 < BNW; g0 += 64) {
      const int ng = n0 + g0;
      if (ng >= p.N) break;
      const bool l2 = ng < e.l2_cols;
      float inv = 1.f;
      if (l2) {
        float ss = 0.f;
#pragma unroll 1
        for (int h = 0; h < 64 && h < BNW; h += 32) {
          load32(g0 + h, va);
          wait32(va);
          epi_math32(e, ng + h, va);
#pragma unroll
          for (int j = 0; j < 32; ++j) ss += va[j] * va[j];
        }
        inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
      }
#pragma unroll 1
      for (int h = 0; h < 64 && h < BNW; h += 32) {
        if (ng + h >= p.N) break;
        load32(g0 + h, va);
        wait32(va);
        epi_math32(e, ng + h, va);
        if (l2) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
            if (e.l2_scale != nullptr) s4 = __ldg(reinterpret_cast<const float4*>(e.l2_scale + h + j));
            va[j] *= inv * s4.x; va[j + 1] *= inv * s4.y; va[j + 2] *= inv * s4.z; va[j + 3] *= inv * s4.w;
          }
        }
        if (p.tma_store) {
          tma_store_wait_read(lane, p.debug);
          stage32(va, stage, lane);
          flush32_tma(mapO, ri, ng + h, stage, lane, p.debug);
        } else {
          stage32(va, stage, lane);
          flush32(p, ri, ng + h, stage, meta, lane);
        }
      }
    }
    return;
  }
  if constexpr (!PIPE) {
    // 16 epilogue warps (4 per scheduler): thread-level parallelism hides the tcgen05.ld latency, one register buffer is enough
#pragma unroll 1
    for (int g0 = 0; g0 < BNW; g0 += 32) {
      const int ng = n0 + g0;
      if (ng >= p.N) break;
      load32(g0, va);
      wait32(va);
      epi_math32(e, ng, va);
      if (p.tma_store) {
        tma_store_wait_read(lane, p.debug);
        stage32(va, stage, lane);
        flush32_tma(mapO, ri, ng, stage, lane, p.debug);
      } else {
        stage32(va, stage, lane);
        flush32(p, ri, ng, stage, meta, lane);
      }
    }
    return;
  }
  float vb[32];
  load32(0, va);
#pragma unroll 1
  for (int g0 = 0; g0 < BNW; g0 += 64) {
    const int ng = n0 + g0;
    if (ng >= p.N) break;
    wait32(va);
    if (g0 + 32 < BNW) load32(g0 + 32, vb);
    epi_math32(e, ng, va);
    if (p.tma_store) {
      tma_store_wait_read(lane, p.debug);
      stage32(va, stage, lane);
      flush32_tma(mapO, ri, ng, stage, lane, p.debug);
    } else {
      stage32(va, stage, lane);
      flush32(p, ri, ng, stage, meta, lane);
    }
    if (g0 + 32 < BNW && ng + 32 < p.N) {
      wait32(vb);
      if (g0 + 64 < BNW) load32(g0 + 64, va);
      epi_math32(e, ng + 32, vb);
      if (p.tma_store) {
        tma_store_wait_read(lane, p.debug);
        stage32(vb, stage, lane);
        flush32_tma(mapO, ri, ng + 32, stage, lane, p.debug);
      } else {
        stage32(vb, stage, lane);
        flush32(p, ri, ng + 32, stage, meta, lane);
      }
    } else if (g0 + 32 < BNW) {
      wait32(vb);   // drain the prefetch before leaving
    }
  }
  tmem_ld_wait();   // a prefetch may still be in flight when the column loop ends at N: never release the accumulator under it
}

__device__ __forceinline__ void tmem_wait_regs32(float* v) {
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) asm volatile("" : "+f"(v[i]));
}

// wait::ld, then pin the 16 destination registers behind the wait so the compiler cannot hoist their uses above it
__device__ __forceinline__ void tmem_wait_regs16(float* v) {
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) asm volatile("" : "+f"(v[i]));
}

// ---- norm-fusing epilogue (b200_epilogue.norm1 / norm2) -----------------------------------------------------------------
// One N tile spans all channels, so a thread (= accumulator row = one pixel / token) sees its whole row in TMEM, BNH columns per
// epilogue warp.  Row statistics are reduced across the HALVES warps that share a TMEM lane quarter through a small shared-memory
// exchange + a named barrier per quarter; the accumulator is simply re-read from TMEM for every pass (TMEM reads are cheap, the
// GEMMs this is used on are bound by the MMA pipe or by HBM, not by the epilogue):
//   [norm1] pass: sum -> mean, pass: centred sum of squares -> rstd          (two-pass LayerNorm, fp32)
//   [norm2] pass: w = norm1(v) + residual; sum(w), sum(w^2)                  (LayerNorm: var = E[w^2] - mean^2; RMS: |w|_2)
//   final pass  : w -> out (raw, optional), norm2(w) -> out_norm; both through the warp's staging tile -> coalesced stores.
// Exchange slots are indexed by (tile parity, round) so that a fast warp's next-tile write can never overtake a slow warp's read.

__device__ __forceinline__ void flush32_to(__nv_bfloat16* base, int ld, const GemmParams& p, const RowInfo& ri, int ng, uint32_t stage, uint32_t meta,
                                           int lane) {
  const long long my_dst = ri.valid ? reinterpret_cast<long long>(base + ri.row * (long long)ld + ng) : 0ll;
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(meta + (uint32_t)lane * 16u), "r"((uint32_t)my_dst), "r"((uint32_t)(my_dst >> 32)) : "memory");
  const int rs = lane >> 2, ch = lane & 3;
  const uint32_t rd = stage + (uint32_t)rs * 64u + (uint32_t)((ch ^ ((rs >> 1) & 3)) << 4);
  __syncwarp();
  uint4 u[4];
  uint32_t m0[4], m1[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    u[r] = ld_shared_v4(rd + (uint32_t)r * 512u);
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(m0[r]), "=r"(m1[r]) : "r"(meta + (uint32_t)(8 * r + rs) * 16u) : "memory");
  }
  if (ng + ch * 8 < p.N) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const long long d = (long long)(((unsigned long long)m1[r] << 32) | m0[r]);
      if (d != 0) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d) + ch * 8) = u[r];
    }
  }
  __syncwarp();
}

// One epilogue warp owns 64 columns of its 32 rows and keeps them in registers: ONE pass over TMEM (the accumulator stage is handed
// back to the MMA warp right after it), every value computed once (a first version re-read TMEM and recomputed bias / GELU /
// norm1 for each statistics pass: 2-3x the instructions, and with 8 epilogue warps the fused GEMMs became instruction bound --
// FF1+GELU+LN 102 us vs 45 + 25 us unfused).
// sum of n register values with four independent accumulators (a 64-deep dependent FADD chain costs ~256 cycles of latency,
// and the epilogue warps are latency bound)
template <int N_, class F>
__device__ __forceinline__ float sum4(F f) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int j = 0; j < N_; j += 4) { a0 += f(j); a1 += f(j + 1); a2 += f(j + 2); a3 += f(j + 3); }
  return (a0 + a1) + (a2 + a3);
}

template <int BNH, int HALVES, class Load32, class Wait32, class Release>
__device__ __forceinline__ void epilogue_norm(const GemmParams& p, const RowInfo& ri, int n0, uint32_t stage, uint32_t meta, float2* xch, int parity,
                                              int q, int half, int lane, Load32 load32, Wait32 wait32, Release release) {
  static_assert(BNH == 32 || BNH == 64, "one or two 32-column groups per epilogue warp, kept in registers");
  constexpr int NG = BNH / 32;
  const EpiDev& e = p.epi;
  const int r = q * 32 + lane;
  const float invN = 1.f / (float)p.N;
  auto exchange = [&](int round, float a, float b) -> float2 {
    if (HALVES == 1) return make_float2(a, b);
    float2* reg = xch + (size_t)((parity * 3 + round) * HALVES) * 128;
    reg[half * 128 + r] = make_float2(a, b);
    asm volatile("bar.sync %0, %1;" ::"r"(8 + q), "n"(32 * HALVES) : "memory");   // the HALVES warps of this lane quarter
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < HALVES; ++h) {      // same order in every warp: bitwise identical statistics
      const float2 t = reg[h * 128 + r];
      s.x += t.x; s.y += t.y;
    }
    return s;
  };
  // ---- accumulator -> registers (columns >= N stay zero and are excluded from every statistic)
  bool hv[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) hv[g] = n0 + 32 * g < p.N;
  float v[BNH];
#pragma unroll
  for (int j = 0; j < BNH; ++j) v[j] = 0.f;
#pragma unroll
  for (int g = 0; g < NG; ++g)
    if (hv[g]) { load32(32 * g, v + 32 * g); wait32(v + 32 * g); }
  release();                                // the MMA warp may overwrite this accumulator stage
#pragma unroll
  for (int g = 0; g < NG; ++g)
    if (hv[g]) epi_math32(e, n0 + 32 * g, v + 32 * g);
  // ---- norm1: two-pass LayerNorm on the fp32 values
  if (e.norm1) {
    const float mean = exchange(0, sum4<BNH>([&](int j) { return v[j]; }), 0.f).x * invN;
    float vs = 0.f;
#pragma unroll
    for (int g = 0; g < NG; ++g)
      if (hv[g]) vs += sum4<32>([&](int j) { const float d = v[32 * g + j] - mean; return d * d; });
    const float rstd = rsqrtf(exchange(1, vs, 0.f).x * invN + 1e-5f);
    const float nmr = -mean * rstd;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if (hv[g]) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(e.norm1_g + n0 + 32 * g + j));
          float* w = v + 32 * g + j;
          w[0] = fmaf(w[0], rstd, nmr) * g4.x; w[1] = fmaf(w[1], rstd, nmr) * g4.y;
          w[2] = fmaf(w[2], rstd, nmr) * g4.z; w[3] = fmaf(w[3], rstd, nmr) * g4.w;
        }
      }
    }
  }
  // ---- residual (the thread's own row: 64-byte pieces)
  if (e.residual != nullptr && ri.valid) {
    const __nv_bfloat16* rrow = e.residual + ri.row * (long long)e.ldr + n0;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if (hv[g]) {
        uint4 u[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) u[t] = *reinterpret_cast<const uint4*>(rrow + 32 * g + 8 * t);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float f[8];
          unpack8(u[t], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[32 * g + 8 * t + i] += f[i];
        }
      }
    }
  }
  if (e.out != nullptr) {
#pragma unroll
    for (int g = 0; g < NG; ++g)
      if (hv[g]) { stage32(v + 32 * g, stage, lane); flush32_to(reinterpret_cast<__nv_bfloat16*>(e.out), e.ldc, p, ri, n0 + 32 * g, stage, meta, lane); }
  }
  if (!e.norm2) return;
  // ---- norm2 on w (fp32): LayerNorm (var = E[w^2] - mean^2) or RMSNorm -> FiLM -> SiLU
  const float s = e.norm2 == 1 ? sum4<BNH>([&](int j) { return v[j]; }) : 0.f;
  const float ss = sum4<BNH>([&](int j) { return v[j] * v[j]; });
  const float2 t = exchange(2, s, ss);
  float m2 = 0.f, k2;
  if (e.norm2 == 1) {
    m2 = t.x * invN;
    k2 = rsqrtf(fmaxf(t.y * invN - m2 * m2, 0.f) + 1e-5f);
  } else {
    k2 = 1.f / fmaxf(sqrtf(t.y), 1e-12f);
  }
  const float nm2 = -m2 * k2;
  const float* film = nullptr;
  if (e.norm2 == 2 && e.film != nullptr && ri.valid) film = e.film + (long long)((unsigned)ri.row / (unsigned)e.rows_per_sample) * e.film_ld;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if (hv[g]) {
      float* w = v + 32 * g;
      const int n = n0 + 32 * g;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(e.norm2_g + n + j));
        w[j] = fmaf(w[j], k2, nm2) * g4.x; w[j + 1] = fmaf(w[j + 1], k2, nm2) * g4.y;
        w[j + 2] = fmaf(w[j + 2], k2, nm2) * g4.z; w[j + 3] = fmaf(w[j + 3], k2, nm2) * g4.w;
      }
      if (e.norm2 == 2) {
        if (film != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(film + n + j));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(film + p.N + n + j));
            w[j] = fmaf(w[j], s4.x, w[j]) + b4.x; w[j + 1] = fmaf(w[j + 1], s4.y, w[j + 1]) + b4.y;
            w[j + 2] = fmaf(w[j + 2], s4.z, w[j + 2]) + b4.z; w[j + 3] = fmaf(w[j + 3], s4.w, w[j + 3]) + b4.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) w[j] = silu_f(w[j]);
      }
      stage32(w, stage, lane);
      flush32_to(reinterpret_cast<__nv_bfloat16*>(e.out_norm), e.ld_norm, p, ri, n, stage, meta, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------ tcgen05 kernel
// Persistent: one CTA per SM loops over output tiles; the smem operand ring (TMA -> MMA) runs straight across
// tile boundaries and the TMEM accumulator is double buffered, so the epilogue of tile i overlaps the MMAs of
// tile i+1.

template <int BN, int STAGES, bool STAGED, int NEPI, bool NORM>
__global__ void __launch_bounds__(64 + 32 * NEPI, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                    const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapO, const __grid_constant__ GemmParams p) {
  constexpr int B_TILE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  constexpr int ACC_COLS = BN < 32 ? 32 : BN;
  constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;   // two accumulator stages (power of two)

  pdl_trigger();   // the next grid may be scheduled; it waits (pdl_wait) for this one to complete before touching memory
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles_n = p.Npad / BN;
  const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
  const int total_tiles = p.ntiles_m * n_tiles_n * ksplit;   // work items: (row tile, column tile, K split), the split index fastest

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    if (p.nchunks[1] > 0) tma_prefetch_desc(&mapA1);
    if (p.nchunks[2] > 0) tma_prefetch_desc(&mapA2);
    if (p.nchunks[3] > 0) tma_prefetch_desc(&mapA3);
    tma_prefetch_desc(&mapB);
    if (STAGED && p.tma_store) tma_prefetch_desc(&mapO);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], NEPI);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform: no ELECT / R2UR.BROADCAST loop around every tcgen05 instruction
  pdl_wait();      // barriers, TMEM and descriptors are set up: from here on the previous grid's results are read

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int split = t % ksplit, tt = t / ksplit;
        const int kc0 = split * p.total_chunks / ksplit, kc1 = (split + 1) * p.total_chunks / ksplit;
        const int tile = tt / n_tiles_n, n0 = (tt % n_tiles_n) * BN;
        const int wblk = tile % p.tiles_w;
        const int hblk = (tile / p.tiles_w) % p.tiles_h;
        const int bblk = tile / (p.tiles_w * p.tiles_h);
        const int w0 = wblk * p.bw, h0 = hblk * p.bh, b0 = bblk * p.bb;
        int kc = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int src = p.seg[s].src;
          const CUtensorMap* mA = src == 0 ? &mapA0 : (src == 1 ? &mapA1 : (src == 2 ? &mapA2 : &mapA3));
          const int dh = p.seg[s].dh, dw = p.seg[s].dw;
          const int nch = p.nchunks[src];
          for (int cc = 0; cc < nch; ++cc, ++kc) {
            if (kc < kc0 || kc >= kc1) continue;   // another split's share of K
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            if (p.debug & 4) {
              mbar_arrive(&full_bar[stage]);
            } else {
              mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
              uint8_t* sA = smem + stage * STAGE_BYTES;
              tma_load_4d(sA, mA, &full_bar[stage], cc * BK, w0 + dw, h0 + dh, b0);
              tma_load_2d(sA + A_TILE_BYTES, &mapB, &full_bar[stage], kc * BK, n0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer (InstrDescriptor: c_format F32 [4,6)=1, a/b BF16 [7,10)=[10,13)=1,
      // K-major both, N>>3 at [17,23), M>>4 at [24,29))
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);   // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        const int split = t % ksplit;
        const int nkc = (split + 1) * p.total_chunks / ksplit - split * p.total_chunks / ksplit;
        for (int kc = 0; kc < nkc; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
          const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + A_TILE_BYTES);
          if (!(p.debug & 2)) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (>>4) start-address field
              umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ---------------- epilogue warps: TMEM lane quarter = warp % 4, NEPI / 4 warps per quarter, each owning BN / (NEPI / 4) columns.
    // 4 warps for narrow tiles, 8 for >= 128 columns, 16 for the small-K GEMMs whose time IS the epilogue: with 2 warps per
    // scheduler the epilogue runs at CPI ~6 per warp (dependent tcgen05.ld -> convert -> st.shared chains), more warps hide it.
    constexpr int HALVES = NEPI / 4;
    constexpr int BNH = BN / HALVES;
    static_assert(BNH >= 32 && BNH % 32 == 0, "each epilogue warp needs at least one 32-column group");
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const RowInTile rit = row_in_tile(p, q * 32 + lane);
    int acc = 0;
    uint32_t acc_phase = 0;
    int tile_parity = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, tile_parity ^= 1) {
      const int tt = (int)((unsigned)t / (unsigned)ksplit), split = t - tt * ksplit;
      const int tile = (int)((unsigned)tt / (unsigned)n_tiles_n), n0 = (tt - tile * n_tiles_n) * BN;
      RowInfo ri = tile_row(p, tile, rit);
      if (p.ksplit > 1) ri.orow += (long long)split * p.B * p.H * p.W;   // partial tiles: [split][row][col] fp32
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + half * BNH);
      if (!(p.debug & 1)) {
        if constexpr (NORM) {
          static_assert(STAGED && (BNH == 64 || BNH == 32), "the norm epilogue keeps its columns in registers and stores through the staging tiles");
          const uint32_t stg = smem_u32(smem + STAGES * STAGE_BYTES + 1024 + (warp - 2) * EPI_STAGE_BYTES);
          float2* xch = reinterpret_cast<float2*>(smem + STAGES * STAGE_BYTES + 1024 + NEPI * EPI_STAGE_BYTES);
          epilogue_norm<BNH, HALVES>(p, ri, n0 + half * BNH, stg, stg + 2048u, xch, tile_parity, q, half, lane,
                                [&](int c, float* v) { tmem_ld32_nowait(taddr + (uint32_t)c, reinterpret_cast<uint32_t*>(v)); },
                                [&](float* v) { tmem_wait_regs32(v); },
                                [&]() {
                                  tc_fence_before();
                                  __syncwarp();
                                  if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                                });
        } else if constexpr (STAGED) {
          const uint32_t stg = smem_u32(smem + STAGES * STAGE_BYTES + 1024 + (warp - 2) * EPI_STAGE_BYTES);
          epilogue_staged<BNH, (NEPI <= 8)>(p, &mapO, ri, n0 + half * BNH, stg, stg + 2048u, lane,
                               [&](int c, float* v) { tmem_ld32_nowait(taddr + (uint32_t)c, reinterpret_cast<uint32_t*>(v)); },
                               [&](float* v) { tmem_wait_regs32(v); });
        } else {
          epilogue_row<BNH, false>(p, ri, n0 + half * BNH, [&](int c, float* v) { tmem_ld16_nowait(taddr + (uint32_t)c, reinterpret_cast<uint32_t*>(v)); },
                                   [&](float* v) { tmem_wait_regs16(v); });
        }
      }
      if (!(NORM && !(p.debug & 1))) {      // (the norm epilogue released the accumulator itself, right after reading it)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  if (STAGED && warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // TMA stores read this CTA's shared memory
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ tcgen05 kernel, CTA pairs
// Same pipeline with cta_group::2: the two CTAs of a 2-wide cluster (one TPC) compute one 256 x BN tile.  Each CTA loads
// the A tile of ITS 128 output rows and HALF of the B tile (BN/2 weight rows); the rank-0 CTA issues M=256 MMAs that read
// both shared memories and write each CTA's 128 accumulator lanes.  Per 128 x BN of output a CTA therefore pulls
// 16 KB + BN*64 B per K chunk instead of 16 KB + BN*128 B -- the conv GEMMs are bound by L2->SM operand traffic
// (profiles/r01_ncu_conv_gemm_summary.txt), so this is where the time goes.
//   full[s]   (rank 0 only): 1 arrival (rank 0's expect_tx of BOTH CTAs' bytes) + the TMA bytes of both CTAs
//   empty[s]  (both CTAs)  : multicast tcgen05.commit -> each producer waits locally
//   tmem_full (both CTAs)  : multicast tcgen05.commit -> each CTA's epilogue warps wait locally
//   tmem_empty (rank 0)    : 8 epilogue warps x 2 CTAs arrive (remote arrive from rank 1)

template <int BN, int STAGES, bool STAGED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 256, 1)
conv_gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                     const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                     const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapO, const __grid_constant__ GemmParams p) {
  constexpr int BNH = BN / 2;                       // B rows held by one CTA == columns owned by one epilogue half
  constexpr int B_TILE_BYTES = BNH * BK * 2;
  constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  constexpr int TMEM_COLS = 2 * BN;                 // two accumulator stages

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_tiles_n = p.Npad / BN;
  const int npairs_m = (p.ntiles_m + 1) / 2;
  const int total_tiles = npairs_m * n_tiles_n;
  const int pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    if (p.nchunks[1] > 0) tma_prefetch_desc(&mapA1);
    if (p.nchunks[2] > 0) tma_prefetch_desc(&mapA2);
    if (p.nchunks[3] > 0) tma_prefetch_desc(&mapA3);
    tma_prefetch_desc(&mapB);
    if (STAGED && p.tma_store) tma_prefetch_desc(&mapO);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 16);   // 8 epilogue warps of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, (uint32_t)TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();                      // barriers of BOTH CTAs are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform: no ELECT / R2UR.BROADCAST loop around every tcgen05 instruction

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer (both CTAs)
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair0; t < total_tiles; t += pair_stride) {
        const int tile = 2 * (t / n_tiles_n) + (int)rank, n0 = (t % n_tiles_n) * BN + (int)rank * BNH;
        const int wblk = tile % p.tiles_w;
        const int hblk = (tile / p.tiles_w) % p.tiles_h;
        const int bblk = tile / (p.tiles_w * p.tiles_h);   // tile == ntiles_m (odd tile count): b0 >= B, TMA zero-fills
        const int w0 = wblk * p.bw, h0 = hblk * p.bh, b0 = bblk * p.bb;
        int kc = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int src = p.seg[s].src;
          const CUtensorMap* mA = src == 0 ? &mapA0 : (src == 1 ? &mapA1 : (src == 2 ? &mapA2 : &mapA3));
          const int dh = p.seg[s].dh, dw = p.seg[s].dw;
          const int nch = p.nchunks[src];
          for (int cc = 0; cc < nch; ++cc, ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            if (p.debug & 4) {
              if (rank == 0) mbar_arrive(&full_bar[stage]);
            } else {
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
              const uint32_t full0 = mapa_u32(smem_u32(&full_bar[stage]), 0);
              uint8_t* sA = smem + stage * STAGE_BYTES;
              tma_load_4d_2sm(sA, mA, full0, cc * BK, w0 + dw, h0 + dh, b0);
              tma_load_2d_2sm(sA + A_TILE_BYTES, &mapB, full0, kc * BK, n0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      // ---------------- MMA issuer (rank 0 only): M = 256 over the pair, N = BN
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = pair0; t < total_tiles; t += pair_stride) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kc = 0; kc < p.total_chunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
          const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + A_TILE_BYTES);
          if (!(p.debug & 2)) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage], 3);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(&tmem_full_bar[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ---------------- epilogue warps (both CTAs): lane quarter = warp % 4, column half = (warp - 2) / 4
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const RowInTile rit = row_in_tile(p, q * 32 + lane);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = pair0; t < total_tiles; t += pair_stride) {
      const int tq = (int)((unsigned)t / (unsigned)n_tiles_n);
      const int tile = 2 * tq + (int)rank, n0 = (t - tq * n_tiles_n) * BN;
      const RowInfo ri = tile_row(p, tile, rit);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * BNH);
      if (!(p.debug & 1)) {
        if constexpr (STAGED) {
          const uint32_t stg = smem_u32(smem + STAGES * STAGE_BYTES + 1024 + (warp - 2) * EPI_STAGE_BYTES);
          epilogue_staged<BNH, true>(p, &mapO, ri, n0 + half * BNH, stg, stg + 2048u, lane,
                               [&](int c, float* v) { tmem_ld32_nowait(taddr + (uint32_t)c, reinterpret_cast<uint32_t*>(v)); },
                               [&](float* v) { tmem_wait_regs32(v); });
        } else {
          epilogue_row<BNH, false>(p, ri, n0 + half * BNH, [&](int c, float* v) { tmem_ld16_nowait(taddr + (uint32_t)c, reinterpret_cast<uint32_t*>(v)); },
                                   [&](float* v) { tmem_wait_regs16(v); });
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  if (STAGED && warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // TMA stores read this CTA's shared memory
  tc_fence_before();
  cluster_sync_all();                      // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, (uint32_t)TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ tcgen05 kernel, K split across a cluster
// The coarsest U-Net levels (8 x 8 pixels x 32 rows = 2048 output rows, 1024 channels, K = 9216 .. 13824) give only 64 tiles of 128 x 256 on
// 148 SMs; with 128 x 128 tiles (128 CTAs) they run at the N = 128 instruction rate (~700 TFLOP/s).  Splitting K through a global-memory
// workspace lost to the extra launch and the 16 MB partial round trip (profiles/r01_gemm_splitk_ab.txt).  Here the K range of ONE 128 x 256
// output tile is split across the 2..4 CTAs of a thread-block cluster: every CTA runs the usual TMA -> tcgen05 pipeline over its share of
// the K chunks, the non-zero ranks park their fp32 accumulator in their own shared memory (over the drained operand ring), and after one
// cluster barrier rank 0 adds the partial tiles through distributed shared memory inside its normal epilogue.  One launch, no HBM traffic,
// partials are added in rank order (deterministic).
constexpr int KS_BN = 256;
constexpr int KS_STAGES = 4;
constexpr int KS_STAGE_BYTES = A_TILE_BYTES + KS_BN * BK * 2;
constexpr int KS_PITCH = KS_BN + 4;                      // floats per parked accumulator row (bank-conflict-free float4 rows)

__global__ void __launch_bounds__(64 + 256, 1)
conv_gemm_tcS_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
                     const __grid_constant__ CUtensorMap mapA3, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapO,
                     const __grid_constant__ GemmParams p) {
  pdl_trigger();
  constexpr int HALVES = 2, BNH = KS_BN / HALVES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + KS_STAGES * KS_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + KS_STAGES;
  uint64_t* tmem_full_bar = empty_bar + KS_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* parked = reinterpret_cast<float*>(smem);        // [128][KS_PITCH] fp32, overlays the operand ring once the MMAs have completed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int nsplit = p.ksplit;
  const int n_tiles_n = p.Npad / KS_BN;
  const int tt = (int)blockIdx.x / nsplit;
  const int tile = tt / n_tiles_n, n0 = (tt % n_tiles_n) * KS_BN;
  const int kc0 = (int)rank * p.total_chunks / nsplit, kc1 = ((int)rank + 1) * p.total_chunks / nsplit;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    if (p.nchunks[1] > 0) tma_prefetch_desc(&mapA1);
    if (p.nchunks[2] > 0) tma_prefetch_desc(&mapA2);
    if (p.nchunks[3] > 0) tma_prefetch_desc(&mapA3);
    tma_prefetch_desc(&mapB);
    if (p.tma_store) tma_prefetch_desc(&mapO);
    for (int s = 0; s < KS_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)KS_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform: no ELECT / R2UR.BROADCAST loop around every tcgen05 instruction
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer: this rank's share of the K chunks
      const int wblk = tile % p.tiles_w;
      const int hblk = (tile / p.tiles_w) % p.tiles_h;
      const int bblk = tile / (p.tiles_w * p.tiles_h);
      const int w0 = wblk * p.bw, h0 = hblk * p.bh, b0 = bblk * p.bb;
      int stage = 0, kc = 0;
      uint32_t phase = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const int src = p.seg[s].src;
        const CUtensorMap* mA = src == 0 ? &mapA0 : (src == 1 ? &mapA1 : (src == 2 ? &mapA2 : &mapA3));
        const int dh = p.seg[s].dh, dw = p.seg[s].dw;
        const int nch = p.nchunks[src];
        for (int cc = 0; cc < nch; ++cc, ++kc) {
          if (kc < kc0 || kc >= kc1) continue;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_expect_tx(&full_bar[stage], KS_STAGE_BYTES);
          uint8_t* sA = smem + stage * KS_STAGE_BYTES;
          tma_load_4d(sA, mA, &full_bar[stage], cc * BK, w0 + dw, h0 + dh, b0);
          tma_load_2d(sA + A_TILE_BYTES, &mapB, &full_bar[stage], kc * BK, n0);
          if (++stage == KS_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KS_BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < kc1 - kc0; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * KS_STAGE_BYTES);
        const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
        const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + A_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == KS_STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tmem_full_bar);
    }
  } else if (rank != 0) {
    // ---------------- non-zero ranks: accumulator -> own shared memory (row = thread, KS_PITCH floats per row)
    const int q = warp & 3, half = (warp - 2) >> 2;
    mbar_wait(tmem_full_bar, 0);          // every MMA of this CTA has completed: the operand ring is free
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * BNH);
    float* row = parked + (size_t)(q * 32 + lane) * KS_PITCH + half * BNH;
#pragma unroll 1
    for (int c = 0; c < BNH; c += 32) {
      float v[32];
      tmem_ld32_nowait(taddr + (uint32_t)c, reinterpret_cast<uint32_t*>(v));
      tmem_wait_regs32(v);
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(row + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    tc_fence_before();
  }
  cluster_sync_all();                                     // #1: the partial tiles are parked and visible cluster-wide
  if (warp >= 2 && rank == 0) {
    // ---------------- rank 0: the normal staged epilogue over (own accumulator + parked partials of ranks 1..nsplit-1, in rank order)
    const int q = warp & 3, half = (warp - 2) >> 2;
    const RowInfo ri = tile_row(p, tile, row_in_tile(p, q * 32 + lane));
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * BNH);
    const uint32_t my_row = smem_u32(parked + (size_t)(q * 32 + lane) * KS_PITCH + half * BNH);
    const uint32_t stg = smem_u32(smem + KS_STAGES * KS_STAGE_BYTES + 1024 + (warp - 2) * EPI_STAGE_BYTES);
    int cur_c = 0;
    epilogue_staged<BNH, false>(p, &mapO, ri, n0 + half * BNH, stg, stg + 2048u, lane,
                                [&](int c, float* v) { cur_c = c; tmem_ld32_nowait(taddr + (uint32_t)c, reinterpret_cast<uint32_t*>(v)); },
                                [&](float* v) {
                                  tmem_wait_regs32(v);
                                  for (int rr = 1; rr < nsplit; ++rr) {
                                    const uint32_t ra = mapa_u32(my_row + (uint32_t)cur_c * 4u, (uint32_t)rr);
#pragma unroll
                                    for (int j = 0; j < 32; j += 4) {
                                      float4 t;
                                      asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(ra + (uint32_t)j * 4u) : "memory");
                                      v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
                                    }
                                  }
                                });
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tc_fence_before();
  }
  cluster_sync_all();                                     // #2: rank 0 has read every partial tile; nobody exits before that
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)KS_BN);
  }
}

// ------------------------------------------------------------------------------------------ tcgen05 kernel, transposed (<= 128 output channels)
// OPT-IN (B200_IMAGEN_GEMM_T=1).  Written when the 128 x 128-tile kernel above measured 878 TFLOP/s MMA-only against 1398 for 256-wide tiles
// (profiles/r01_gemm_bottleneck.txt) -- attributed then to operand fetch, in fact the per-instruction ELECT loops of its issuer thread (see
// elect_one in ptx.cuh): with those gone the row-major kernel reaches 1140 TFLOP/s MMA-only and wins (profiles/r02_gemm_after_elect.txt).
// Per 128 x 128 tile the row-major CTA pulls 32 KB of operands through L2 per 2.1 MFLOP.  This kernel computes
// the TRANSPOSED product  D^T[c][pixel] = sum_k W[c][k] * X[pixel][k]:  the weights are the M = 128 operand, a tile of 256 output pixels is the
// N = 256 operand (same shared-memory layouts, roles swapped), i.e. the instruction shape and operand traffic of the 256-wide tiles.
// The accumulator then holds channels in TMEM lanes and pixels in columns, so the epilogue transposes through shared memory:
//   phase A (thread = channel): tcgen05.ld 16 pixels -> + bias, activation -> fp32 staging tile [16 pixels][128 channels] (conflict-free pitch)
//   phase B (8 threads = one pixel row, 16 channels each): + residual -> raw bf16 row store and / or row statistics (shuffles over the 8
//            lanes) -> LayerNorm | RMSNorm * FiLM -> SiLU -> bf16 row store: every store instruction writes whole 256-byte pixel rows.
// Two groups of four epilogue warps (one per TMEM lane quarter) each own 128 of the tile's 256 pixels.
constexpr int T_BP = 256;                               // pixels per tile (MMA N)
constexpr int T_W_BYTES = 128 * BK * 2;                 // weight tile (MMA A operand): 16 KB
constexpr int T_X_BYTES = T_BP * BK * 2;                // pixel tile (MMA B operand): 32 KB
constexpr int T_STAGE_BYTES = T_W_BYTES + T_X_BYTES;
constexpr int T_STG_PITCH = 144;                        // floats per staged pixel row: 128 channels, +4 per 32-channel block, +12 pad (bank-conflict free)
constexpr int T_STG_BYTES = 16 * T_STG_PITCH * 4;       // per group: 16 pixel rows

__device__ __forceinline__ int t_stg_col(int c) { return c + 4 * (c >> 5); }

template <int STAGES>
__global__ void __launch_bounds__(64 + 256, 1)
conv_gemm_tcT_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
                     const __grid_constant__ CUtensorMap mapA3, const __grid_constant__ CUtensorMap mapW, const __grid_constant__ GemmParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * T_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* stg_all = reinterpret_cast<float*>(smem + STAGES * T_STAGE_BYTES + 1024);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.ntiles_m;               // 256-pixel tiles; one channel tile (Npad == 128)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    if (p.nchunks[1] > 0) tma_prefetch_desc(&mapA1);
    if (p.nchunks[2] > 0) tma_prefetch_desc(&mapA2);
    if (p.nchunks[3] > 0) tma_prefetch_desc(&mapA3);
    tma_prefetch_desc(&mapW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 8);             // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);        // two accumulator stages of 256 pixel columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform: no ELECT / R2UR.BROADCAST loop around every tcgen05 instruction
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer: weight tile {64 k, 128 channels} + pixel tile {64 ch, bw, bh, bb} shifted by the tap
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int wblk = tile % p.tiles_w;
        const int hblk = (tile / p.tiles_w) % p.tiles_h;
        const int bblk = tile / (p.tiles_w * p.tiles_h);
        const int w0 = wblk * p.bw, h0 = hblk * p.bh, b0 = bblk * p.bb;
        int kc = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int src = p.seg[s].src;
          const CUtensorMap* mA = src == 0 ? &mapA0 : (src == 1 ? &mapA1 : (src == 2 ? &mapA2 : &mapA3));
          const int dh = p.seg[s].dh, dw = p.seg[s].dw;
          const int nch = p.nchunks[src];
          for (int cc = 0; cc < nch; ++cc, ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_expect_tx(&full_bar[stage], T_STAGE_BYTES);
            uint8_t* sW = smem + stage * T_STAGE_BYTES;
            tma_load_2d(sW, &mapW, &full_bar[stage], kc * BK, 0);
            tma_load_4d(sW + T_W_BYTES, mA, &full_bar[stage], cc * BK, w0 + dw, h0 + dh, b0);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer: M = 128 (channels), N = 256 (pixels), both operands K-major
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(T_BP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * T_BP);
        for (int kc = 0; kc < p.total_chunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t w_addr = smem_u32(smem + stage * T_STAGE_BYTES);
          const uint64_t adesc = make_sw128_kmajor_desc(w_addr);
          const uint64_t bdesc = make_sw128_kmajor_desc(w_addr + T_W_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ---------------- epilogue: q = TMEM lane quarter (channels 32q..32q+31), grp = pixel half of the tile
    const EpiDev& e = p.epi;
    const int q = warp & 3, grp = (warp - 2) >> 2;
    const int c_lane = q * 32 + lane;                  // phase A: this thread's channel
    const int t = q * 32 + lane;                       // phase B: thread index inside the group
    const int prow = t >> 3, cs = t & 7;               //          pixel row of the 16-row chunk, 16-channel slice
    const int c0 = cs * 16;
    const bool cvalid = c0 < p.N;
    float* stg = stg_all + grp * (T_STG_BYTES / 4);
    const float bias = (e.bias != nullptr) ? __ldg(e.bias + c_lane) : 0.f;   // padded to Npad by the caller
    const float invN = 1.f / (float)p.N;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * T_BP + grp * 128);
#pragma unroll 1
      for (int ch = 0; ch < 8; ++ch) {
        // ---- phase A: 16 pixels of this thread's channel -> staging (transposed)
        float v[16];
        tmem_ld16(taddr + (uint32_t)(16 * ch), v);
        if (ch == 7) {                                 // the accumulator stage is in registers / shared memory: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        const int col = t_stg_col(c_lane);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float y = v[i] + bias;
          if (e.act == B200_ACT_SILU) y = silu_f(y);
          else if (e.act == B200_ACT_GELU) y = gelu_erf_f(y);
          stg[i * T_STG_PITCH + col] = y * e.out_scale;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(8 + grp) : "memory");
        // ---- phase B: one pixel row x 16 channels per thread
        const RowInfo ri = tile_row(p, tile, row_in_tile(p, grp * 128 + ch * 16 + prow));
        float f[16];
        {
          const float4* src = reinterpret_cast<const float4*>(stg + prow * T_STG_PITCH + t_stg_col(c0));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 x4 = src[k];
            f[4 * k] = x4.x; f[4 * k + 1] = x4.y; f[4 * k + 2] = x4.z; f[4 * k + 3] = x4.w;
          }
        }
        const bool live = ri.valid && cvalid;
        if (live && e.residual != nullptr) {
          const uint4* rp = reinterpret_cast<const uint4*>(e.residual + ri.row * (long long)e.ldr + c0);
          float g[8];
          unpack8(rp[0], g);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] += g[k];
          unpack8(rp[1], g);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[8 + k] += g[k];
        }
        if (live && e.out != nullptr) {
          uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out) + ri.row * (long long)e.ldc + c0);
          op[0] = pack8(f);
          op[1] = pack8(f + 8);
        }
        if (e.norm2) {
          float s = 0.f, ss = 0.f;
          if (cvalid) {
#pragma unroll
            for (int k = 0; k < 16; ++k) { s += f[k]; ss += f[k] * f[k]; }
          }
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {            // the 8 threads of a pixel row are 8 consecutive lanes
            s += __shfl_xor_sync(0xffffffffu, s, o);
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
          }
          float m2 = 0.f, k2;
          if (e.norm2 == 1) {
            m2 = s * invN;
            k2 = rsqrtf(fmaxf(ss * invN - m2 * m2, 0.f) + 1e-5f);
          } else {
            k2 = 1.f / fmaxf(sqrtf(ss), 1e-12f);
          }
          if (live) {
            const float nm2 = -m2 * k2;
            const float* film = (e.norm2 == 2 && e.film != nullptr) ? e.film + (long long)((unsigned)ri.row / (unsigned)e.rows_per_sample) * e.film_ld : nullptr;
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(e.norm2_g + c0 + k));
              f[k] = fmaf(f[k], k2, nm2) * g4.x; f[k + 1] = fmaf(f[k + 1], k2, nm2) * g4.y;
              f[k + 2] = fmaf(f[k + 2], k2, nm2) * g4.z; f[k + 3] = fmaf(f[k + 3], k2, nm2) * g4.w;
              if (film != nullptr) {
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(film + c0 + k));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(film + p.N + c0 + k));
                f[k] = fmaf(f[k], s4.x, f[k]) + b4.x; f[k + 1] = fmaf(f[k + 1], s4.y, f[k + 1]) + b4.y;
                f[k + 2] = fmaf(f[k + 2], s4.z, f[k + 2]) + b4.z; f[k + 3] = fmaf(f[k + 3], s4.w, f[k + 3]) + b4.w;
              }
            }
            if (e.norm2 == 2) {
#pragma unroll
              for (int k = 0; k < 16; ++k) f[k] = silu_f(f[k]);
            }
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out_norm) + ri.row * (long long)e.ld_norm + c0);
            op[0] = pack8(f);
            op[1] = pack8(f + 8);
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(8 + grp) : "memory");   // the staging tile may be overwritten
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
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

template <int BNW>
__global__ void conv_gemm_ref_epilogue_kernel(GemmParams p, const float* __restrict__ scratch) {
  const long long M = (long long)p.B * p.H * p.W;
  const int chunks = p.Npad / BNW;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * chunks) return;
  const int ch = (int)(idx % chunks);
  const long long m = idx / chunks;
  RowInfo ri;
  ri.valid = true;
  ri.w = (int)(m % p.W); ri.h = (int)((m / p.W) % p.H); ri.b = (int)(m / ((long long)p.W * p.H));
  ri.row = m;
  ri.orow = m;
  if (p.epi.rows_per_group > 0)
    ri.orow = (m / p.epi.rows_per_group) * (long long)p.epi.group_stride + p.epi.row_offset + (m % p.epi.rows_per_group);
  const float* src = scratch + m * p.Npad + ch * BNW;
  epilogue_row<BNW, false>(p, ri, ch * BNW, [&](int c, float* v) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = src[c + j];
  }, [](float*) {});
}

// split-K finish: out = epilogue(sum_s partial[s]) -- partials are added in split order, so the result does not depend on timing
template <int BNW>
__global__ void conv_gemm_splitk_epilogue_kernel(GemmParams p, const float* __restrict__ scratch, int ksplit) {
  const long long M = (long long)p.B * p.H * p.W;
  const int chunks = p.Npad / BNW;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * chunks) return;
  const int ch = (int)(idx % chunks);
  const long long m = idx / chunks;
  RowInfo ri;
  ri.valid = true;
  ri.w = (int)(m % p.W); ri.h = (int)((m / p.W) % p.H); ri.b = (int)(m / ((long long)p.W * p.H));
  ri.row = m;
  ri.orow = m;
  if (p.epi.rows_per_group > 0)
    ri.orow = (m / p.epi.rows_per_group) * (long long)p.epi.group_stride + p.epi.row_offset + (m % p.epi.rows_per_group);
  const float* src = scratch + m * p.Npad + ch * BNW;
  const long long stride = M * p.Npad;
  epilogue_row<BNW, false>(p, ri, ch * BNW, [&](int c, float* v) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      float4 a = *reinterpret_cast<const float4*>(src + c + j);
      for (int s = 1; s < ksplit; ++s) {
        const float4 b = *reinterpret_cast<const float4*>(src + s * stride + c + j);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      v[j] = a.x; v[j + 1] = a.y; v[j + 2] = a.z; v[j + 3] = a.w;
    }
  }, [](float*) {});
}

// B200_IMAGEN_GEMM_SPLITK=1: few-row, long-K GEMMs (the 8x8 levels: 64 tiles of 128 x 256 on 148 SMs) run as 2 K-splits of
// 128 x 256 tiles + a finishing kernel instead of 128 x 128 tiles (N=128 MMAs issue at 0.63 of the N=256 rate)
bool splitk_enabled() {
  static const bool on = [] {
    const char* e = getenv("B200_IMAGEN_GEMM_SPLITK");
    return e != nullptr && atoi(e) != 0;
  }();
  return on;
}
int splitk_factor(long long ntiles, int Npad, int total_chunks) {
  if (!splitk_enabled() || Npad % 256 != 0) return 1;
  const long long t256 = ntiles * (Npad / 256);
  return (t256 * 2 <= sm_count() && total_chunks >= 32) ? 2 : 1;   // always 2: the summation order must not depend on the batch size
}

// ------------------------------------------------------------------------------------------ host side

int next_pow2(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

template <int BN, int STAGES, bool SIMPLE, int NEPI, bool NORM = false>
int launch_tc2(const CUtensorMap* maps, const CUtensorMap& mapB, const CUtensorMap& mapO, const GemmParams& p, int ntiles, cudaStream_t st) {
  constexpr int smem = STAGES * (A_TILE_BYTES + BN * BK * 2) + 1024 /*align*/ + 256 /*barriers*/ + (SIMPLE ? 768 + NEPI * EPI_STAGE_BYTES : 0) /*staging*/ +
                       (NORM ? 6 * (NEPI / 4) * 128 * 8 : 0) /*row-statistics exchange*/;
  static_assert(smem <= 232448, "shared memory budget");
  B200_SMEM_OPT_IN((conv_gemm_tc_kernel<BN, STAGES, SIMPLE, NEPI, NORM>), smem);
  const long long total = (long long)ntiles * (p.Npad / BN) * (p.ksplit > 1 ? p.ksplit : 1);
  const int grid = (int)(total < sm_count() ? total : sm_count());   // persistent: one CTA per SM
  B200_CUDA_OK(b200_launch(conv_gemm_tc_kernel<BN, STAGES, SIMPLE, NEPI, NORM>, dim3(grid), dim3(64 + 32 * NEPI), smem, st, maps[0], maps[1], maps[2], maps[3], mapB, mapO, p));
  return B200_OK;
}

// every bf16 destination with 16-byte-aligned 8-column chunks takes the staged (coalescing) epilogue; the rest (fp32 outputs,
// odd channel counts, < 64 columns) keep the per-thread-row epilogue
bool gemm_is_simple(const GemmParams& p) {
  const EpiDev& e = p.epi;
  const auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!(e.out_mode == B200_OUT_BF16 || e.out_mode == B200_OUT_PIXEL_SHUFFLE)) return false;
  if (p.Npad < 64 || (p.N & 7) != 0 || (e.ldc & 7) != 0 || !al16(e.out)) return false;
  if (e.out_mode == B200_OUT_PIXEL_SHUFFLE && (e.ps_C % 64 != 0 || e.split_col != 0)) return false;
  if (e.split_col > 0 && ((e.ldc2 & 7) != 0 || !al16(e.out2))) return false;
  if (e.residual != nullptr && ((e.ldr & 7) != 0 || !al16(e.residual))) return false;
  if (e.dup_rows > 0 && e.split_col > 0) return false;
  if (e.l2_scale != nullptr && !al16(e.l2_scale)) return false;
  return (long long)p.B * p.H * p.W < (1ll << 31);
}

// B200_IMAGEN_GEMM_EPI16=0 keeps the small-K GEMMs on the 8-epilogue-warp kernels (A/B comparisons; measured on B200 with 16 warps:
// to_q + L2 norm 61 -> 49 us, FF1 + GELU 53 -> 45 us, plain 131072x128->512 41 -> 39 us, K = 512 linears 5 % slower -> K <= 256 only)
bool epi16_enabled() {
  static const bool on = [] {
    const char* e = getenv("B200_IMAGEN_GEMM_EPI16");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}

template <int BN, int STAGES>
int launch_tc(const CUtensorMap* maps, const CUtensorMap& mapB, const CUtensorMap& mapO, const GemmParams& p, int ntiles, cudaStream_t st) {
  constexpr int NEPI = BN >= 128 ? 8 : 4;
  if (!gemm_is_simple(p)) return launch_tc2<BN, STAGES, false, NEPI>(maps, mapB, mapO, p, ntiles, st);
  if constexpr (BN >= 128) {
    // K <= 256: a tile's MMAs take ~1-2k cycles, its epilogue ~8k with 8 warps -> 16 epilogue warps, shallower operand ring.
    // (the per-head L2 norm needs 64 columns per warp: 16 warps only at BN = 256)
    if (epi16_enabled() && p.total_chunks <= 4 && (p.epi.l2_cols == 0 || BN == 256))
      return launch_tc2<BN, (BN == 256 ? 3 : 4), true, 16>(maps, mapB, mapO, p, ntiles, st);
  }
  return launch_tc2<BN, STAGES, true, NEPI>(maps, mapB, mapO, p, ntiles, st);
}

template <int BN, int STAGES, bool SIMPLE>
int launch_pair2(const CUtensorMap* maps, const CUtensorMap& mapB, const CUtensorMap& mapO, const GemmParams& p, int ntiles, cudaStream_t st) {
  constexpr int smem = STAGES * (A_TILE_BYTES + (BN / 2) * BK * 2) + 1024 /*align*/ + 256 /*barriers*/ + (SIMPLE ? 768 + 8 * EPI_STAGE_BYTES : 0) /*staging*/;
  static_assert(smem <= 232448, "shared memory budget");
  B200_SMEM_OPT_IN((conv_gemm_tc2_kernel<BN, STAGES, SIMPLE>), smem);
  const long long total = (long long)((ntiles + 1) / 2) * (p.Npad / BN);     // 256 x BN tiles
  const int pairs = (int)(total < sm_count() / 2 ? total : sm_count() / 2);  // persistent: one CTA pair per TPC
  conv_gemm_tc2_kernel<BN, STAGES, SIMPLE><<<2 * pairs, 64 + 256, smem, st>>>(maps[0], maps[1], maps[2], maps[3], mapB, mapO, p);
  B200_LAUNCH_OK();
  return B200_OK;
}

template <int BN, int STAGES>
int launch_pair(const CUtensorMap* maps, const CUtensorMap& mapB, const CUtensorMap& mapO, const GemmParams& p, int ntiles, cudaStream_t st) {
  return gemm_is_simple(p) ? launch_pair2<BN, STAGES, true>(maps, mapB, mapO, p, ntiles, st) : launch_pair2<BN, STAGES, false>(maps, mapB, mapO, p, ntiles, st);
}

// B200_IMAGEN_GEMM_PAIR=1 routes the GEMMs with >= 128 columns to the CTA-pair kernel.  Off by default: on B200 it measured
// ~10 % slower than the single-CTA kernel on every shape of the workload (profiles/r01_gemm_bottleneck.txt).
bool pair_enabled() {
  static const bool on = [] {
    const char* e = getenv("B200_IMAGEN_GEMM_PAIR");
    return e != nullptr && atoi(e) != 0;
  }();
  return on;
}

}  // namespace

extern "C" int b200_conv_gemm_npad(int N) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  return (N + 127) / 128 * 128;
}

extern "C" int b200_conv_gemm_splitk(int B, int H, int W, int N, int Ktot) {
  if (B <= 0 || H <= 0 || W <= 0 || N <= 0 || Ktot <= 0) return 1;
  const int bw = W >= 128 ? 128 : next_pow2(W);
  const int bh = next_pow2(H) < 128 / bw ? next_pow2(H) : 128 / bw;
  const int bb = 128 / (bw * bh);
  const long long ntiles = (long long)((W + bw - 1) / bw) * ((H + bh - 1) / bh) * ((B + bb - 1) / bb);
  return splitk_factor(ntiles, b200_conv_gemm_npad(N), Ktot / BK);
}

extern "C" int b200_conv_gemm(const b200_src* srcs, int nsrc, const b200_seg* segs, int nseg, int B, int H, int W,
                              const void* w_packed, int N, const b200_epilogue* epi, int impl, void* f32_scratch,
                              void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(nsrc >= 1 && nsrc <= B200_MAX_SRC, "conv_gemm: nsrc %d out of range", nsrc);
  B200_REQUIRE(nseg >= 1 && nseg <= B200_MAX_SEG, "conv_gemm: nseg %d out of range", nseg);
  B200_REQUIRE(B > 0 && H > 0 && W > 0 && N > 0, "conv_gemm: bad shape B=%d H=%d W=%d N=%d", B, H, W, N);
  B200_REQUIRE(epi != nullptr && w_packed != nullptr, "conv_gemm: null pointer");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W;
  p.N = N;
  p.Npad = b200_conv_gemm_npad(N);
  p.nseg = nseg;
  {
    static const int dbg = [] { const char* e = getenv("B200_IMAGEN_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
    p.debug = dbg;
  }
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
  for (p.lbw = 0; (1 << p.lbw) < p.bw; ++p.lbw) {}
  for (p.lbh = 0; (1 << p.lbh) < p.bh; ++p.lbh) {}
  p.tiles_w = (W + p.bw - 1) / p.bw;
  p.tiles_h = (H + p.bh - 1) / p.bh;
  const int tiles_b = (B + p.bb - 1) / p.bb;
  const long long ntiles_ll = (long long)p.tiles_w * p.tiles_h * tiles_b;
  B200_REQUIRE(ntiles_ll < (1ll << 31), "conv_gemm: too many tiles");
  const int ntiles = (int)ntiles_ll;
  p.ntiles_m = ntiles;
  // epilogue
  EpiDev& e = p.epi;
  e.bias = epi->bias; e.residual = reinterpret_cast<const __nv_bfloat16*>(epi->residual);
  e.out = epi->out; e.out2 = epi->out2; e.l2_scale = epi->l2_scale;
  e.out_scale = epi->out_scale; e.act = epi->act; e.ldr = epi->ldr; e.out_mode = epi->out_mode; e.ldc = epi->ldc;
  e.ldc2 = epi->ldc2; e.split_col = epi->split_col; e.rows_per_group = epi->rows_per_group;
  e.group_stride = epi->group_stride; e.row_offset = epi->row_offset; e.l2_cols = epi->l2_cols; e.ps_C = epi->ps_C;
  e.dup_rows = epi->dup_rows;
  e.norm1 = epi->norm1; e.norm1_g = epi->norm1_g; e.norm2 = epi->norm2; e.norm2_g = epi->norm2_g; e.film = epi->film;
  e.film_ld = epi->film_ld; e.rows_per_sample = epi->rows_per_sample; e.out_norm = epi->out_norm; e.ld_norm = epi->ld_norm;
  const bool has_norm = e.norm1 != 0 || e.norm2 != 0;
  if (has_norm) {
    const auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    B200_REQUIRE(impl == 0, "conv_gemm: the norm epilogue exists on the tcgen05 path only");
    B200_REQUIRE(N >= 64 && N <= 256 && N % 32 == 0, "conv_gemm: norm epilogue needs one N tile over all channels (64 <= N <= 256, N %% 32 == 0), got N=%d", N);
    B200_REQUIRE(e.out_mode == B200_OUT_BF16 && e.split_col == 0 && e.rows_per_group == 0 && e.l2_cols == 0 && e.dup_rows == 0,
                 "conv_gemm: norm epilogue is incompatible with split / remap / l2 / dup outputs");
    B200_REQUIRE((e.norm1 == 0 || e.norm1 == 1) && e.norm2 >= 0 && e.norm2 <= 2, "conv_gemm: bad norm1/norm2 %d/%d", e.norm1, e.norm2);
    B200_REQUIRE(e.norm1 == 0 || (e.norm1_g != nullptr && al16(e.norm1_g)), "conv_gemm: norm1 needs a 16-byte aligned gain vector");
    B200_REQUIRE(e.norm2 == 0 || (e.norm2_g != nullptr && al16(e.norm2_g) && e.out_norm != nullptr && al16(e.out_norm) && (e.ld_norm & 7) == 0),
                 "conv_gemm: norm2 needs a gain vector and an aligned out_norm");
    B200_REQUIRE(e.film == nullptr || (e.norm2 == 2 && al16(e.film) && (e.film_ld & 3) == 0 && e.rows_per_sample > 0), "conv_gemm: bad FiLM operands");
    B200_REQUIRE(e.out != nullptr || e.norm2 != 0, "conv_gemm: norm epilogue without any output");
    B200_REQUIRE(e.out == nullptr || ((e.ldc & 7) == 0 && al16(e.out)), "conv_gemm: norm epilogue needs an aligned out");
    B200_REQUIRE(e.residual == nullptr || ((e.ldr & 7) == 0 && al16(e.residual)), "conv_gemm: norm epilogue needs an aligned residual");
    B200_REQUIRE((long long)B * H * W < (1ll << 31), "conv_gemm: too many rows for the norm epilogue");
  } else {
    B200_REQUIRE(epi->out != nullptr, "conv_gemm: null output");
  }
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
      const long long t2 = M;
      conv_gemm_ref_epilogue_kernel<32><<<(unsigned)ceil_div64(t2, 128), 128, 0, st>>>(p, reinterpret_cast<const float*>(f32_scratch));
    }
    B200_LAUNCH_OK();
    return B200_OK;
  }

  // ---- tensor maps
  EncodeTiledFn enc = get_encode_fn();
  B200_REQUIRE(enc != nullptr, "conv_gemm: cuTensorMapEncodeTiled not available from the driver");
  // ---- transposed kernel for MMA-bound layers with <= 128 output channels (see conv_gemm_tcT_kernel)
  {
    // off by default since the MMAs are issued under elect.sync: the 128 x 128 kernel went from 878 to 1140 TFLOP/s MMA-only and beats the
    // transposed one (46.1 vs 48.1 us on the 64x64 conv, profiles/r02_gemm_after_elect.txt); B200_IMAGEN_GEMM_T=1 selects it
    static const bool t_on = [] { const char* ev = getenv("B200_IMAGEN_GEMM_T"); return ev != nullptr && atoi(ev) != 0; }();
    const auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const long long Mll = (long long)B * H * W;
    const bool t_epi = e.out_mode == B200_OUT_BF16 && e.split_col == 0 && e.rows_per_group == 0 && e.l2_cols == 0 && e.dup_rows == 0 && e.norm1 == 0 &&
                       (N & 15) == 0 && (e.out == nullptr || ((e.ldc & 7) == 0 && al16(e.out))) && (e.out != nullptr || e.norm2 != 0) &&
                       (e.residual == nullptr || ((e.ldr & 7) == 0 && al16(e.residual))) && (e.norm2 == 0 || (al16(e.out_norm) && (e.ld_norm & 7) == 0)) &&
                       Mll < (1ll << 31);
    // measured on B200 (profiles/r02_gemm_transposed_ab.txt): plain 3x3 conv 128->128 @64x64 54.8 -> 47.6 us; the HBM-bound K = 512 linears and the
    // norm-fused K = 1152 convs are slower on it (its 8-warp transposing epilogue is not hidden behind 18 K chunks) -> long-K or norm-free only
    const bool t_shape = total >= 16 && (e.norm2 == 0 || total >= 32);
    if (t_on && p.Npad == 128 && t_epi && t_shape && Mll >= (long long)T_BP * sm_count()) {   // at least one 256-pixel tile per SM
      GemmParams pt = p;
      pt.bw = W >= T_BP ? T_BP : next_pow2(W);
      pt.bh = next_pow2(H) < T_BP / pt.bw ? next_pow2(H) : T_BP / pt.bw;
      pt.bb = T_BP / (pt.bw * pt.bh);
      for (pt.lbw = 0; (1 << pt.lbw) < pt.bw; ++pt.lbw) {}
      for (pt.lbh = 0; (1 << pt.lbh) < pt.bh; ++pt.lbh) {}
      pt.tiles_w = (W + pt.bw - 1) / pt.bw;
      pt.tiles_h = (H + pt.bh - 1) / pt.bh;
      const long long ntl = (long long)pt.tiles_w * pt.tiles_h * ((B + pt.bb - 1) / pt.bb);
      pt.ntiles_m = (int)ntl;
      pt.ksplit = 1;
      pt.tma_store = 0;
      if (pt.bb <= 256 && ntl < (1ll << 31)) {
        CUtensorMap mapsT[B200_MAX_SRC];
        memset(mapsT, 0, sizeof(mapsT));
        for (int i = 0; i < nsrc; ++i) {
          cuuint64_t dims[4] = {(cuuint64_t)srcs[i].C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
          cuuint64_t strides[3] = {(cuuint64_t)srcs[i].ld * 2, (cuuint64_t)srcs[i].ld * 2 * W, (cuuint64_t)srcs[i].ld * 2 * W * H};
          cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)pt.bw, (cuuint32_t)pt.bh, (cuuint32_t)pt.bb};
          cuuint32_t estr[4] = {1, 1, 1, 1};
          CUresult r = enc(&mapsT[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(srcs[i].ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          B200_REQUIRE(r == CUDA_SUCCESS, "conv_gemm(T): cuTensorMapEncodeTiled(X%d) failed with %d (box %d,%d,%d)", i, (int)r, pt.bw, pt.bh, pt.bb);
        }
        for (int i = nsrc; i < B200_MAX_SRC; ++i) mapsT[i] = mapsT[0];
        CUtensorMap mapW;
        {
          cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)p.Npad};
          cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
          cuuint32_t box[2] = {(cuuint32_t)BK, 128u};
          cuuint32_t estr[2] = {1, 1};
          CUresult r = enc(&mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          B200_REQUIRE(r == CUDA_SUCCESS, "conv_gemm(T): cuTensorMapEncodeTiled(W) failed with %d", (int)r);
        }
        constexpr int TS = 4;
        constexpr int smemT = TS * T_STAGE_BYTES + 1024 /*align*/ + 1024 /*barriers*/ + 2 * T_STG_BYTES;
        static_assert(smemT <= 232448, "shared memory budget");
        B200_SMEM_OPT_IN(conv_gemm_tcT_kernel<TS>, smemT);
        const int grid = (int)(ntl < sm_count() ? ntl : sm_count());
        B200_CUDA_OK(b200_launch(conv_gemm_tcT_kernel<TS>, dim3(grid), dim3(64 + 256), smemT, st, mapsT[0], mapsT[1], mapsT[2], mapsT[3], mapW, pt));
        return B200_OK;
      }
    }
  }
  // tile shape: CTA pairs (256 x BN per pair, B split across the pair) whenever there are >= 2 row tiles and >= 128 columns
  const bool pair = pair_enabled() && !has_norm && ntiles >= 2 && p.Npad >= 128;
  int BN;
  if (p.Npad <= 64) BN = p.Npad;
  else if (pair) BN = (p.Npad % 256 == 0 && (long long)((ntiles + 1) / 2) * (p.Npad / 256) >= sm_count() / 2) ? 256 : 128;
  else if (p.Npad % 256 == 0) {
    // MMA-only rates measured on B200 (tools/gemm_bench.py, debug 5): 1140 TFLOP/s for 128 x 128 tiles, 1398 for 128 x 256
    // (profiles/r02_gemm_after_elect.txt; 878 / 1398 before the issue loop was fixed): cost of a tile ~ BN / rate(BN);
    // pick the width with fewer (waves x tile cost)
    const long long t128 = (long long)ntiles * (p.Npad / 128), t256 = (long long)ntiles * (p.Npad / 256);
    const long long sms = sm_count();
    static const double rate128 = [] { const char* e = getenv("B200_IMAGEN_GEMM_RATE128"); return e ? atof(e) : 1140.0; }();   // tuning hook
    const double c128 = (double)((t128 + sms - 1) / sms) * (128.0 / rate128), c256 = (double)((t256 + sms - 1) / sms) * (256.0 / 1398.0);
    BN = c256 <= c128 ? 256 : 128;
  } else BN = 128;
  if (has_norm) BN = p.Npad;   // one tile spans all channels
  p.ksplit = 1;
  const int ks = (!pair && !has_norm && f32_scratch != nullptr) ? splitk_factor(ntiles, p.Npad, total) : 1;   // caller-provided partial-tile workspace
  if (ks > 1) BN = 256;
  const int boxN = pair ? BN / 2 : BN;
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
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)boxN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "conv_gemm: cuTensorMapEncodeTiled(B) failed with %d (Ktot=%d Npad=%d BN=%d)", (int)r, Ktot, p.Npad, BN);
  }
  // output map for the TMA-store epilogue: plain bf16 row-major destination, nothing but bias / activation / scale / L2 norm
  CUtensorMap mapO = mapB;
  p.tma_store = 0;
  {
    static const bool tma_on = [] { const char* ev = getenv("B200_IMAGEN_GEMM_TMA_STORE"); return ev == nullptr || atoi(ev) != 0; }();
    if (tma_on && !has_norm && gemm_is_simple(p) && e.out_mode == B200_OUT_BF16 && e.residual == nullptr && e.split_col == 0 && e.rows_per_group == 0 &&
        e.dup_rows == 0 && (((long long)e.ldc * 2) & 15) == 0) {
      // the 32 rows of one epilogue warp are a sub-box of the tile box {bw, bh, bb}
      const int sw_ = p.bw < 32 ? p.bw : 32;
      const int sh_ = (32 / sw_) < p.bh ? (32 / sw_) : p.bh;
      const int sb_ = 32 / (sw_ * sh_);
      cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
      cuuint64_t strides[3] = {(cuuint64_t)e.ldc * 2, (cuuint64_t)e.ldc * 2 * W, (cuuint64_t)e.ldc * 2 * W * H};
      cuuint32_t box[4] = {32u, (cuuint32_t)sw_, (cuuint32_t)sh_, (cuuint32_t)sb_};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = enc(&mapO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, e.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      B200_REQUIRE(r == CUDA_SUCCESS, "conv_gemm: cuTensorMapEncodeTiled(out) failed with %d (N=%d ldc=%d W=%d H=%d B=%d)", (int)r, N, e.ldc, W, H, B);
      p.tma_store = 1;
    }
  }
  // ---- K split across a thread-block cluster for few-tile, long-K GEMMs (conv_gemm_tcS_kernel)
  {
    static const bool ks_on = [] { const char* ev = getenv("B200_IMAGEN_GEMM_CLUSTER_K"); return ev == nullptr || atoi(ev) != 0; }();
    const long long t256 = (long long)ntiles * (p.Npad / 256);
    // measured on B200 (profiles/r02_gemm_clusterk_ab.txt): 8x8 conv 1024->1024 alone 65.5 -> 55.3 us, but inside the step only the K = 13824
    // convs gain (78 -> 72 us); K = 4608 x N = 512 (32 tiles, 4-way split) got slower (31 -> 45 us): every CTA still streams its full share of
    // the weights through L2 (19 TB/s demanded at 128 CTAs vs ~15 TB/s L2->SM) -> only very long K
    if (ks_on && !pair && !has_norm && ks <= 1 && p.Npad % 256 == 0 && gemm_is_simple(p) && e.l2_cols == 0 && t256 * 2 <= sm_count() && total >= 192) {
      int split = (int)(sm_count() / t256);
      if (split > 2) split = 2;
      while (split > 1 && total / split < 16) --split;
      if (split >= 2) {
        // the B map of this path always has a 256-row box
        CUtensorMap mapB2;
        cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)p.Npad};
        cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)KS_BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&mapB2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        B200_REQUIRE(r == CUDA_SUCCESS, "conv_gemm(K-split): cuTensorMapEncodeTiled(B) failed with %d", (int)r);
        GemmParams pk = p;
        pk.ksplit = split;
        constexpr int smemS = KS_STAGES * KS_STAGE_BYTES + 1024 /*align*/ + 1024 /*barriers*/ + 8 * EPI_STAGE_BYTES;
        static_assert(smemS <= 232448 && 128 * KS_PITCH * 4 <= KS_STAGES * KS_STAGE_BYTES, "shared memory budget");
        B200_SMEM_OPT_IN(conv_gemm_tcS_kernel, smemS);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(t256 * split));
        cfg.blockDim = dim3(64 + 256);
        cfg.dynamicSmemBytes = smemS;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)split; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        B200_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_gemm_tcS_kernel, maps[0], maps[1], maps[2], maps[3], mapB2, mapO, pk));
        return B200_OK;
      }
    }
  }
  if (ks > 1) {
    // main kernel: raw fp32 partial tiles [ks][M][Npad] into the workspace; then the finishing kernel applies the real epilogue
    GemmParams ps = p;
    ps.ksplit = ks;
    ps.N = p.Npad;
    ps.tma_store = 0;
    memset(&ps.epi, 0, sizeof(ps.epi));
    ps.epi.out = f32_scratch;
    ps.epi.out_mode = B200_OUT_F32;
    ps.epi.ldc = p.Npad;
    ps.epi.out_scale = 1.f;
    const int rc = launch_tc2<256, 4, false, 8>(maps, mapB, mapO, ps, ntiles, st);
    if (rc != B200_OK) return rc;
    const long long t2 = (long long)B * H * W * (p.Npad / 64);
    conv_gemm_splitk_epilogue_kernel<64><<<(unsigned)ceil_div64(t2, 128), 128, 0, st>>>(p, reinterpret_cast<const float*>(f32_scratch), ks);
    B200_LAUNCH_OK();
    return B200_OK;
  }
  if (has_norm) {
    switch (BN) {
      case 64: return launch_tc2<64, 8, true, 4, true>(maps, mapB, mapO, p, ntiles, st);
      case 128: return launch_tc2<128, 5, true, 16, true>(maps, mapB, mapO, p, ntiles, st);
      default: return launch_tc2<256, 3, true, 16, true>(maps, mapB, mapO, p, ntiles, st);
    }
  }
  if (pair) return BN == 256 ? launch_pair<256, 6>(maps, mapB, mapO, p, ntiles, st) : launch_pair<128, 8>(maps, mapB, mapO, p, ntiles, st);
  switch (BN) {
    // persistent, one CTA per SM: the whole 227 KB of shared memory goes to the operand ring
    case 32: return launch_tc<32, 8>(maps, mapB, mapO, p, ntiles, st);
    case 64: return launch_tc<64, 8>(maps, mapB, mapO, p, ntiles, st);
    case 128: return launch_tc<128, 6>(maps, mapB, mapO, p, ntiles, st);
    default: return launch_tc<256, 4>(maps, mapB, mapO, p, ntiles, st);
  }
}
