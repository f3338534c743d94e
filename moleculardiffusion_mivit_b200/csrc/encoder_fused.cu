// One post-norm encoder layer of the frame-token transformer as ONE persistent kernel (reference helpers/models.py:81-108
// TransformerEncoderLayerWithSkip.forward: x = LN1(x + out_proj(softmax(QK^T / sqrt(d)) V)); x = LN2(x + fc2(relu(fc1(x))))
// with MultiHeadAttention :33-59 and FeedForward :72-77).
//
// The unfused forward of a layer is 7 launches (q/k/v GEMM, attention, out-proj, LayerNorm, fc1, fc2, LayerNorm) of 10-19 us
// each for ~2.5 us of HBM traffic: a fixed latency chain per launch (barrier / TMEM set-up, weight fetch, tile fetch, MMA, TMEM
// read, transposed store).  Here a CTA owns a tile of 128 token rows = floor(128 / S) whole sequences and walks the layer with
// the tile resident on chip:
//   x tile  --TMA-->  shared memory ([row][128 B] SWIZZLE_128B boxes of 32 floats, the tcgen05 K-major A operand)
//   [q|k|v] = x W_qkv^T    ONE tcgen05.mma chain (kind::tf32, M = 128, N = 3E) -> TMEM -> +bias -> row-major tiles in shared memory
//   attention              per (sequence, head) on one warp with warp-level tensor-core MMAs (mma.sync m16n8k8 tf32; scores in
//                          3xTF32, S <= 64, d = 16) -> ctx written straight into the A-operand layout, row log-sum-exps to HBM
//   ao = ctx W_o^T         tcgen05 -> TMEM -> epilogue with thread = token row: + bias + residual, LayerNorm in registers
//   h  = relu(x1 W_1^T)    tcgen05 -> TMEM -> +bias, ReLU -> shared memory (A layout)
//   x2 = LN2(x1 + h W_2^T) tcgen05 -> TMEM -> epilogue: + bias + residual, LayerNorm
// Weights arrive by TMA (one box per 32-float slice of the reduction dimension, SWIZZLE_128B) in two alternating buffers, one
// phase ahead; activations produced on chip (ctx, x1, relu(h)) are written by the threads in the no-swizzle core-matrix order.
// Everything the backward reads -- q, k, v, ctx, z1, x1, relu(h), z2, x2 -- leaves through row-major tiles in the TMA box layout
// and TMA stores (the thread-copied version spent 25 % of its time in the copy loops, profiles/r02_ncu_encoder_fwd.md), the
// LayerNorm / softmax row statistics directly.  tf32 products, fp32 accumulation / softmax / LayerNorm.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kRows = 128;
constexpr int kThreads = 512;   // 16 warps: warp w reads TMEM lane quarter w & 3 and owns column slice w >> 2 of every epilogue

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct EncArgs {
  const float* x;   // [T, E] layer input
  const float *wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo, *g1, *be1, *w1, *bf1, *w2, *bf2, *g2, *be2;
  float *q, *k, *v, *ctx, *lse, *z1, *m1, *r1, *x1, *hact, *z2, *m2, *r2, *x2;
  int B, S, spt, n_tiles;
  float ln_eps;
};

struct EncMaps {   // fp32 SWIZZLE_128B tensor maps (tma.cuh: make_f32_tensor_map_sw), boxes of 32 floats x rows
  CUtensorMap x;                       // [T, E], 128-row boxes
  CUtensorMap wq, wk, wv, wo;          // [E, E], E-row boxes
  CUtensorMap w1;                      // [HD, E], HD-row boxes
  CUtensorMap w2;                      // [E, HD], E-row boxes
  // stores: boxes of spt * S rows (the rows a tile owns; the last tile is clipped at the end of the tensor)
  CUtensorMap q, k, v, ctx, z1, x1, z2, x2;   // [T, E]
  CUtensorMap hact;                            // [T, HD]
};

// D[128 x N] (TMEM columns [0, N)) = A[128 x K] * B^T.  B = TMA-loaded weights: K / 32 boxes of [N rows][128 B] SWIZZLE_128B
// (8 reduction elements = 32 B per MMA inside the 128-byte atom).  A = the same layout (A_SW: the TMA-loaded x tile, boxes of
// [128 rows][128 B]) or the no-swizzle core-matrix slab [K/4][128][16 B] the epilogues write.
template <bool A_SW>
__device__ __forceinline__ void issue_gemm(uint32_t tmem, const uint8_t* a, const uint8_t* b, int N, int K, uint64_t* bar) {
  const uint32_t idesc = idesc_tf32(kRows, N);
  const uint64_t da = A_SW ? tma::make_desc_sw(umma::smem_u32(a), 0u, 128u) : umma::make_desc(umma::smem_u32(a), (uint32_t)kRows * 16u, 128u);
  const uint64_t db = tma::make_desc_sw(umma::smem_u32(b), 0u, 128u);
  const uint32_t a_lo0 = (uint32_t)da, b_lo0 = (uint32_t)db;
  const uint32_t a_hi = (uint32_t)(da >> 32), b_hi = (uint32_t)(db >> 32);
  for (int ks = 0; ks < K / 8; ++ks) {
    const uint32_t kb = (uint32_t)(ks >> 2), j = (uint32_t)(ks & 3);
    const uint32_t a_lo = A_SW ? a_lo0 + kb * (uint32_t)(kRows * 8) + 2u * j : a_lo0 + (uint32_t)ks * 2u * kRows;   // 16-byte units
    const uint32_t b_lo = b_lo0 + kb * (uint32_t)(N * 8) + 2u * j;
    mma_tf32(tmem, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc, ks > 0 ? 1u : 0u);
  }
  umma::commit(bar);
}

// Row-major tiles (q, k, v for the attention; every staging tile) live in the layout TMA consumes for fp32 SWIZZLE_128B boxes of
// 32 floats: box c / 8 holds [128 rows][128 B], inside a row the 16-byte chunk c % 8 sits at position (c % 8) ^ (r % 8).  Lanes
// holding consecutive rows and the same chunk -- the thread-per-row epilogues, the MMA fragment loads -- hit every bank once,
// and a tile goes to HBM as E / 32 TMA stores instead of 2048 LDS + STG pairs.
template <int CH>
__device__ __forceinline__ uint32_t rm_off(int r, int c) {
  return (uint32_t)(c >> 3) * (kRows * 128) + (uint32_t)r * 128 + (uint32_t)(((c & 7) ^ (r & 7)) << 4);
}
template <int CH>
__device__ __forceinline__ void rm_st(uint8_t* tile, int r, int c, float4 v) {
  *reinterpret_cast<float4*>(tile + rm_off<CH>(r, c)) = v;
}
// element (row, col); rows past the tile are clamped (finite data: a masked 0 * NaN would poison the MMA)
template <int CH>
__device__ __forceinline__ float tile_el(const uint8_t* tile, int row, int col) {
  row = min(row, kRows - 1);
  return *reinterpret_cast<const float*>(tile + rm_off<CH>(row, col >> 2) + ((col & 3) << 2));
}
__device__ __forceinline__ void tma_store_box(const CUtensorMap* tm, const void* smem_src, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(col), "r"(row),
               "r"(umma::smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be reused
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// a whole E-wide tile -> rows [row0, ...) of a [T, W] matrix at column col0 (one elected thread)
template <int E>
__device__ __forceinline__ void store_tile(const CUtensorMap* tm, const uint8_t* tile, int col0, int row0) {
#pragma unroll
  for (int kb = 0; kb < E / 32; ++kb) tma_store_box(tm, tile + (size_t)kb * kRows * 128, col0 + kb * 32, row0);
}

// ---- warp-level tensor-core attention (mma.sync m16n8k8 tf32; fragment layouts: see encoder_fused_bwd.cu) -------------------
__device__ __forceinline__ void mma_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {   // 3xTF32: x = hi + lo, hi exact in tf32
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// softmax(Q K^T / sqrt(d)) V of ONE (sequence, head) on one warp: rows rb .. rb + S of the q / k / v tiles, columns hc .. hc + 16.
// Scores AND P V in 3xTF32 (error ~ fp32: the backward recomputes P from the stored log-sum-exp, and a plain-tf32 P V moved the
// forward by 1e-4 -- enough to flip borderline ReLU masks of the feed-forward block against the fp32 attention of the unfused
// path), base-2 exponentials; the score C fragments are reused as A fragments of P V (reduction index permuted inside groups of
// 8, V rows loaded with the same permutation).
// ctx goes straight into the K-major A slab of the out-projection, the natural-log row log-sum-exps to HBM.
template <int E, int CH, int SMAX>
__device__ __forceinline__ void attention_fwd_mma(const uint8_t* Qt, const uint8_t* Kt, const uint8_t* Vt, uint8_t* ctx_slab, int rb,
                                                  int hc, int S, float scale, int lane, float* __restrict__ lse) {
  constexpr int MT = SMAX / 16, NT = SMAX / 8;
  const int gid = lane >> 2, tig = lane & 3;
  const float sl2 = scale * 1.4426950408889634f;
#pragma unroll 1
  for (int mt = 0; mt < MT; ++mt) {
    const int r0 = mt * 16 + gid, r1 = r0 + 8;
    if (mt * 16 >= S) break;
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int ca = hc + ks * 8 + tig, cb = ca + 4;
      tf32_split(tile_el<CH>(Qt, rb + r0, ca) * sl2, ah[ks][0], al[ks][0]);
      tf32_split(tile_el<CH>(Qt, rb + r1, ca) * sl2, ah[ks][1], al[ks][1]);
      tf32_split(tile_el<CH>(Qt, rb + r0, cb) * sl2, ah[ks][2], al[ks][2]);
      tf32_split(tile_el<CH>(Qt, rb + r1, cb) * sl2, ah[ks][3], al[ks][3]);
    }
    float sc[NT][4];
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int cr = rb + nt * 8 + gid;
      uint32_t bh[2][2], bl[2][2];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int ca = hc + ks * 8 + tig;
        tf32_split(tile_el<CH>(Kt, cr, ca), bh[ks][0], bl[ks][0]);
        tf32_split(tile_el<CH>(Kt, cr, ca + 4), bh[ks][1], bl[ks][1]);
      }
      sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        mma_16x8x8(sc[nt], al[ks], bh[ks]);
        mma_16x8x8(sc[nt], ah[ks], bl[ks]);
        mma_16x8x8(sc[nt], ah[ks], bh[ks]);
      }
      const int c0 = nt * 8 + 2 * tig;
      if (c0 >= S) sc[nt][0] = sc[nt][2] = -INFINITY;
      if (c0 + 1 >= S) sc[nt][1] = sc[nt][3] = -INFINITY;
      m0 = fmaxf(m0, fmaxf(sc[nt][0], sc[nt][1]));
      m1 = fmaxf(m1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    // a row lives in the four lanes of a quad
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float s0 = 0.f, s1 = 0.f;
    uint32_t pah[NT][4], pal[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float p00 = exp2f(sc[nt][0] - m0), p01 = exp2f(sc[nt][1] - m0);     // masked columns: 2^(-inf) = 0
      const float p10 = exp2f(sc[nt][2] - m1), p11 = exp2f(sc[nt][3] - m1);
      s0 += p00 + p01;
      s1 += p10 + p11;
      tf32_split(p00, pah[nt][0], pal[nt][0]); tf32_split(p10, pah[nt][1], pal[nt][1]);
      tf32_split(p01, pah[nt][2], pal[nt][2]); tf32_split(p11, pah[nt][3], pal[nt][3]);
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    float o[2][4] = {};
#pragma unroll
    for (int ks = 0; ks < NT; ++ks) {
      const int ta = rb + ks * 8 + 2 * tig, tb = ta + 1;            // tokens behind slots tig and tig + 4
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int col = hc + nt * 8 + gid;
        uint32_t bh[2], bl[2];
        tf32_split(tile_el<CH>(Vt, ta, col), bh[0], bl[0]);
        tf32_split(tile_el<CH>(Vt, tb, col), bh[1], bl[1]);
        mma_16x8x8(o[nt], pal[ks], bh);
        mma_16x8x8(o[nt], pah[ks], bl);
        mma_16x8x8(o[nt], pah[ks], bh);
      }
    }
    const float i0 = 1.0f / s0, i1 = 1.0f / s1;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int col = hc + nt * 8 + 2 * tig;
      uint8_t* d = ctx_slab + ((size_t)(col >> 2) * kRows) * 16 + (size_t)((col & 3) << 2);
      if (r0 < S) *reinterpret_cast<float2*>(d + (size_t)(rb + r0) * 16) = make_float2(o[nt][0] * i0, o[nt][1] * i0);
      if (r1 < S) *reinterpret_cast<float2*>(d + (size_t)(rb + r1) * 16) = make_float2(o[nt][2] * i1, o[nt][3] * i1);
    }
    if (tig == 0) {
      if (r0 < S) lse[r0] = m0 * 0.6931471805599453f + logf(s0);
      if (r1 < S) lse[r1] = m1 * 0.6931471805599453f + logf(s1);
    }
  }
}

// LayerNorm over a token row whose E columns are split into 16-column slices over E/16 threads (one per column-slice warp):
// partial sums go through shared memory `red` ([128 rows][4 slices]), two passes (mean, then centred sum of squares) like the
// one-warp-per-token kernel.  v[] = z on entry, the normalised output on exit.  Every thread of the CTA must call it.
template <int E>
__device__ __forceinline__ void layernorm_sliced(float (&v)[16], bool active, int r, int sl, float* __restrict__ red,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                 float& mean, float& rstd) {
  constexpr int NS = E / 16;
  float s = 0.f;
  if (active) {
#pragma unroll
    for (int c = 0; c < 16; ++c) s += v[c];
    red[r * 4 + sl] = s;
  }
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int k = 0; k < NS; ++k) s += red[r * 4 + k];
  mean = s / (float)E;
  __syncthreads();
  float q = 0.f;
  if (active) {
#pragma unroll
    for (int c = 0; c < 16; ++c) { const float d = v[c] - mean; q = fmaf(d, d, q); }
    red[r * 4 + sl] = q;
  }
  __syncthreads();
  q = 0.f;
#pragma unroll
  for (int k = 0; k < NS; ++k) q += red[r * 4 + k];
  rstd = rsqrtf(q / (float)E + eps);
  __syncthreads();
  if (active) {
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = (v[c] - mean) * rstd * __ldg(gamma + sl * 16 + c) + __ldg(beta + sl * 16 + c);
  }
}

template <int E, int HD, int NH, int SMAX>
__global__ void __launch_bounds__(kThreads, 1) encoder_layer_fwd_kernel(const __grid_constant__ EncArgs a, const __grid_constant__ EncMaps tm) {
  constexpr int D = E / NH;                 // head dim (16 in every reference configuration)
  constexpr int CH = E / 4;                 // 16-byte chunks of an E-wide row
  constexpr int TILE = kRows * E * 4;       // bytes of an E-wide tile (row-major in the TMA box layout, or a K-major slab)
  constexpr int WA_BYTES = (3 * E * E > E * HD ? 3 * E * E : E * HD) * 4;
  constexpr int WB_BYTES = (E * E > HD * E ? E * E : HD * E) * 4;
  constexpr int NS = E / 16;                // active column slices of an E-wide epilogue
  static_assert(HD == 2 * E && E % 32 == 0 && D == 16 && NS <= 4, "tile shapes");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Xs = smem;                       // A operand: x (TMA), then ctx, then x1 (core-matrix slabs); staging tile of relu(h)[:, E:]
  uint8_t* Wa = Xs + TILE;                  // W_qkv, then W_1
  uint8_t* Wb = Wa + WA_BYTES;              // W_o, then W_2
  uint8_t* Qs = Wb + WB_BYTES;              // q tile; then staging tile: ctx, z1, relu(h)[:, :E], z2
  uint8_t* Ks = Qs + TILE;                  // k tile; then staging tile: x1, x2
  uint8_t* Vs = Ks + TILE;                  // v tile
  uint8_t* Hs = Ks;                         // A operand relu(h): [HD/4][128][16 B] over the k and v tiles
  float* red = reinterpret_cast<float*>(Vs + TILE);                  // [128][4] LayerNorm partial sums
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + kRows * 4);      // MMA completion
  uint64_t* lbar = bar + 1;                                          // [3] TMA completion: x + W_qkv + W_o | W_1 | W_2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lbar + 3);
  const int tid = threadIdx.x, warp = umma::warp_idx_uniform(), lane = tid & 31;
  const int qd = warp & 3, sl = warp >> 2;              // TMEM lane quarter, column slice
  const int r = qd * 32 + lane;                         // token row of this thread in every epilogue
  const int S = a.S;

  if (tid == 0) {
    umma::mbar_init(bar, 1);
    for (int i = 0; i < 3; ++i) umma::mbar_init(lbar + i, 1);
    umma::mbar_fence_init();
    tma::prefetch_map(&tm.x); tma::prefetch_map(&tm.wq); tma::prefetch_map(&tm.wk); tma::prefetch_map(&tm.wv);
    tma::prefetch_map(&tm.wo); tma::prefetch_map(&tm.w1); tma::prefetch_map(&tm.w2);
    tma::prefetch_map(&tm.q); tma::prefetch_map(&tm.k); tma::prefetch_map(&tm.v); tma::prefetch_map(&tm.ctx);
    tma::prefetch_map(&tm.z1); tma::prefetch_map(&tm.x1); tma::prefetch_map(&tm.hact); tma::prefetch_map(&tm.z2); tma::prefetch_map(&tm.x2);
  }
  if (warp == 0) umma::tmem_alloc<256>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(qd * 32) << 16);
  uint32_t parity = 0, lpar = 0;   // lpar: phase of the three load barriers (each completes once per tile)
  const float scale = rsqrtf((float)D);
  const bool act = sl < NS;

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int seq0 = tile * a.spt;
    const int nseq = min(a.spt, a.B - seq0);
    const long long row0 = (long long)seq0 * S;
    const int nrows = nseq * S;
    const bool live = r < nrows;
    // ---- x tile + W_qkv (-> Wa) + W_o (-> Wb): 10 TMA boxes, one elected thread.  (Rows of the tile past nrows hold the next
    //      tile's tokens, or zeros past the end of the tensor: every epilogue masks them with `live`, the stores do not reach them.)
    if (tid == 0) {
      bulk_wait_read0();           // the previous tile's last stores have read their staging tiles (Xs among them)
      tma::expect_tx(lbar, (uint32_t)(TILE + 4 * E * E * 4));
#pragma unroll
      for (int kb = 0; kb < E / 32; ++kb) {
        tma::load_tile(Xs + (size_t)kb * kRows * 128, &tm.x, kb * 32, (int)row0, lbar);
        tma::load_tile(Wa + ((size_t)kb * 3 * E) * 128, &tm.wq, kb * 32, 0, lbar);
        tma::load_tile(Wa + ((size_t)kb * 3 * E + E) * 128, &tm.wk, kb * 32, 0, lbar);
        tma::load_tile(Wa + ((size_t)kb * 3 * E + 2 * E) * 128, &tm.wv, kb * 32, 0, lbar);
        tma::load_tile(Wb + (size_t)kb * E * 128, &tm.wo, kb * 32, 0, lbar);
      }
    }
    // ---- [q | k | v] = x W_qkv^T
    if (warp == 4) {
      umma::mbar_wait(lbar, lpar);
      umma::fence_after_sync();
      if (umma::elect_one()) issue_gemm<true>(tmem, Xs, Wa, 3 * E, E, bar);
      __syncwarp();
    }
    umma::mbar_wait(bar, parity); parity ^= 1;
    umma::fence_after_sync();
    if (tid == 0) {                // W_1 -> Wa (free now), in flight during the q/k/v epilogue and the attention
      tma::expect_tx(lbar + 1, (uint32_t)(HD * E * 4));
#pragma unroll
      for (int kb = 0; kb < E / 32; ++kb) tma::load_tile(Wa + (size_t)kb * HD * 128, &tm.w1, kb * 32, 0, lbar + 1);
    }
#pragma unroll 1
    for (int g = sl; g < 3 * E / 16; g += 4) {            // 16-column groups of [q | k | v], dealt over the four column slices
      float v[16];
      umma::tmem_ld16(trow + (uint32_t)(g * 16), v);
      const int which = (g * 16) / E, c0 = (g * 16) % E;
      const float* bias = which == 0 ? a.bq : which == 1 ? a.bk : a.bv;
      uint8_t* dst = which == 0 ? Qs : which == 1 ? Ks : Vs;
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + c));
        rm_st<CH>(dst, r, (c0 + c) >> 2, make_float4(v[c] + bb.x, v[c + 1] + bb.y, v[c + 2] + bb.z, v[c + 3] + bb.w));
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- q, k, v to HBM (the backward reads them) and the attention: one (sequence, head) per warp
    if (tid == 0) {
      store_tile<E>(&tm.q, Qs, 0, (int)row0);
      store_tile<E>(&tm.k, Ks, 0, (int)row0);
      store_tile<E>(&tm.v, Vs, 0, (int)row0);
      bulk_commit();
    }
    for (int pair = warp; pair < nseq * NH; pair += kThreads / 32) {
      const int sq = pair / NH, h = pair % NH;
      attention_fwd_mma<E, CH, SMAX>(Qs, Ks, Vs, Xs, sq * S, h * D, S, scale, lane, a.lse + ((size_t)(seq0 + sq) * NH + h) * S);
    }
    if (tid == 0) bulk_wait_read0();   // the q / k / v stores have read the tiles: Qs becomes the ctx staging tile
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- ao = ctx W_o^T ; ctx to HBM meanwhile (core-matrix slab -> row-major tile -> TMA store)
    if (warp == 4) {
      umma::fence_after_sync();
      if (umma::elect_one()) issue_gemm<false>(tmem, Xs, Wb, E, E, bar);
      __syncwarp();
    }
    for (int i = tid; i < kRows * CH; i += kThreads) {   // lanes walk rows: conflict-free on both sides
      const int c = i / kRows, rr = i % kRows;
      rm_st<CH>(Qs, rr, c, *reinterpret_cast<const float4*>(Xs + ((size_t)c * kRows + rr) * 16));
    }
    umma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      store_tile<E>(&tm.ctx, Qs, 0, (int)row0);
      bulk_commit();
    }
    float4 xres[4];                // the residual x of this thread's 16 columns (L2: the TMA load above brought the rows in): in flight during the MMA
#pragma unroll
    for (int c = 0; c < 4; ++c)
      xres[c] = (act && live) ? __ldg(reinterpret_cast<const float4*>(a.x + (row0 + r) * E + sl * 16) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    umma::mbar_wait(bar, parity); parity ^= 1;
    umma::fence_after_sync();
    if (tid == 0) bulk_wait_read0();
    __syncthreads();               // ctx staging consumed, W_o consumed: Wb and Qs / Ks are free
    if (tid == 0) {                // W_2 -> Wb, in flight during LayerNorm 1 and fc1
      tma::expect_tx(lbar + 2, (uint32_t)(HD * E * 4));
#pragma unroll
      for (int kb = 0; kb < HD / 32; ++kb) tma::load_tile(Wb + (size_t)kb * E * 128, &tm.w2, kb * 32, 0, lbar + 2);
    }
    float x1v[16];                 // this thread's 16 columns of the x1 row stay in registers for the second residual
    {
      float v[16];
      if (act) {
        umma::tmem_ld16(trow + (uint32_t)(sl * 16), v);
        if (live) {
#pragma unroll
          for (int c = 0; c < 16; c += 4) {
            const float4 xx = xres[c / 4], bb = __ldg(reinterpret_cast<const float4*>(a.bo + sl * 16 + c));
            v[c] += xx.x + bb.x; v[c + 1] += xx.y + bb.y; v[c + 2] += xx.z + bb.z; v[c + 3] += xx.w + bb.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < 16; c += 4)          // z1 -> staging (Qs)
          rm_st<CH>(Qs, r, sl * 4 + c / 4, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
      }
      umma::fence_before_sync();
      float mean, rstd;
      layernorm_sliced<E>(v, act, r, sl, red, a.g1, a.be1, a.ln_eps, mean, rstd);
      if (act) {
        if (live && sl == 0) { a.m1[row0 + r] = mean; a.r1[row0 + r] = rstd; }
#pragma unroll
        for (int c = 0; c < 16; ++c) x1v[c] = live ? v[c] : 0.f;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {        // x1 -> A operand (Xs) and staging (Ks)
          const float4 t = make_float4(x1v[c], x1v[c + 1], x1v[c + 2], x1v[c + 3]);
          *reinterpret_cast<float4*>(Xs + ((size_t)((sl * 16 + c) / 4) * kRows + r) * 16) = t;
          rm_st<CH>(Ks, r, sl * 4 + c / 4, t);
        }
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- h = relu(x1 W_1^T + b) ; z1, x1 to HBM meanwhile
    if (warp == 4) {
      umma::mbar_wait(lbar + 1, lpar);   // W_1 has landed
      umma::fence_after_sync();
      if (umma::elect_one()) issue_gemm<false>(tmem, Xs, Wa, HD, E, bar);
      __syncwarp();
    }
    if (tid == 0) {
      store_tile<E>(&tm.z1, Qs, 0, (int)row0);
      store_tile<E>(&tm.x1, Ks, 0, (int)row0);
      bulk_commit();
    }
    umma::mbar_wait(bar, parity); parity ^= 1;
    umma::fence_after_sync();
    if (tid == 0) bulk_wait_read0();
    __syncthreads();               // staging tiles consumed, the x1 slab consumed: Hs (over Ks / Vs), Qs and Xs may be written
    static_assert(HD / E == 2, "two hidden chunks: staged in Qs and Xs");
#pragma unroll
    for (int chunk = 0; chunk < HD / E; ++chunk) {          // E hidden columns at a time: A operand + staging tile
      if (act) {
        const int c0 = chunk * E + sl * 16;
        float v[16];
        umma::tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bf1 + c0 + c));
          float4 t = make_float4(fmaxf(v[c] + bb.x, 0.f), fmaxf(v[c + 1] + bb.y, 0.f), fmaxf(v[c + 2] + bb.z, 0.f),
                                 fmaxf(v[c + 3] + bb.w, 0.f));
          if (!live) t = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(Hs + ((size_t)((c0 + c) / 4) * kRows + r) * 16) = t;
          rm_st<CH>(chunk == 0 ? Qs : Xs, r, sl * 4 + c / 4, t);
        }
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- x2 = LN2(x1 + h W_2^T + b) ; relu(h) to HBM meanwhile
    if (warp == 4) {
      umma::mbar_wait(lbar + 2, lpar);   // W_2 has landed
      umma::fence_after_sync();
      if (umma::elect_one()) issue_gemm<false>(tmem, Hs, Wb, E, HD, bar);
      __syncwarp();
    }
    if (tid == 0) {
      store_tile<E>(&tm.hact, Qs, 0, (int)row0);
      store_tile<E>(&tm.hact, Xs, E, (int)row0);
      bulk_commit();
    }
    umma::mbar_wait(bar, parity); parity ^= 1;
    umma::fence_after_sync();
    if (tid == 0) bulk_wait_read0();
    __syncthreads();               // Qs is free for z2; Ks (under Hs) has been consumed by the MMA
    {
      float z[16];
      if (act) {
        umma::tmem_ld16(trow + (uint32_t)(sl * 16), z);
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bf2 + sl * 16 + c));
          z[c] += x1v[c] + bb.x; z[c + 1] += x1v[c + 1] + bb.y; z[c + 2] += x1v[c + 2] + bb.z; z[c + 3] += x1v[c + 3] + bb.w;
        }
#pragma unroll
        for (int c = 0; c < 16; c += 4) rm_st<CH>(Qs, r, sl * 4 + c / 4, make_float4(z[c], z[c + 1], z[c + 2], z[c + 3]));
      }
      umma::fence_before_sync();
      float mean, rstd;
      layernorm_sliced<E>(z, act, r, sl, red, a.g2, a.be2, a.ln_eps, mean, rstd);
      if (act) {
        if (live && sl == 0) { a.m2[row0 + r] = mean; a.r2[row0 + r] = rstd; }
#pragma unroll
        for (int c = 0; c < 16; c += 4) rm_st<CH>(Ks, r, sl * 4 + c / 4, make_float4(z[c], z[c + 1], z[c + 2], z[c + 3]));
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();               // the next tile overwrites every buffer (its first shared-memory writes are TMA loads by thread 0)
    if (tid == 0) {
      store_tile<E>(&tm.z2, Qs, 0, (int)row0);
      store_tile<E>(&tm.x2, Ks, 0, (int)row0);
      bulk_commit();
    }
    lpar ^= 1;
  }
  if (tid == 0) bulk_wait0();
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<256>(tmem);
}

int sm_count() {
  static int sms = 0;
  if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  return sms ? sms : 148;
}

template <int E, int HD, int NH, int SMAX>
int launch_fwd(const EncArgs& a, cudaStream_t st) {
  constexpr int WA = (3 * E * E > E * HD ? 3 * E * E : E * HD) * 4, WB = (E * E > HD * E ? E * E : HD * E) * 4;
  const int smem = 4 * kRows * E * 4 + WA + WB + kRows * 4 * 4 + 64;
  auto kern = encoder_layer_fwd_kernel<E, HD, NH, SMAX>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = a.n_tiles < sm_count() ? a.n_tiles : sm_count();
  const double T = (double)a.B * a.S;
  EncMaps tm;
  {
    const long long rows = (long long)a.B * a.S;
    const int own = a.spt * a.S;
    int rc = make_f32_tensor_map_sw(&tm.x, a.x, E, rows, kRows);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.wq, a.wq, E, E, E);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.wk, a.wk, E, E, E);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.wv, a.wv, E, E, E);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.wo, a.wo, E, E, E);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.w1, a.w1, E, HD, HD);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.w2, a.w2, HD, E, E);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.q, a.q, E, rows, own);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.k, a.k, E, rows, own);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.v, a.v, E, rows, own);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.ctx, a.ctx, E, rows, own);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.z1, a.z1, E, rows, own);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.x1, a.x1, E, rows, own);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.z2, a.z2, E, rows, own);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.x2, a.x2, E, rows, own);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.hact, a.hact, HD, rows, own);
    if (rc) return rc;
  }
  MivitProfScope prof("encoder_layer_fwd", 2.0 * T * (4.0 * E * E + 2.0 * E * HD) + 4.0 * a.B * NH * (double)a.S * a.S * (E / NH), st);
  kern<<<grid, kThreads, smem, st>>>(a, tm);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

}  // namespace

bool encoder_fused_supported(int B, int S, int E, int HD, int H) {
  if (!(S >= 1 && S <= 64 && H >= 1 && E == 16 * H)) return false;
  if (!((E == 64 && HD == 128) || (E == 32 && HD == 64))) return false;
  return (long long)B * S >= 512;      // small batches keep the fp32 SIMT path (and its 1e-4 parity tests)
}

int encoder_layer_fwd(const EncoderLayerIO& io, int B, int S, int E, int HD, int H, float ln_eps, cudaStream_t st) {
  MIVIT_CHECK_ARG(encoder_fused_supported(B, S, E, HD, H), "fused encoder layer: unsupported shape");
  EncArgs a;
  a.x = io.x;
  a.wq = io.wq; a.bq = io.bq; a.wk = io.wk; a.bk = io.bk; a.wv = io.wv; a.bv = io.bv; a.wo = io.wo; a.bo = io.bo;
  a.g1 = io.g1; a.be1 = io.be1; a.w1 = io.w1; a.bf1 = io.bf1; a.w2 = io.w2; a.bf2 = io.bf2; a.g2 = io.g2; a.be2 = io.be2;
  a.q = io.q; a.k = io.k; a.v = io.v; a.ctx = io.ctx; a.lse = io.lse; a.z1 = io.z1; a.m1 = io.m1; a.r1 = io.r1; a.x1 = io.x1;
  a.hact = io.hact; a.z2 = io.z2; a.m2 = io.m2; a.r2 = io.r2; a.x2 = io.x2;
  a.B = B; a.S = S; a.spt = kRows / S; a.n_tiles = (B + a.spt - 1) / a.spt; a.ln_eps = ln_eps;
  if (E == 64) return S <= 32 ? launch_fwd<64, 128, 4, 32>(a, st) : launch_fwd<64, 128, 4, 64>(a, st);
  return S <= 32 ? launch_fwd<32, 64, 2, 32>(a, st) : launch_fwd<32, 64, 2, 64>(a, st);
}
