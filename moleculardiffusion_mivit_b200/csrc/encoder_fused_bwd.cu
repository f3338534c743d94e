// Input-gradient half of the backward of one post-norm encoder layer as ONE persistent kernel (reference
// helpers/models.py:81-108 TransformerEncoderLayerWithSkip, differentiated by autograd there):
//   dz2  = LN2'(dy)                                   thread = token row, 16-column slices, cross-slice sums in shared memory
//   dh2  = (dz2 W_2) * 1[relu(h) > 0]                 tcgen05 kind::tf32, M = 128 tokens, N = HD
//   dz1  = LN1'(dz2 + dh2 W_1)                        tcgen05, N = E; residual from registers
//   dctx = dz1 W_o                                    tcgen05, N = E
//   dq, dk, dv = attention'(q, k, v, lse, ctx, dctx)  one (sequence, head) per warp, flash style (P recomputed from the row
//                                                     log-sum-exps), operands and results in shared memory
//   dx   = dz1 + dq W_q + dk W_k + dv W_v             three tcgen05 chains into one accumulator
// The unfused path runs this as 9 launches per layer (two LayerNorm backwards, four tf32 GEMM launches, the attention
// backward, ...), each a 10-19 us latency chain over ~2.5 us of HBM traffic.  Here a CTA owns a tile of 128 token rows =
// floor(128 / S) whole sequences and keeps it on chip; what the weight-gradient GEMMs need (dz2, dh2, dz1, dq, dk, dv) is
// written once, coalesced, through an XOR-swizzled row-major staging tile.  The weight gradients stay in linear_tc.cu: their
// reduction runs over ALL tokens, so they are separate (batched) launches behind this kernel.
//
// The input-gradient GEMMs multiply by W (not W^T); instead of an MN-major B operand the (tiny) weights are transposed once
// per step by pack_kernel, so every GEMM here is the same K-major x K-major form as the forward kernel (encoder_fused.cu).
// Weights (and, for S <= 32, the dq / dk / dv slabs) are staged by TMA: fp32 boxes of 32 reduction elements x rows that land as
// [row][128 B] SWIZZLE_128B tiles, completion on mbarriers; A operands produced on chip stay in the no-swizzle core-matrix order.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kRows = 128;
constexpr int kThreads = 512;   // 16 warps: warp w reads TMEM lane quarter w & 3 and owns column slice w >> 2 of every epilogue

__device__ __forceinline__ void tma_store_box(const CUtensorMap* tm, const void* smem_src, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(col), "r"(row),
               "r"(umma::smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be reused
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
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

struct BwdArgs {
  const float* dy;   // [T, E] gradient of the layer output
  float* dx;         // [T, E] gradient of the layer input (may alias dy: a tile reads its dy rows before it writes them)
  const float *z2, *m2, *r2, *hact, *z1, *m1, *r1, *q, *k, *v, *lse, *ctx;
  const float *g2, *g1;                     // LayerNorm weights
  const float *w2t, *w1t, *wot, *wqkvt;     // transposed weights (pack_kernel)
  float *dz2, *dh2, *dz1, *dq, *dk, *dv;    // operands of the weight-gradient launches
  float *dg2, *db2, *dg1, *db1;             // LayerNorm parameter gradients (accumulated with atomics)
  int B, S, spt, n_tiles;
};

struct BwdMaps {   // fp32 SWIZZLE_128B tensor maps (tma.cuh: make_f32_tensor_map_sw), boxes of 32 floats x rows
  CUtensorMap w2t;           // [HD, E],  HD-row boxes
  CUtensorMap wot;           // [E, E],   E-row boxes
  CUtensorMap w1t;           // [E, HD],  E-row boxes
  CUtensorMap wqkvt;         // [E, 3E],  E-row boxes
  CUtensorMap dq, dk, dv;    // [T, E],   128-row boxes (read back as A slabs, S <= 32)
  CUtensorMap q, k, v;       // [T, E],   128-row boxes (the attention's row-major tiles)
  CUtensorMap dy, z2;        // [T, E],   128-row boxes (inputs of LayerNorm 2', prefetched for the next tile)
  CUtensorMap dz2, dh2, dx;  // stores: [T, E] / [T, HD] / [T, E], boxes of spt * S rows (the rows a tile owns)
};

// D[128 x N] (TMEM columns from `tmem`) (+)= A[128 x K] * B^T.  B = TMA-loaded weights: K / 32 boxes of [N rows][128 B]
// SWIZZLE_128B; A = the same layout (A_SW: boxes of [128 rows][128 B]) or the no-swizzle core-matrix slab [K/4][128][16 B].
template <bool A_SW>
__device__ __forceinline__ void issue_chain(uint32_t tmem, const uint8_t* a, const uint8_t* b, int N, int K, bool first) {
  const uint32_t idesc = idesc_tf32(kRows, N);
  const uint64_t da = A_SW ? tma::make_desc_sw(umma::smem_u32(a), 0u, 128u) : umma::make_desc(umma::smem_u32(a), (uint32_t)kRows * 16u, 128u);
  const uint64_t db = tma::make_desc_sw(umma::smem_u32(b), 0u, 128u);
  const uint32_t a_lo0 = (uint32_t)da, b_lo0 = (uint32_t)db;
  const uint32_t a_hi = (uint32_t)(da >> 32), b_hi = (uint32_t)(db >> 32);
  for (int ks = 0; ks < K / 8; ++ks) {
    const uint32_t kb = (uint32_t)(ks >> 2), j = (uint32_t)(ks & 3);
    const uint32_t a_lo = A_SW ? a_lo0 + kb * (uint32_t)(kRows * 8) + 2u * j : a_lo0 + (uint32_t)ks * 2u * kRows;   // 16-byte units
    const uint32_t b_lo = b_lo0 + kb * (uint32_t)(N * 8) + 2u * j;
    mma_tf32(tmem, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc, (first && ks == 0) ? 0u : 1u);
  }
}

// Row-major tiles live in the layout TMA produces (and consumes) for fp32 SWIZZLE_128B boxes of 32 floats: box c / 8 holds
// [128 rows][128 B], and inside a row the 16-byte chunk c % 8 sits at position (c % 8) ^ (r % 8).  The q / k / v tiles are TMA
// loads, the staging tiles TMA stores; lanes holding consecutive rows and the same chunk (every thread-per-row access, and the
// MMA fragment loads of the attention) hit every bank once.
// (Measured alternative for the thread-filled version of these tiles: swizzling whole 64-byte head blocks, so that the SIMT
// attention loops address a block with one computed offset + immediates, executed 7 % fewer instructions but paid 8-way bank
// conflicts on every thread-per-row access: 121.5 vs 106.9 us per launch under ncu, profiles/r02_ncu_encoder_bwd.md.)
template <int CH>
__device__ __forceinline__ uint32_t rm_off(int r, int c) {
  return (uint32_t)(c >> 3) * (kRows * 128) + (uint32_t)r * 128 + (uint32_t)(((c & 7) ^ (r & 7)) << 4);
}
template <int CH>
__device__ __forceinline__ float4 rm_ld(const uint8_t* tile, int r, int c) {
  return *reinterpret_cast<const float4*>(tile + rm_off<CH>(r, c));
}
template <int CH>
__device__ __forceinline__ void rm_st(uint8_t* tile, int r, int c, float4 v) {
  *reinterpret_cast<float4*>(tile + rm_off<CH>(r, c)) = v;
}

// swizzled tile -> global rows [row0, row0 + nrows) of a [.., gw] matrix at column offset gc (coalesced 16-byte stores)
template <int CH>
__device__ __forceinline__ void copy_out(const uint8_t* tile, float* __restrict__ g, long long row0, int nrows, int gw, int gc, int tid) {
  for (int i = tid; i < nrows * CH; i += kThreads) {
    const int r = i / CH, c = i % CH;
    *reinterpret_cast<float4*>(g + (row0 + r) * gw + gc + 4 * c) = rm_ld<CH>(tile, r, c);
  }
}
// swizzled row-major tile -> K-major A slab [CH][128][16 B]; lanes walk rows (conflict-free on both sides)
template <int CH>
__device__ __forceinline__ void tile_to_slab(const uint8_t* tile, uint8_t* slab, int tid) {
  for (int i = tid; i < kRows * CH; i += kThreads) {
    const int c = i / kRows, r = i % kRows;
    *reinterpret_cast<float4*>(slab + ((size_t)c * kRows + r) * 16) = rm_ld<CH>(tile, r, c);
  }
}

// 32 values per lane -> lane l returns the sum over the warp's lanes of value l (31 shuffles instead of 160)
template <int N>
__device__ __forceinline__ void tr_step(float (&v)[32], int lane) {
  const bool up = (lane & N) != 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float keep = up ? v[i + N] : v[i];
    const float send = up ? v[i] : v[i + N];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, N);
  }
}
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
  tr_step<16>(v, lane); tr_step<8>(v, lane); tr_step<4>(v, lane); tr_step<2>(v, lane); tr_step<1>(v, lane);
  return v[0];
}

// ---- warp-level tensor-core pieces of the attention backward (S <= 32): mma.sync m16n8k8 tf32 --------------------------------
// Fragment layouts (PTX ISA, m16n8k8 .tf32; gid = lane / 4, tig = lane % 4):
//   A 16x8 (row):  a0 = (gid, tig)      a1 = (gid + 8, tig)      a2 = (gid, tig + 4)      a3 = (gid + 8, tig + 4)
//   B  8x8 (col):  b0 = (k = tig, n = gid)                       b1 = (k = tig + 4, n = gid)
//   C 16x8:        c0 = (gid, 2 tig)    c1 = (gid, 2 tig + 1)    c2 = (gid + 8, 2 tig)    c3 = (gid + 8, 2 tig + 1)
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
// 3xTF32 split: x = hi + lo with hi exactly representable in tf32 (the score GEMMs must reproduce the forward's fp32 scores:
// P_ij = exp(s_ij - L_i) is only normalised if s_ij matches the s_ij the log-sum-exp L_i was computed from)
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// element (row, col) of a swizzled row-major tile; rows past the tile are clamped (finite data: a masked 0 * NaN would poison the MMA)
template <int CH>
__device__ __forceinline__ float tile_el(const uint8_t* tile, int row, int col) {
  row = min(row, kRows - 1);
  return *reinterpret_cast<const float*>(tile + rm_off<CH>(row, col >> 2) + ((col & 3) << 2));
}
// Attention backward of ONE (sequence, head) on one warp with tensor-core MMAs (S <= 32, head dim 16).  Q / K / V / dctx are
// the swizzled tiles Qt / Kt / Vt / Gt (rows rb .. rb + S of the CTA's tile, columns hc .. hc + 16); Ls = this warp's row
// log-sum-exps (base 2), Dl = scratch for D_i = dctx_i . ctx_i.  Two orientations, each in two halves of 16 rows:
//   rows = queries i:  s = (scale Q) K^T (3xTF32), dP = G V^T  ->  D_i = sum_j P dP, dS = P (dP - D)  ->  dQ = scale dS K
//   rows = keys j:     s^T = (scale K) Q^T,        dP^T = V G^T ->  P^T, dS^T        ->  dV = P^T G,  dK = scale dS^T Q
// The C fragments of the first stage are the A fragments of the second with the reduction index permuted inside each group of 8
// (slot tig <-> token 2 tig, slot tig + 4 <-> token 2 tig + 1); the B fragments of the second stage are loaded with the same
// permutation, so nothing moves between lanes.  Results go straight to HBM (the weight-gradient launches read them there; the
// CTA reads them back from L2 as K-major slabs for the q/k/v input-gradient GEMM).
template <int E, int CH>
__device__ __forceinline__ void attention_bwd_mma(const uint8_t* Qt, const uint8_t* Kt, const uint8_t* Vt, const uint8_t* Gt,
                                                  const float* Ls, float* Dl, int rb, int hc, int S, float scale, int lane,
                                                  float* __restrict__ gdq, float* __restrict__ gdk, float* __restrict__ gdv) {
  const int gid = lane >> 2, tig = lane & 3;
  constexpr float kLog2e = 1.4426950408889634f;   // P = 2^(s log2e - L log2e): Ls holds L log2e, the scaled operand carries log2e
  const float sl2 = scale * kLog2e;
#pragma unroll 1
  for (int orient = 0; orient < 2; ++orient) {
    // orient 0: rows = queries (A operands Q, G; B operands K, V; second stage B = K -> dQ)
    // orient 1: rows = keys    (A operands K, V; B operands Q, G; second stage B = G -> dV and B = Q -> dK)
    const uint8_t* A1 = orient == 0 ? Qt : Kt;     // scaled, 3xTF32
    const uint8_t* A2 = orient == 0 ? Gt : Vt;
    const uint8_t* B1 = orient == 0 ? Kt : Qt;     // 3xTF32
    const uint8_t* B2 = orient == 0 ? Vt : Gt;
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      const int r0 = mt * 16 + gid, r1 = r0 + 8;                    // this lane's two rows (of the orientation)
      // both first-stage products in 3xTF32: dS = P (dP - D) cancels (dP_ij ~ D_i for diffuse attention), so a plain tf32 dP
      // (1e-3 relative) showed up as 9 % on the last layer's q_proj gradient against the fp32 oracle
      uint32_t ah[2][4], al[2][4], a2h[2][4], a2l[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int ca = hc + ks * 8 + tig, cb = ca + 4;
        tf32_split(tile_el<CH>(A1, rb + r0, ca) * sl2, ah[ks][0], al[ks][0]);
        tf32_split(tile_el<CH>(A1, rb + r1, ca) * sl2, ah[ks][1], al[ks][1]);
        tf32_split(tile_el<CH>(A1, rb + r0, cb) * sl2, ah[ks][2], al[ks][2]);
        tf32_split(tile_el<CH>(A1, rb + r1, cb) * sl2, ah[ks][3], al[ks][3]);
        tf32_split(tile_el<CH>(A2, rb + r0, ca), a2h[ks][0], a2l[ks][0]);
        tf32_split(tile_el<CH>(A2, rb + r1, ca), a2h[ks][1], a2l[ks][1]);
        tf32_split(tile_el<CH>(A2, rb + r0, cb), a2h[ks][2], a2l[ks][2]);
        tf32_split(tile_el<CH>(A2, rb + r1, cb), a2h[ks][3], a2l[ks][3]);
      }
      // orientation 0: L belongs to the rows; orientation 1: L and D to the columns.  D_i = dctx_i . ctx_i = sum_j P_ij dP_ij is
      // taken from the fragments of orientation 0 (a row lives in the four lanes of a quad) instead of re-reading ctx from HBM.
      const float Lr0 = Ls[r0], Lr1 = Ls[r1];
      float pv[4][4], dv[4][4];        // P and dP (then dS) of this lane's 2 rows x 8 columns
      float Dr0 = 0.f, Dr1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int cr = rb + nt * 8 + gid;                            // row of the B-operand tiles
        uint32_t bh[2][2], bl[2][2], b2h[2][2], b2l[2][2];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const int ca = hc + ks * 8 + tig;
          tf32_split(tile_el<CH>(B1, cr, ca), bh[ks][0], bl[ks][0]);
          tf32_split(tile_el<CH>(B1, cr, ca + 4), bh[ks][1], bl[ks][1]);
          tf32_split(tile_el<CH>(B2, cr, ca), b2h[ks][0], b2l[ks][0]);
          tf32_split(tile_el<CH>(B2, cr, ca + 4), b2h[ks][1], b2l[ks][1]);
        }
        float sc[4] = {0.f, 0.f, 0.f, 0.f};
        dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          mma_16x8x8(sc, al[ks], bh[ks]);
          mma_16x8x8(sc, ah[ks], bl[ks]);
          mma_16x8x8(sc, ah[ks], bh[ks]);
          mma_16x8x8(dv[nt], a2l[ks], b2h[ks]);
          mma_16x8x8(dv[nt], a2h[ks], b2l[ks]);
          mma_16x8x8(dv[nt], a2h[ks], b2h[ks]);
        }
        const int c0 = nt * 8 + 2 * tig, c1 = c0 + 1;               // this lane's two columns
        float L00, L01, L10, L11;                                   // (row r0 | r1, column c0 | c1)
        if (orient == 0) {
          L00 = L01 = Lr0; L10 = L11 = Lr1;
        } else {
          L00 = L10 = Ls[c0]; L01 = L11 = Ls[c1];
        }
        const bool v00 = r0 < S && c0 < S, v01 = r0 < S && c1 < S, v10 = r1 < S && c0 < S, v11 = r1 < S && c1 < S;
        pv[nt][0] = v00 ? exp2f(sc[0] - L00) : 0.f; pv[nt][1] = v01 ? exp2f(sc[1] - L01) : 0.f;
        pv[nt][2] = v10 ? exp2f(sc[2] - L10) : 0.f; pv[nt][3] = v11 ? exp2f(sc[3] - L11) : 0.f;
        if (!v00) dv[nt][0] = 0.f;
        if (!v01) dv[nt][1] = 0.f;
        if (!v10) dv[nt][2] = 0.f;
        if (!v11) dv[nt][3] = 0.f;
        Dr0 = fmaf(pv[nt][0], dv[nt][0], fmaf(pv[nt][1], dv[nt][1], Dr0));
        Dr1 = fmaf(pv[nt][2], dv[nt][2], fmaf(pv[nt][3], dv[nt][3], Dr1));
      }
      if (orient == 0) {
        Dr0 += __shfl_xor_sync(0xffffffffu, Dr0, 1); Dr0 += __shfl_xor_sync(0xffffffffu, Dr0, 2);
        Dr1 += __shfl_xor_sync(0xffffffffu, Dr1, 1); Dr1 += __shfl_xor_sync(0xffffffffu, Dr1, 2);
        if (tig == 0) { Dl[r0] = Dr0; Dl[r1] = Dr1; }   // for orientation 1 (rows >= S: masked there)
      }
      uint32_t pa[4][4], da[4][4];     // second-stage A fragments: P (orientation 1 only) and dS, one k-step per first-stage n-tile
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c0 = nt * 8 + 2 * tig, c1 = c0 + 1;
        float D00, D01, D10, D11;
        if (orient == 0) {
          D00 = D01 = Dr0; D10 = D11 = Dr1;
        } else {
          D00 = D10 = Dl[c0]; D01 = D11 = Dl[c1];
        }
        // (masked entries: P = 0 and dP was zeroed above, but D of a masked row / column may be anything: select, not multiply)
        const bool v00 = r0 < S && c0 < S, v01 = r0 < S && c1 < S, v10 = r1 < S && c0 < S, v11 = r1 < S && c1 < S;
        const float d00 = v00 ? pv[nt][0] * (dv[nt][0] - D00) : 0.f, d01 = v01 ? pv[nt][1] * (dv[nt][1] - D01) : 0.f;
        const float d10 = v10 ? pv[nt][2] * (dv[nt][2] - D10) : 0.f, d11 = v11 ? pv[nt][3] * (dv[nt][3] - D11) : 0.f;
        // A fragment of k-step nt: slot tig <- column c0, slot tig + 4 <- column c1
        da[nt][0] = tf32_rna(d00); da[nt][1] = tf32_rna(d10); da[nt][2] = tf32_rna(d01); da[nt][3] = tf32_rna(d11);
        if (orient == 1) {
          pa[nt][0] = tf32_rna(pv[nt][0]); pa[nt][1] = tf32_rna(pv[nt][2]); pa[nt][2] = tf32_rna(pv[nt][1]); pa[nt][3] = tf32_rna(pv[nt][3]);
        }
      }
      float o1[2][4] = {}, o2[2][4] = {};     // orientation 0: o1 = dQ;  orientation 1: o1 = dK, o2 = dV
      const uint8_t* S1 = orient == 0 ? Kt : Qt;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int ta = rb + ks * 8 + 2 * tig, tb = ta + 1;            // tokens behind slots tig and tig + 4
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int col = hc + nt * 8 + gid;
          uint32_t b[2] = {tf32_rna(tile_el<CH>(S1, ta, col)), tf32_rna(tile_el<CH>(S1, tb, col))};
          mma_16x8x8(o1[nt], da[ks], b);
          if (orient == 1) {
            uint32_t bg[2] = {tf32_rna(tile_el<CH>(Gt, ta, col)), tf32_rna(tile_el<CH>(Gt, tb, col))};
            mma_16x8x8(o2[nt], pa[ks], bg);
          }
        }
      }
      float* g1 = orient == 0 ? gdq : gdk;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int col = hc + nt * 8 + 2 * tig;
        if (r0 < S) {
          *reinterpret_cast<float2*>(g1 + (size_t)(rb + r0) * E + col) = make_float2(o1[nt][0] * scale, o1[nt][1] * scale);
          if (orient == 1) *reinterpret_cast<float2*>(gdv + (size_t)(rb + r0) * E + col) = make_float2(o2[nt][0], o2[nt][1]);
        }
        if (r1 < S) {
          *reinterpret_cast<float2*>(g1 + (size_t)(rb + r1) * E + col) = make_float2(o1[nt][2] * scale, o1[nt][3] * scale);
          if (orient == 1) *reinterpret_cast<float2*>(gdv + (size_t)(rb + r1) * E + col) = make_float2(o2[nt][2], o2[nt][3]);
        }
      }
    }
    __syncwarp();   // orientation 1 reads the D_i orientation 0 wrote
  }
}

// 16 floats of a row's head block (4 chunks from chunk c0) of a swizzled tile
template <int CH, int D>
__device__ __forceinline__ void ld_head(const uint8_t* tile, int row, int c0, float (&o)[D], float mul) {
#pragma unroll
  for (int j = 0; j < D / 4; ++j) {
    const float4 t = rm_ld<CH>(tile, row, c0 + j);
    o[4 * j] = t.x * mul; o[4 * j + 1] = t.y * mul; o[4 * j + 2] = t.z * mul; o[4 * j + 3] = t.w * mul;
  }
}
template <int CH, int D>
__device__ __forceinline__ void st_head(uint8_t* tile, int row, int c0, const float (&o)[D], float mul) {
#pragma unroll
  for (int j = 0; j < D / 4; ++j)
    rm_st<CH>(tile, row, c0 + j, make_float4(o[4 * j] * mul, o[4 * j + 1] * mul, o[4 * j + 2] * mul, o[4 * j + 3] * mul));
}

// LayerNorm backward of a token row whose E columns are split into 16-column slices over the column-slice warps.
// On entry dyv = upstream gradient, zv = the saved pre-norm sum; on exit dz = the input gradient (0 for dead rows: rs = 0),
// and the row's contributions to dgamma / dbeta have been added to acc[0 .. E) / acc[E .. 2E).  Every thread must call it.
template <int E>
__device__ __forceinline__ void layernorm_bwd_sliced(const float (&dyv)[16], const float (&zv)[16], float m, float rs, bool act,
                                                     int r, int sl, int lane, float* __restrict__ red,
                                                     const float* __restrict__ gamma, float* __restrict__ acc, float (&dz)[16]) {
  constexpr int NS = E / 16;
  float xh[16], g[16];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    xh[c] = (zv[c] - m) * rs;
    g[c] = act ? dyv[c] * __ldg(gamma + sl * 16 + c) : 0.f;
    s1 += g[c];
    s2 = fmaf(g[c], xh[c], s2);
  }
  if (act) { red[r * 4 + sl] = s1; red[kRows * 4 + r * 4 + sl] = s2; }
  __syncthreads();
  s1 = s2 = 0.f;
#pragma unroll
  for (int k = 0; k < NS; ++k) { s1 += red[r * 4 + k]; s2 += red[kRows * 4 + r * 4 + k]; }
  s1 *= 1.0f / (float)E;
  s2 *= 1.0f / (float)E;
#pragma unroll
  for (int c = 0; c < 16; ++c) dz[c] = rs * (g[c] - s1 - xh[c] * s2);
  float pv[32];
#pragma unroll
  for (int c = 0; c < 16; ++c) { pv[c] = dyv[c] * xh[c]; pv[16 + c] = dyv[c]; }
  const float tot = warp_transpose_reduce(pv, lane);
  if (act) atomicAdd(acc + (lane >> 4) * E + sl * 16 + (lane & 15), tot);
}

template <int E, int HD, int NH, int SMAX>
__global__ void __launch_bounds__(kThreads, 1) encoder_layer_bwd_kernel(const __grid_constant__ BwdArgs a, const __grid_constant__ BwdMaps tm) {
  constexpr int D = E / NH;                 // head dim (16)
  constexpr int CH = E / 4;                 // 16-byte chunks of an E-wide row
  constexpr int NS = E / 16;                // active column slices of an E-wide epilogue
  constexpr int RPL = SMAX / 32;            // sequence rows per lane in the attention passes
  constexpr int TILE = kRows * E * 4;       // bytes of an E-wide tile (swizzled row-major, or a K-major slab)
  constexpr int WA_BYTES = (3 * E * E > E * HD + E * E ? 3 * E * E : E * HD + E * E) * 4;
  constexpr int WB_BYTES = E * HD * 4;
  constexpr int C1 = 0, C2 = HD, C3 = HD + E, C4 = 0;   // TMEM columns of the four accumulators
  static_assert(HD == 2 * E && D == 16 && NS <= 4 && (CH & (CH - 1)) == 0 && CH >= 8, "tile shapes");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Wa = smem;                         // [W_2^T | W_o^T], then [W_q^T | W_k^T | W_v^T]
  uint8_t* Wb = Wa + WA_BYTES;                // W_1^T (resident over the CTA's tiles)
  uint8_t* R0 = Wb + WB_BYTES;                // dz2 slab -> dz1 slab -> q tile (dk on exit) -> dv slab
  uint8_t* R1 = R0 + TILE;                    // dh2 slab (R1, R2) -> k tile (dq on exit) -> dk slab
  uint8_t* R2 = R1 + TILE;                    //                   -> v tile (dv on exit) -> dx staging
  uint8_t* R3 = R2 + TILE;                    // staging tile -> dctx tile -> dq slab
  float* red = reinterpret_cast<float*>(R3 + TILE);      // [2][128][4] LayerNorm partial sums
  float* lnacc = red + 2 * kRows * 4;                    // [LN1 | LN2][dgamma | dbeta][E]
  float* lsdl = lnacc + 4 * E;                           // [16 warps][L | D][SMAX]
  uint64_t* bar = reinterpret_cast<uint64_t*>(lsdl + 16 * 2 * SMAX);   // MMA completion
  uint64_t* lbar = bar + 1;              // [7] TMA completion: W_2^T + W_o^T | W_1^T (once) | W_qkv^T | dq, dk, dv slabs | k, v | q | dy, z2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lbar + 7);
  const int tid = threadIdx.x, warp = umma::warp_idx_uniform(), lane = tid & 31;
  const int qd = warp & 3, sl = warp >> 2;
  const int r = qd * 32 + lane;               // token row of this thread in every epilogue
  const bool act = sl < NS;
  const int S = a.S;

  if (tid == 0) {
    umma::mbar_init(bar, 1);
    for (int i = 0; i < 7; ++i) umma::mbar_init(lbar + i, 1);
    tma::prefetch_map(&tm.dy); tma::prefetch_map(&tm.z2);
    umma::mbar_fence_init();
    tma::prefetch_map(&tm.q); tma::prefetch_map(&tm.k); tma::prefetch_map(&tm.v);
    tma::prefetch_map(&tm.dz2); tma::prefetch_map(&tm.dh2); tma::prefetch_map(&tm.dx);
    tma::prefetch_map(&tm.w2t); tma::prefetch_map(&tm.wot); tma::prefetch_map(&tm.w1t); tma::prefetch_map(&tm.wqkvt);
    tma::prefetch_map(&tm.dq); tma::prefetch_map(&tm.dk); tma::prefetch_map(&tm.dv);
  }
  if (warp == 0) umma::tmem_alloc<HD + 2 * E>(tmem_slot);
  for (int i = tid; i < 4 * E; i += kThreads) lnacc[i] = 0.f;
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(qd * 32) << 16);
  uint32_t parity = 0, lpar = 0;   // lpar: phase of the per-tile load barriers (each completes once per tile)
  const float scale = rsqrtf((float)D);
  bool wa_ready = false;       // [W_2^T | W_o^T] of the next tile already requested
  bool w1_waited = false;      // W_1^T is loaded once and stays resident
  // one elected thread issues every TMA load; the MMA warp waits for them
  auto load_wa_first = [&]() {   // [W_2^T | W_o^T] -> Wa
    tma::expect_tx(lbar, (uint32_t)((E * HD + E * E) * 4));
#pragma unroll
    for (int kb = 0; kb < E / 32; ++kb) {
      tma::load_tile(Wa + (size_t)kb * HD * 128, &tm.w2t, kb * 32, 0, lbar);
      tma::load_tile(Wa + E * HD * 4 + (size_t)kb * E * 128, &tm.wot, kb * 32, 0, lbar);
    }
  };
  auto load_in_first = [&](long long rows0) {   // dy -> R3 (the staging tile's place), z2 -> R1: both free between two tiles
    tma::expect_tx(lbar + 6, (uint32_t)(2 * TILE));
#pragma unroll
    for (int kb = 0; kb < E / 32; ++kb) {
      tma::load_tile(R3 + (size_t)kb * kRows * 128, &tm.dy, kb * 32, (int)rows0, lbar + 6);
      tma::load_tile(R1 + (size_t)kb * kRows * 128, &tm.z2, kb * 32, (int)rows0, lbar + 6);
    }
  };
  if (tid == 0 && (int)blockIdx.x < a.n_tiles) {
    tma::expect_tx(lbar + 1, (uint32_t)(E * HD * 4));
#pragma unroll
    for (int kb = 0; kb < HD / 32; ++kb) tma::load_tile(Wb + (size_t)kb * E * 128, &tm.w1t, kb * 32, 0, lbar + 1);
  }
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int seq0 = tile * a.spt;
    const int nseq = min(a.spt, a.B - seq0);
    const long long row0 = (long long)seq0 * S;
    const int nrows = nseq * S;
    const bool live = act && r < nrows;
    if (!wa_ready && tid == 0) {
      load_wa_first();
      load_in_first(row0);
    }
    // ---- dz2 = LN2'(dy)
    float dz2v[16];
    {
      float dyv[16], zv[16];
      float m = 0.f, rs = 0.f;
      if (live) {
        m = __ldg(a.m2 + row0 + r);
        rs = __ldg(a.r2 + row0 + r);
      }
      umma::mbar_wait(lbar + 6, lpar);   // dy and z2 tiles (TMA, requested during the previous tile)
      if (live) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {   // this thread's own 64 bytes of both tiles (it overwrites exactly these with dz2 below)
          const float4 d4 = rm_ld<CH>(R3, r, sl * 4 + c), z4 = rm_ld<CH>(R1, r, sl * 4 + c);
          dyv[4 * c] = d4.x; dyv[4 * c + 1] = d4.y; dyv[4 * c + 2] = d4.z; dyv[4 * c + 3] = d4.w;
          zv[4 * c] = z4.x; zv[4 * c + 1] = z4.y; zv[4 * c + 2] = z4.z; zv[4 * c + 3] = z4.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) dyv[c] = zv[c] = 0.f;
      }
      layernorm_bwd_sliced<E>(dyv, zv, m, rs, act, r, sl, lane, red, a.g2, lnacc + 2 * E, dz2v);
      if (act) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 t = make_float4(dz2v[4 * c], dz2v[4 * c + 1], dz2v[4 * c + 2], dz2v[4 * c + 3]);
          *reinterpret_cast<float4*>(R0 + ((size_t)(sl * 4 + c) * kRows + r) * 16) = t;
          rm_st<CH>(R3, r, sl * 4 + c, t);
        }
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- dh = dz2 W_2 ; dz2 to HBM meanwhile
    if (warp == 4) {
      umma::mbar_wait(lbar, lpar);       // [W_2^T | W_o^T] have landed
      umma::fence_after_sync();
      if (umma::elect_one()) {
        issue_chain<false>(tmem + C1, R0, Wa, HD, E, true);
        umma::commit(bar);
      }
      __syncwarp();
    }
    if (tid == 0) {                // dz2 to HBM: TMA store of the staging tile (boxes of the rows this tile owns)
#pragma unroll
      for (int kb = 0; kb < E / 32; ++kb) tma_store_box(&tm.dz2, R3 + (size_t)kb * kRows * 128, kb * 32, (int)row0);
      bulk_commit();
    }
    // the ReLU gate (sign of the stored relu(h)) of this thread's columns of both hidden chunks: in flight during the MMA
    float4 gate[HD / E][4];
#pragma unroll
    for (int chunk = 0; chunk < HD / E; ++chunk)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        gate[chunk][c] = live ? __ldg(reinterpret_cast<const float4*>(a.hact + (row0 + r) * HD + chunk * E + sl * 16) + c)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
    umma::mbar_wait(bar, parity); parity ^= 1;
    umma::fence_after_sync();
    if (tid == 0) bulk_wait_read0();
    __syncthreads();               // staging tile consumed (the TMA store has read it); the dz2 slab (R0) is consumed by the MMA
    static_assert(HD / E == 2, "two hidden chunks: staged in R3 and R0");
#pragma unroll
    for (int chunk = 0; chunk < HD / E; ++chunk) {          // E hidden columns at a time: ReLU gate, A operand, staging tile
      if (act) {
        const int c0 = chunk * E + sl * 16;
        float v[16];
        umma::tmem_ld16(trow + (uint32_t)(C1 + c0), v);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 h4 = gate[chunk][c];
          const float4 t = make_float4(h4.x > 0.f ? v[4 * c] : 0.f, h4.y > 0.f ? v[4 * c + 1] : 0.f, h4.z > 0.f ? v[4 * c + 2] : 0.f,
                                       h4.w > 0.f ? v[4 * c + 3] : 0.f);
          *reinterpret_cast<float4*>(R1 + ((size_t)(c0 / 4 + c) * kRows + r) * 16) = t;
          rm_st<CH>(chunk == 0 ? R3 : R0, r, sl * 4 + c, t);
        }
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0) {                // dh2 to HBM (the weight gradient of fc1 reads it there)
#pragma unroll
      for (int kb = 0; kb < E / 32; ++kb) {
        tma_store_box(&tm.dh2, R3 + (size_t)kb * kRows * 128, kb * 32, (int)row0);
        tma_store_box(&tm.dh2, R0 + (size_t)kb * kRows * 128, E + kb * 32, (int)row0);
      }
      bulk_commit();
    }
    // ---- d(x1) = dz2 + dh2 W_1 ; dz1 = LN1'(d(x1))
    if (warp == 4) {
      if (!w1_waited) umma::mbar_wait(lbar + 1, 0);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        issue_chain<false>(tmem + C2, R1, Wb, E, HD, true);
        umma::commit(bar);
      }
      __syncwarp();
    }
    float dz1v[16];
    {
      float zv[16];
      float m = 0.f, rs = 0.f;
      if (live) {
        const float4* zp = reinterpret_cast<const float4*>(a.z1 + (row0 + r) * E + sl * 16);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 z4 = __ldg(zp + c);
          zv[4 * c] = z4.x; zv[4 * c + 1] = z4.y; zv[4 * c + 2] = z4.z; zv[4 * c + 3] = z4.w;
        }
        m = __ldg(a.m1 + row0 + r);
        rs = __ldg(a.r1 + row0 + r);
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) zv[c] = 0.f;
      }
      umma::mbar_wait(bar, parity); parity ^= 1;
      umma::fence_after_sync();
      // the dh2 slab is consumed: k and v tiles of the attention backward land in its place
      if (tid == 0) {
        bulk_wait_read0();         // the dh2 staging tiles (R3, R0) have been read: LayerNorm 1' below rewrites them
        tma::expect_tx(lbar + 4, (uint32_t)(2 * TILE));
#pragma unroll
        for (int kb = 0; kb < E / 32; ++kb) {
          tma::load_tile(R1 + (size_t)kb * kRows * 128, &tm.k, kb * 32, (int)row0, lbar + 4);
          tma::load_tile(R2 + (size_t)kb * kRows * 128, &tm.v, kb * 32, (int)row0, lbar + 4);
        }
      }
      float dyv[16];
      if (act) {
        umma::tmem_ld16(trow + (uint32_t)(C2 + sl * 16), dyv);
#pragma unroll
        for (int c = 0; c < 16; ++c) dyv[c] = live ? dyv[c] + dz2v[c] : 0.f;
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) dyv[c] = 0.f;
      }
      layernorm_bwd_sliced<E>(dyv, zv, m, rs, act, r, sl, lane, red, a.g1, lnacc, dz1v);
      if (act) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 t = make_float4(dz1v[4 * c], dz1v[4 * c + 1], dz1v[4 * c + 2], dz1v[4 * c + 3]);
          *reinterpret_cast<float4*>(R0 + ((size_t)(sl * 4 + c) * kRows + r) * 16) = t;
          rm_st<CH>(R3, r, sl * 4 + c, t);
        }
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- dctx = dz1 W_o ; dz1 to HBM meanwhile
    if (warp == 4) {
      umma::fence_after_sync();
      if (umma::elect_one()) {
        issue_chain<false>(tmem + C3, R0, Wa + E * HD * 4, E, E, true);
        umma::commit(bar);
      }
      __syncwarp();
    }
    copy_out<CH>(R3, a.dz1, row0, nrows, E, 0, tid);
    umma::mbar_wait(bar, parity); parity ^= 1;
    umma::fence_after_sync();
    __syncthreads();               // staging consumed; the dz1 slab and [W_2^T | W_o^T] are consumed too
    if (tid == 0) {                // q tile -> R0 and W_qkv^T -> Wa, in flight during the dctx epilogue / the attention
      tma::expect_tx(lbar + 5, (uint32_t)TILE);
#pragma unroll
      for (int kb = 0; kb < E / 32; ++kb) tma::load_tile(R0 + (size_t)kb * kRows * 128, &tm.q, kb * 32, (int)row0, lbar + 5);
      tma::expect_tx(lbar + 2, (uint32_t)(3 * E * E * 4));
#pragma unroll
      for (int kb = 0; kb < 3 * E / 32; ++kb) tma::load_tile(Wa + (size_t)kb * E * 128, &tm.wqkvt, kb * 32, 0, lbar + 2);
    }
    if (act) {                     // dctx -> swizzled row-major tile
      float v[16];
      umma::tmem_ld16(trow + (uint32_t)(C3 + sl * 16), v);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        rm_st<CH>(R3, r, sl * 4 + c, live ? make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    umma::mbar_wait(lbar + 4, lpar);   // k, v tiles
    umma::mbar_wait(lbar + 5, lpar);   // q tile
    umma::fence_before_sync();
    __syncthreads();
    if constexpr (SMAX == 32) {
      // ---- attention backward on warp-level tensor-core MMAs: one (sequence, head) per warp (attention_bwd_mma above)
      float* Ls = lsdl + warp * 2 * SMAX;
      float* Dl = Ls + SMAX;
      for (int pair = warp; pair < nseq * NH; pair += kThreads / 32) {
        const int sq = pair / NH, h = pair % NH, rb = sq * S;
        Ls[lane] = lane < S ? __ldg(a.lse + ((size_t)(seq0 + sq) * NH + h) * S + lane) * 1.4426950408889634f : 0.f;   // base 2
        Dl[lane] = 0.f;
        __syncwarp();
        attention_bwd_mma<E, CH>(R0, R1, R2, R3, Ls, Dl, rb, h * D, S, scale, lane, a.dq + row0 * E, a.dk + row0 * E, a.dv + row0 * E);
        __syncwarp();
      }
      asm volatile("fence.proxy.async;" ::: "memory");   // this thread's dq / dk / dv stores (generic proxy) before the TMA reads them
      __syncthreads();             // dq / dk / dv of the tile are in HBM / L2; every tile is dead
      if (tid == 0) {              // read them back as K-major slabs (rows past the tile are other tiles' rows: never stored)
        tma::expect_tx(lbar + 3, (uint32_t)(3 * TILE));
#pragma unroll
        for (int kb = 0; kb < E / 32; ++kb) {
          tma::load_tile(R3 + (size_t)kb * kRows * 128, &tm.dq, kb * 32, (int)row0, lbar + 3);
          tma::load_tile(R1 + (size_t)kb * kRows * 128, &tm.dk, kb * 32, (int)row0, lbar + 3);
          tma::load_tile(R0 + (size_t)kb * kRows * 128, &tm.dv, kb * 32, (int)row0, lbar + 3);
        }
      }
    } else {
      // ---- attention backward: one (sequence, head) per warp.  Results replace dead operands of the same (sequence, head)
      //      block: dq -> k tile, dk -> q tile, dv -> v tile.
      {
        float* Ls = lsdl + warp * 2 * SMAX;
        float* Dl = Ls + SMAX;
        for (int pair = warp; pair < nseq * NH; pair += kThreads / 32) {
          const int sq = pair / NH, h = pair % NH, c0 = h * (D / 4), rb = sq * S;
          // pass A (lane = query row i): D_i = dctx_i . ctx_i, dQ_i = scale * sum_j dS_ij K_j
          float dqv[RPL][D];
#pragma unroll
          for (int rr = 0; rr < RPL; ++rr) {
            const int i = lane + 32 * rr;
            if (i < S) {
              float gi[D], qi[D];
              ld_head<CH, D>(R3, rb + i, c0, gi, 1.f);
              ld_head<CH, D>(R0, rb + i, c0, qi, scale);
              float di = 0.f;
              const float4* op = reinterpret_cast<const float4*>(a.ctx + (row0 + rb + i) * E + h * D);
#pragma unroll
              for (int c = 0; c < D / 4; ++c) {
                const float4 o4 = __ldg(op + c);
                di = fmaf(gi[4 * c], o4.x, di); di = fmaf(gi[4 * c + 1], o4.y, di);
                di = fmaf(gi[4 * c + 2], o4.z, di); di = fmaf(gi[4 * c + 3], o4.w, di);
              }
              const float li = __ldg(a.lse + ((size_t)(seq0 + sq) * NH + h) * S + i);
              Ls[i] = li;
              Dl[i] = di;
              float acc[D];
#pragma unroll
              for (int c = 0; c < D; ++c) acc[c] = 0.f;
              for (int j = 0; j < S; ++j) {
                float kj[D], vj[D];
                ld_head<CH, D>(R1, rb + j, c0, kj, 1.f);
                ld_head<CH, D>(R2, rb + j, c0, vj, 1.f);
                float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f;
#pragma unroll
                for (int c = 0; c < D; c += 2) {
                  s0 = fmaf(qi[c], kj[c], s0); s1 = fmaf(qi[c + 1], kj[c + 1], s1);
                  p0 = fmaf(gi[c], vj[c], p0); p1 = fmaf(gi[c + 1], vj[c + 1], p1);
                }
                const float p = __expf((s0 + s1) - li);
                const float ds = p * ((p0 + p1) - di);
#pragma unroll
                for (int c = 0; c < D; ++c) acc[c] = fmaf(ds, kj[c], acc[c]);
              }
#pragma unroll
              for (int c = 0; c < D; ++c) dqv[rr][c] = acc[c] * scale;
            }
          }
          __syncwarp();
          // pass B (lane = key row j): dK_j = scale * sum_i dS_ij Q_i,  dV_j = sum_i P_ij dctx_i
          float kjs[RPL][D], vjs[RPL][D];
#pragma unroll
          for (int rr = 0; rr < RPL; ++rr) {
            const int j = lane + 32 * rr;
            if (j < S) {
              ld_head<CH, D>(R1, rb + j, c0, kjs[rr], scale);
              ld_head<CH, D>(R2, rb + j, c0, vjs[rr], 1.f);
            }
          }
          __syncwarp();              // every lane holds its k / v rows: the k block is dead, dq moves in
#pragma unroll
          for (int rr = 0; rr < RPL; ++rr) {
            const int i = lane + 32 * rr;
            if (i < S) st_head<CH, D>(R1, rb + i, c0, dqv[rr], 1.f);
          }
          float dkv[RPL][D];
#pragma unroll
          for (int rr = 0; rr < RPL; ++rr) {
            const int j = lane + 32 * rr;
            if (j < S) {
              float ak[D], av[D];
#pragma unroll
              for (int c = 0; c < D; ++c) ak[c] = av[c] = 0.f;
              for (int i = 0; i < S; ++i) {
                float qi[D], gi[D];
                ld_head<CH, D>(R0, rb + i, c0, qi, 1.f);
                ld_head<CH, D>(R3, rb + i, c0, gi, 1.f);
                float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f;
#pragma unroll
                for (int c = 0; c < D; c += 2) {
                  s0 = fmaf(kjs[rr][c], qi[c], s0); s1 = fmaf(kjs[rr][c + 1], qi[c + 1], s1);
                  p0 = fmaf(vjs[rr][c], gi[c], p0); p1 = fmaf(vjs[rr][c + 1], gi[c + 1], p1);
                }
                const float p = __expf((s0 + s1) - Ls[i]);
                const float ds = p * ((p0 + p1) - Dl[i]);
#pragma unroll
                for (int c = 0; c < D; ++c) { ak[c] = fmaf(ds, qi[c], ak[c]); av[c] = fmaf(p, gi[c], av[c]); }
              }
              st_head<CH, D>(R2, rb + j, c0, av, 1.f);      // the v block is dead since the loads above
#pragma unroll
              for (int c = 0; c < D; ++c) dkv[rr][c] = ak[c] * scale;
            }
          }
          __syncwarp();              // every lane is done with the q block: dk moves in
#pragma unroll
          for (int rr = 0; rr < RPL; ++rr) {
            const int j = lane + 32 * rr;
            if (j < S) st_head<CH, D>(R0, rb + j, c0, dkv[rr], 1.f);
          }
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- dq / dk / dv to HBM and into K-major slabs (each slab replaces a tile that has just been consumed)
      copy_out<CH>(R1, a.dq, row0, nrows, E, 0, tid);
      tile_to_slab<CH>(R1, R3, tid);          // dq slab <- dctx tile's place
      __syncthreads();
      copy_out<CH>(R0, a.dk, row0, nrows, E, 0, tid);
      tile_to_slab<CH>(R0, R1, tid);          // dk slab <- dq tile's place
      __syncthreads();
      copy_out<CH>(R2, a.dv, row0, nrows, E, 0, tid);
      tile_to_slab<CH>(R2, R0, tid);          // dv slab <- dk tile's place
      umma::fence_proxy_async();
      umma::fence_before_sync();
      __syncthreads();
    }
    // ---- dx = dz1 + dq W_q + dk W_k + dv W_v
    if (warp == 4) {
      umma::mbar_wait(lbar + 2, lpar);                       // W_qkv^T
      if (SMAX == 32) umma::mbar_wait(lbar + 3, lpar);       // the dq / dk / dv slabs
      umma::fence_after_sync();
      if (umma::elect_one()) {
        constexpr size_t WSTEP = (size_t)(E / 32) * E * 128;  // boxes of one of the three matrices
        issue_chain<SMAX == 32>(tmem + C4, R3, Wa, E, E, true);
        issue_chain<SMAX == 32>(tmem + C4, R1, Wa + WSTEP, E, E, false);
        issue_chain<SMAX == 32>(tmem + C4, R0, Wa + 2 * WSTEP, E, E, false);
        umma::commit(bar);
      }
      __syncwarp();
    }
    // the residual dz1: this CTA wrote the rows above (L2); re-reading them frees 16 registers across the attention phase, and the
    // loads are in flight during the last MMA chain
    float4 zres[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      zres[c] = live ? __ldcg(reinterpret_cast<const float4*>(a.dz1 + (row0 + r) * E + sl * 16) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    umma::mbar_wait(bar, parity); parity ^= 1;
    umma::fence_after_sync();
    w1_waited = true;
    lpar ^= 1;
    wa_ready = tile + (int)gridDim.x < a.n_tiles;
    if (wa_ready && tid == 0) {   // the next tile's first weights and LayerNorm inputs, in flight during this epilogue
      load_wa_first();
      load_in_first((long long)(tile + (int)gridDim.x) * a.spt * S);
    }
    if (act) {
      float v[16];
      umma::tmem_ld16(trow + (uint32_t)(C4 + sl * 16), v);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 z4 = zres[c];
        rm_st<CH>(R2, r, sl * 4 + c, make_float4(v[4 * c] + z4.x, v[4 * c + 1] + z4.y, v[4 * c + 2] + z4.z, v[4 * c + 3] + z4.w));
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();               // the next tile overwrites every buffer; R2 is next written by a TMA load issued by thread 0
    if (tid == 0) {                // dx to HBM
#pragma unroll
      for (int kb = 0; kb < E / 32; ++kb) tma_store_box(&tm.dx, R2 + (size_t)kb * kRows * 128, kb * 32, (int)row0);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait0();
  umma::fence_before_sync();
  __syncthreads();
  for (int i = tid; i < 4 * E; i += kThreads) {
    float* dst = i < E ? a.dg1 + i : i < 2 * E ? a.db1 + (i - E) : i < 3 * E ? a.dg2 + (i - 2 * E) : a.db2 + (i - 3 * E);
    atomicAdd(dst, lnacc[i]);
  }
  if (warp == 0) umma::tmem_dealloc<HD + 2 * E>(tmem);
}

// transposed weights of every layer, one launch: out[l] = [W_2^T (HD x E) | W_1^T (E x HD) | W_o^T (E x E) | W_qkv^T (E x 3E)]
struct PackArgs {
  long long off[kEncMaxLayers][6];   // offsets of q_w, k_w, v_w, o_w, f1_w, f2_w in the flat parameter buffer
};
__global__ void pack_kernel(const float* __restrict__ params, const __grid_constant__ PackArgs pa, float* __restrict__ out, int E, int HD) {
  const int l = blockIdx.y;
  const size_t per = (size_t)2 * E * HD + 4 * (size_t)E * E;
  float* o = out + l * per;
  const float* wq = params + pa.off[l][0];
  const float* wk = params + pa.off[l][1];
  const float* wv = params + pa.off[l][2];
  const float* wo = params + pa.off[l][3];
  const float* w1 = params + pa.off[l][4];   // [HD][E]
  const float* w2 = params + pa.off[l][5];   // [E][HD]
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int)per; i += gridDim.x * blockDim.x) {
    int j = i;
    float v;
    if (j < E * HD) {                        // W_2^T [h][e]
      const int h = j / E, e = j % E;
      v = w2[(size_t)e * HD + h];
    } else if ((j -= E * HD) < E * HD) {     // W_1^T [e][h]
      const int e = j / HD, h = j % HD;
      v = w1[(size_t)h * E + e];
    } else if ((j -= E * HD) < E * E) {      // W_o^T [i][o]
      const int ii = j / E, oo = j % E;
      v = wo[(size_t)oo * E + ii];
    } else {                                 // [W_q^T | W_k^T | W_v^T]  [i][w * E + o]
      j -= E * E;
      const int ii = j / (3 * E), rem = j % (3 * E), w = rem / E, oo = rem % E;
      const float* ww = w == 0 ? wq : w == 1 ? wk : wv;
      v = ww[(size_t)oo * E + ii];
    }
    o[i] = v;
  }
}

int sm_count() {
  static int sms = 0;
  if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  return sms ? sms : 148;
}

template <int E, int HD, int NH, int SMAX>
int launch_bwd(const BwdArgs& a, cudaStream_t st) {
  constexpr int WA = (3 * E * E > E * HD + E * E ? 3 * E * E : E * HD + E * E) * 4, WB = E * HD * 4;
  const int smem = WA + WB + 4 * kRows * E * 4 + 2 * kRows * 4 * 4 + 4 * E * 4 + 16 * 2 * SMAX * 4 + 128;
  auto kern = encoder_layer_bwd_kernel<E, HD, NH, SMAX>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = a.n_tiles < sm_count() ? a.n_tiles : sm_count();
  const double T = (double)a.B * a.S;
  BwdMaps tm;
  {
    const long long rows = (long long)a.B * a.S;
    int rc = make_f32_tensor_map_sw(&tm.w2t, a.w2t, E, HD, HD);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.wot, a.wot, E, E, E);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.w1t, a.w1t, HD, E, E);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.wqkvt, a.wqkvt, 3 * E, E, E);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.dq, a.dq, E, rows, kRows);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.dk, a.dk, E, rows, kRows);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.dv, a.dv, E, rows, kRows);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.q, a.q, E, rows, kRows);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.k, a.k, E, rows, kRows);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.v, a.v, E, rows, kRows);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.dy, a.dy, E, rows, kRows);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.z2, a.z2, E, rows, kRows);
    const int own = a.spt * a.S;   // rows a tile owns: the store boxes (the last tile is clipped at the end of the tensor)
    if (!rc) rc = make_f32_tensor_map_sw(&tm.dz2, a.dz2, E, rows, own);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.dh2, a.dh2, HD, rows, own);
    if (!rc) rc = make_f32_tensor_map_sw(&tm.dx, a.dx, E, rows, own);
    if (rc) return rc;
  }
  MivitProfScope prof("encoder_layer_bwd", 2.0 * T * (4.0 * E * E + 2.0 * E * HD) + 10.0 * a.B * NH * (double)a.S * a.S * (E / NH), st);
  kern<<<grid, kThreads, smem, st>>>(a, tm);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

}  // namespace

size_t encoder_bwd_packed_floats(int E, int HD) { return (size_t)2 * E * HD + 4 * (size_t)E * E; }

int encoder_pack_bwd_weights(const float* params, const long long (*offs)[6], int L, int E, int HD, float* out, cudaStream_t st) {
  MIVIT_CHECK_ARG(L >= 1 && L <= kEncMaxLayers, "fused encoder backward: more than %d layers", kEncMaxLayers);
  PackArgs pa;
  for (int l = 0; l < L; ++l)
    for (int i = 0; i < 6; ++i) pa.off[l][i] = offs[l][i];
  pack_kernel<<<dim3(8, L), 256, 0, st>>>(params, pa, out, E, HD);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

int encoder_layer_bwd(const EncoderLayerBwdIO& io, int B, int S, int E, int HD, int H, cudaStream_t st) {
  MIVIT_CHECK_ARG(encoder_fused_supported(B, S, E, HD, H), "fused encoder layer: unsupported shape");
  BwdArgs a;
  a.dy = io.dy; a.dx = io.dx;
  a.z2 = io.z2; a.m2 = io.m2; a.r2 = io.r2; a.hact = io.hact; a.z1 = io.z1; a.m1 = io.m1; a.r1 = io.r1;
  a.q = io.q; a.k = io.k; a.v = io.v; a.lse = io.lse; a.ctx = io.ctx;
  a.g2 = io.g2; a.g1 = io.g1;
  a.w2t = io.packed;
  a.w1t = a.w2t + (size_t)E * HD;
  a.wot = a.w1t + (size_t)E * HD;
  a.wqkvt = a.wot + (size_t)E * E;
  a.dz2 = io.dz2; a.dh2 = io.dh2; a.dz1 = io.dz1; a.dq = io.dq; a.dk = io.dk; a.dv = io.dv;
  a.dg2 = io.dg2; a.db2 = io.db2; a.dg1 = io.dg1; a.db1 = io.db1;
  a.B = B; a.S = S; a.spt = kRows / S; a.n_tiles = (B + a.spt - 1) / a.spt;
  if (E == 64) return S <= 32 ? launch_bwd<64, 128, 4, 32>(a, st) : launch_bwd<64, 128, 4, 64>(a, st);
  return S <= 32 ? launch_bwd<32, 64, 2, 32>(a, st) : launch_bwd<32, 64, 2, 64>(a, st);
}
