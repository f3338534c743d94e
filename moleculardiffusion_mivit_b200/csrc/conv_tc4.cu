// CTA-pair (tcgen05 cta_group::2) version of the TMA-fed shifted-row convolution for C_out = 128
// (64 -> 128 [+ fused 1x1 skip], 128 -> 128; forward and input gradient).  Layout / algorithm: conv_tc.cu, conv_tc3.cu.
//
// Why: with the output columns split over two independent CTAs (conv_tc3.cu, N = 64 per MMA) every 128 x 64 x 16 MMA
// re-reads its 4 KB A operand and 2 KB B operand from shared memory for 32 cycles of tensor work -- 192 B/cycle against
// the ~128 B/cycle the SM can feed -- and ncu shows the issuing warps back-pressured at 54 % tensor-pipe activity
// (profiles/r01_ncu_full_summary.md).  A CTA pair issues ONE 256 x 128 x 16 MMA: each SM streams its own 128 rows of A
// (4 KB) and its HALF of the weights (64 output columns, 2 KB) for 64 cycles of tensor work (96 B/cycle) and the hardware
// exchanges the B halves between the two SMs.  Each CTA of the pair owns a different 128-row tile and all 128 output
// columns of it, so a row slab is also loaded exactly once (no multicast needed).
//
// Protocol (cluster = 2 CTAs, rank 0 = leader):
//   producer (each CTA): waits its own empty[slot], TMA-loads ITS tile's unit into ITS shared memory, completion bytes
//                        go to the LEADER's full[slot] (the leader expects 2 x unit bytes);
//   MMA issuers (leader only, warps 4/5 alternate pair-tiles): wait full[slot] (+ tempty[buf], 8 arrivals = 4 epilogue
//                        warps of both CTAs), issue cta_group::2 MMAs, commit with multicast to empty[slot] / tfull[buf]
//                        of BOTH CTAs;
//   epilogue (each CTA): waits its tfull[buf], reads its own TMEM (its 128 rows x 128 columns), stages / stores /
//                        accumulates statistics like conv_tc3, then arrives on the LEADER's tempty[buf].
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kTileM = 128;
constexpr int kBoxRows = 32;
constexpr int kMaxRing = 8;

__device__ __forceinline__ void epi_bar_sync4() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void tma_store_tile4(const CUtensorMap* tm, const void* smem_src, int ch0, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(ch0), "r"(row),
               "r"(umma::smem_u32(smem_src))
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read4() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// Relaxed: the arrival only hands TMEM columns back to the leader's MMA warps, and tcgen05.fence::before_thread_sync /
// after_thread_sync order the tensor-memory accesses on both sides.  (.release compiled to MEMBAR.ALL.CTA, which also waits
// for every global load / store the warp has in flight -- with the BatchNorm-backward epilogue that exposed the full DRAM
// latency of the raw-tile prefetch once per tile: 1.03 -> 1.53 ms per launch, ncu source page of r01_bstat.)
__device__ __forceinline__ void arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void load_tile_2sm(void* smem_dst, const CUtensorMap* tm, int ch0, int row, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          umma::smem_u32(smem_dst)),
      "l"(tm), "r"(ch0), "r"(row), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit_2sm(uint64_t* bar) {   // arrives on `bar` of BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   umma::smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma::smem_u32(smem_dst)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// NT = C_out = accumulator columns per CTA (128 or 64); NS = NT / 2 = weight columns resident per CTA
template <int CIN, bool SKIP, int TAPS, int NT>
struct Cfg4 {
  static constexpr int NS = NT / 2;
  static constexpr int kCH = CIN / 8;
  static constexpr int kWBytes = TAPS * CIN * NS * 2;
  static constexpr int kWSkipBytes = SKIP ? CIN * NS * 2 : 0;
  static constexpr int kStageBufs = kWBytes + kWSkipBytes > 128 * 1024 ? 1 : 2;
  static constexpr int kRegions = CIN / 64;
  static constexpr int kAccCols = NT * (SKIP ? 2 : 1);
  static constexpr int kTmemCols = 2 * kAccCols <= 128 ? 128 : 2 * kAccCols <= 256 ? 256 : 512;
  static constexpr int kOutPerAcc = NT / 64;
  static constexpr int kOutRegions = kOutPerAcc * (SKIP ? 2 : 1);
  static constexpr int kStageBytes = kTileM * 128;
};

// BSTAT (input-gradient launches): the output Y is the gradient g of a ReLU(BatchNorm(raw)) activation, and instead of
// (sum y, sum y^2) the epilogue accumulates that BatchNorm's backward sums over the masked gradient,
//   stats[0][c] = sum_r m*g,  stats[1][c] = sum_r m*g*raw,   m = 1[raw*scale + shift > 0]   (bn_ss = scale | shift),
// from the staged (bf16-rounded, i.e. exactly the stored) g and a raw tile prefetched into registers while the MMAs of the
// tile are still running.  This removes bn.cu's reduction pass over g and raw (2 x 1.5 GB at 128 channels).
template <int CIN, bool SKIP, int TAPS, int NT, bool BSTAT>
__global__ void __launch_bounds__(256, 1)
conv_rows_tc4_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                     const __grid_constant__ CUtensorMap tmYsk, const __nv_bfloat16* __restrict__ Wp,
                     const __nv_bfloat16* __restrict__ Wsk, float* __restrict__ stats, float* __restrict__ stats_sk, long long rows,
                     int n_tiles, int P, ConvShifts shifts, int halo, int xslab_rows, int ring, int guard,
                     const __nv_bfloat16* __restrict__ bn_raw, const float* __restrict__ bn_ss, uint4* __restrict__ y_out) {
  static_assert(!(BSTAT && SKIP), "BatchNorm-backward sums are for single-output launches");
  using C = Cfg4<CIN, SKIP, TAPS, NT>;
  constexpr int taps = TAPS;
  constexpr int NS = C::NS;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_uniform(), lane = tid & 31;
  const int unit_bytes = xslab_rows * 128;
  uint8_t* wsm = smem;                                   // [taps][CH][64][8]   this CTA's half of the output columns
  uint8_t* wsk = wsm + C::kWBytes;                       // [CH][64][8]         (SKIP)
  uint8_t* stage0 = wsk + C::kWSkipBytes;
  uint8_t* slab0 = stage0 + C::kStageBufs * C::kStageBytes;
  uint8_t* tail = slab0 + (size_t)ring * unit_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);    // [kMaxRing] used in the leader only
  uint64_t* empty = full + kMaxRing;                     // [kMaxRing]
  uint64_t* tfull = empty + kMaxRing;                    // [2]
  uint64_t* tempty = tfull + 2;                          // [2] used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* red = reinterpret_cast<float*>(tmem_slot + 4);  // [kOutRegions][2][64]

  const uint32_t rank = tma::cluster_ctarank();          // 0 = leader
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_pair_tiles = (n_tiles + 1) >> 1;
  const int col0 = (int)rank * NS;                       // resident weight columns of this CTA

  if (tid == 0) {
    for (int i = 0; i < kMaxRing; ++i) {
      umma::mbar_init(full + i, 1);
      umma::mbar_init(empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(tfull + i, 1);
      umma::mbar_init(tempty + i, 8);
    }
    umma::mbar_fence_init();
    tma::prefetch_map(&tmX);
    tma::prefetch_map(&tmY);
    if (SKIP) tma::prefetch_map(&tmYsk);
  }
  if (warp == 0) tmem_alloc_2sm<C::kTmemCols>(tmem_slot);
  for (int i = tid; i < C::kOutRegions * 2 * 64; i += 256) red[i] = 0.f;
  {
    const int per_tap = C::kCH * NS;
    for (int i = tid; i < taps * per_tap; i += 256) {
      const int t = i / per_tap, r = i - t * per_tap, ch = r / NS, n = r - ch * NS;
      reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(Wp) + ((size_t)(t * C::kCH + ch) * NT + col0 + n));
    }
    if (SKIP)
      for (int i = tid; i < per_tap; i += 256) {
        const int ch = i / NS, n = i - ch * NS;
        reinterpret_cast<uint4*>(wsk)[i] = __ldg(reinterpret_cast<const uint4*>(Wsk) + ((size_t)ch * NT + col0 + n));
      }
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  tma::cluster_sync();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 6) {
    // ===================== TMA producer (both CTAs; each loads its own tile) =====================
    if (umma::elect_one()) {
      const int nboxes = xslab_rows / kBoxRows;
      int slot = 0;
      uint32_t ph = 0;
      for (int pt = pair; pt < n_pair_tiles; pt += n_pairs) {
        const int tile = 2 * pt + (int)rank;            // may be == n_tiles (odd tile count): lands in the zero guard rows
        const int row0 = guard + tile * kTileM - halo;
#pragma unroll 1
        for (int rg = 0; rg < C::kRegions; ++rg) {
          umma::mbar_wait(empty + slot, ph ^ 1);
          uint8_t* slab = slab0 + (size_t)slot * unit_bytes;
          if (rank == 0) tma::expect_tx(full + slot, (uint32_t)(2 * unit_bytes));
          const uint32_t lbar = map_to_rank(umma::smem_u32(full + slot), 0);
          for (int b = 0; b < nboxes; ++b)
            load_tile_2sm(slab + (size_t)b * kBoxRows * 128, &tmX, rg * 64, row0 + b * kBoxRows, lbar);
          if (++slot == ring) { slot = 0; ph ^= 1; }
        }
      }
      for (int i = 0; i < ring; ++i) {   // drain
        umma::mbar_wait(empty + slot, ph ^ 1);
        if (++slot == ring) { slot = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if ((warp == 4 || warp == 5) && rank == 0) {
    // ===================== MMA issuers (leader only) =====================
    constexpr uint32_t idesc = umma::make_idesc_bf16(2 * kTileM, NT, 0, 0);   // 256 x 128 across the pair
    const uint64_t da_base = tma::make_desc_sw(umma::smem_u32(slab0) + (uint32_t)(halo * 128), 0u, 128u);
    const uint64_t db_base = umma::make_desc(umma::smem_u32(wsm), (uint32_t)NS * 16u, 128u);
    const uint64_t dbsk_base = umma::make_desc(umma::smem_u32(wsk), (uint32_t)NS * 16u, 128u);
    const uint32_t a_hi = (uint32_t)(da_base >> 32), b_hi = (uint32_t)(db_base >> 32);
    const uint32_t unit_units = (uint32_t)unit_bytes >> 4;
    int dl[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) dl[t] = shifts.d[t] * 8;
    const int my_buf = warp - 4;
    int k = 0;
    for (int pt = pair; pt < n_pair_tiles; pt += n_pairs, ++k) {
      if ((k & 1) != my_buf) continue;
      const uint32_t tph = (k >> 1) & 1;
      const uint32_t acc = tmem + (uint32_t)(my_buf * C::kAccCols);
#pragma unroll 1
      for (int rg = 0; rg < C::kRegions; ++rg) {
        const int u = k * C::kRegions + rg;
        const int slot = u % ring;
        const uint32_t ph = (uint32_t)(u / ring) & 1u;
        umma::mbar_wait(full + slot, ph);
        if (rg == 0) umma::mbar_wait(tempty + my_buf, tph ^ 1);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          const uint32_t a_lo0 = (uint32_t)da_base + (uint32_t)slot * unit_units;
          const uint32_t b_lo0 = (uint32_t)db_base + (uint32_t)(rg * 8 * NS);
#pragma unroll
          for (int t = 0; t < TAPS; ++t) {
            const uint32_t a_t = a_lo0 + (uint32_t)dl[t];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t da = ((uint64_t)a_hi << 32) | (a_t + (uint32_t)(2 * j));
              const uint64_t db = ((uint64_t)b_hi << 32) | (b_lo0 + (uint32_t)((t * C::kCH + 2 * j) * NS));
              mma_bf16_2sm(acc, da, db, idesc, (rg > 0 || t > 0 || j > 0) ? 1u : 0u);
            }
          }
          if (SKIP) {
            const uint32_t bs_lo0 = (uint32_t)dbsk_base + (uint32_t)(rg * 8 * NS);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t da = ((uint64_t)a_hi << 32) | (a_lo0 + (uint32_t)(2 * j));
              const uint64_t db = ((uint64_t)b_hi << 32) | (bs_lo0 + (uint32_t)((2 * j) * NS));
              mma_bf16_2sm(acc + NT, da, db, idesc, (rg > 0 || j > 0) ? 1u : 0u);
            }
          }
          commit_2sm(empty + slot);
          if (rg == C::kRegions - 1) commit_2sm(tfull + my_buf);
        }
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    // ===================== epilogue warps 0-3 (both CTAs) =====================
    float ssum[C::kOutRegions][8], ssq[C::kOutRegions][8];
#pragma unroll
    for (int o = 0; o < C::kOutRegions; ++o)
#pragma unroll
      for (int i = 0; i < 8; ++i) ssum[o][i] = ssq[o][i] = 0.f;
    const int my_row = warp * 32 + lane;
    const int my_swz = my_row & 7;
    const int sc = tid & 7, sg = tid >> 3;
    // BSTAT: raw rows of this thread's (8-row group, 16-byte chunk) for the region being processed and for the NEXT one
    // (prefetched a whole region ahead: one region of epilogue work is ~1000 cycles, about the loaded DRAM latency; a
    // prefetch issued only at the top of its own region stalled the epilogue and cost 0.4 ms per launch).
    static_assert(!BSTAT || C::kOutPerAcc == 2, "the raw-tile double buffer assumes two 64-column regions per tile");
    uint4 rawv[BSTAT ? 2 : 1][BSTAT ? 8 : 1];
    auto prefetch_raw = [&](int tile_, int q_, uint4 (&dst)[BSTAT ? 8 : 1]) {
      const uint4* rp = reinterpret_cast<const uint4*>(bn_raw + ((long long)tile_ * kTileM + sg * 8) * NT + q_ * 64 + sc * 8);
#pragma unroll
      for (int i = 0; i < (BSTAT ? 8 : 1); ++i) {
        // A COHERENT load on purpose: ptxas is free to sink a non-coherent (ld.global.nc / __ldg) load below the bar.sync
        // of the next region to save registers -- it did, which put these loads ~100 instructions before their consumers
        // instead of a region ahead (ncu source page: the unpack of rawv carried the long-scoreboard stalls).  A weak
        // ld.global may not move down across a barrier.
        const uint4* p = rp + (size_t)i * (NT / 8);
        asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(dst[i].x), "=r"(dst[i].y), "=r"(dst[i].z), "=r"(dst[i].w)
                     : "l"(p)
                     : "memory");
      }
    };
    if (BSTAT && pair < n_pair_tiles) prefetch_raw(2 * pair + (int)rank, 0, rawv[0]);
    uint32_t sidx = 0;
    int k = 0;
    RowWalker rw;
    rw.init((long long)(2 * pair + (int)rank) * kTileM + my_row, (long long)2 * n_pairs * kTileM, P);
    for (int pt = pair; pt < n_pair_tiles; pt += n_pairs, ++k, rw.next()) {
      const int tile = 2 * pt + (int)rank;
      const int buf = k & 1;
      umma::mbar_wait(tfull + buf, (k >> 1) & 1);
      umma::fence_after_sync();
      const bool valid = rw.valid(rows, P);
      const uint32_t acc = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * C::kAccCols);
      const uint32_t leader_tempty = map_to_rank(umma::smem_u32(tempty + buf), 0);
#pragma unroll
      for (int o = 0; o < (SKIP ? 2 : 1); ++o) {
        const bool want_stats = (o == 0 ? stats : stats_sk) != nullptr;
#pragma unroll
        for (int q = 0; q < C::kOutPerAcc; ++q, ++sidx) {
          uint8_t* stg = stage0 + (sidx % C::kStageBufs) * C::kStageBytes;
          if (BSTAT) {
            if (q == 0) prefetch_raw(tile, 1, rawv[1]);
            else if (pt + n_pairs < n_pair_tiles) prefetch_raw(2 * (pt + n_pairs) + (int)rank, 0, rawv[0]);
          }
          // BSTAT stores Y with plain 16-byte stores from the statistics loop below (no TMA store, hence no generic->async
          // proxy fence in this path: the fence drained the raw-tile loads in flight and exposed their full DRAM latency)
          if (!BSTAT && tid == 0) bulk_wait_read4<C::kStageBufs - 1>();
          epi_bar_sync4();
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float v[32];
            umma::tmem_ld32(acc + (uint32_t)(o * NT + q * 64 + g * 32), v);
            if (o == (SKIP ? 1 : 0) && q == C::kOutPerAcc - 1 && g == 1) {
              umma::fence_before_sync();   // last TMEM read of this accumulator: hand it back to the leader's MMA warps
              __syncwarp();
              if (lane == 0) arrive_cluster(leader_tempty);
            }
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              uint4 pk;
              uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = valid ? v[c4 * 8 + 2 * e] : 0.f, b = valid ? v[c4 * 8 + 2 * e + 1] : 0.f;
                __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                pw[e] = *reinterpret_cast<uint32_t*>(&h);
              }
              const int chunk = g * 4 + c4;
              *reinterpret_cast<uint4*>(stg + my_row * 128 + ((chunk ^ my_swz) << 4)) = pk;
            }
          }
          if (!BSTAT) umma::fence_proxy_async();
          epi_bar_sync4();
          if (!BSTAT && tid == 0 && tile < n_tiles) tma_store_tile4(o == 0 ? &tmY : &tmYsk, stg, q * 64, guard + tile * kTileM);
          if (want_stats && (!BSTAT || tile < n_tiles)) {   // (the odd leftover tile reads uninitialised guard rows of raw)
            float bsa[8], bha[8];   // BatchNorm scale / shift of this thread's 8 channels of region q (L1 resident)
            if (BSTAT) {
              const float4* sp = reinterpret_cast<const float4*>(bn_ss + q * 64 + sc * 8);
              const float4 a0 = __ldg(sp), a1 = __ldg(sp + 1), h0 = __ldg(sp + NT / 4), h1 = __ldg(sp + NT / 4 + 1);
              bsa[0] = a0.x; bsa[1] = a0.y; bsa[2] = a0.z; bsa[3] = a0.w; bsa[4] = a1.x; bsa[5] = a1.y; bsa[6] = a1.z; bsa[7] = a1.w;
              bha[0] = h0.x; bha[1] = h0.y; bha[2] = h0.z; bha[3] = h0.w; bha[4] = h1.x; bha[5] = h1.y; bha[6] = h1.z; bha[7] = h1.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = sg * 8 + i;
              const uint4 u = *reinterpret_cast<const uint4*>(stg + row * 128 + ((sc ^ (row & 7)) << 4));
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
              if (BSTAT) {
                y_out[((long long)tile * kTileM + row) * (NT / 8) + q * 8 + sc] = u;
                const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&rawv[BSTAT ? (q & 1) : 0][BSTAT ? i : 0]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(h[e]), r = __bfloat1622float2(rh[e]);
                  const float gx = fmaf(r.x, bsa[2 * e], bha[2 * e]) > 0.f ? f.x : 0.f;           // == bn.cu bn_pre<false>
                  const float gy = fmaf(r.y, bsa[2 * e + 1], bha[2 * e + 1]) > 0.f ? f.y : 0.f;
                  ssum[q][2 * e] += gx;
                  ssum[q][2 * e + 1] += gy;
                  ssq[q][2 * e] = fmaf(gx, r.x, ssq[q][2 * e]);
                  ssq[q][2 * e + 1] = fmaf(gy, r.y, ssq[q][2 * e + 1]);
                }
                continue;
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                ssum[o * C::kOutPerAcc + q][2 * e] += f.x;
                ssum[o * C::kOutPerAcc + q][2 * e + 1] += f.y;
                ssq[o * C::kOutPerAcc + q][2 * e] = fmaf(f.x, f.x, ssq[o * C::kOutPerAcc + q][2 * e]);
                ssq[o * C::kOutPerAcc + q][2 * e + 1] = fmaf(f.y, f.y, ssq[o * C::kOutPerAcc + q][2 * e + 1]);
              }
            }
          }
        }
      }
    }
    if (tid == 0) bulk_wait_read4<0>();
    if (k > 0) {
#pragma unroll
      for (int o = 0; o < (SKIP ? 2 : 1); ++o) {
        if ((o == 0 ? stats : stats_sk) == nullptr) continue;
#pragma unroll
        for (int q = 0; q < C::kOutPerAcc; ++q) {
          float* rr = red + (o * C::kOutPerAcc + q) * 2 * 64;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            atomicAdd(rr + sc * 8 + i, ssum[o * C::kOutPerAcc + q][i]);
            atomicAdd(rr + 64 + sc * 8 + i, ssq[o * C::kOutPerAcc + q][i]);
          }
        }
      }
      epi_bar_sync4();
#pragma unroll
      for (int o = 0; o < (SKIP ? 2 : 1); ++o) {
        float* st = o == 0 ? stats : stats_sk;
        if (st == nullptr) continue;
        for (int i = tid; i < C::kOutPerAcc * 2 * 64; i += 128) {
          const int q = i / 128, rem = i - q * 128, which = rem / 64, ch = rem - which * 64;
          atomicAdd(st + which * NT + q * 64 + ch, red[(o * C::kOutPerAcc + q) * 2 * 64 + rem]);
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  tma::cluster_sync();
  if (warp == 0) tmem_dealloc_2sm<C::kTmemCols>(tmem);
}

template <int CIN, bool SKIP, int TAPS, int NT, bool BSTAT = false>
int launch4(const __nv_bfloat16* X, const __nv_bfloat16* Wp, const __nv_bfloat16* Wsk, __nv_bfloat16* Y, __nv_bfloat16* Ysk,
            float* stats, float* stats_sk, long long rows, int P, const ConvShifts& sh, cudaStream_t st, bool* fits,
            const __nv_bfloat16* bn_raw = nullptr, const float* bn_ss = nullptr) {
  using C = Cfg4<CIN, SKIP, TAPS, NT>;
  constexpr int taps = TAPS;
  constexpr int guard = 128;
  const int halo = taps == 1 ? 0 : P + 2;
  const int xslab_rows = (kTileM + 2 * halo + kBoxRows - 1) / kBoxRows * kBoxRows;
  *fits = halo <= guard - kBoxRows;
  if (!*fits) return MIVIT_OK;
  const int unit_bytes = xslab_rows * 128;
  const int fixed = C::kWBytes + C::kWSkipBytes + C::kStageBufs * C::kStageBytes;
  const int tail = (2 * kMaxRing + 4) * 8 + 16 + C::kOutRegions * 2 * 64 * 4 + 64;
  int ring = (227 * 1024 - fixed - tail) / unit_bytes;
  if (ring > kMaxRing) ring = kMaxRing;
  *fits = ring >= 2 && ring >= C::kRegions;
  if (!*fits) return MIVIT_OK;
  int smem = fixed + ring * unit_bytes + tail;
  if (smem < 120 * 1024) smem = 120 * 1024;            // one CTA per SM
  auto kern = conv_rows_tc4_kernel<CIN, SKIP, TAPS, NT, BSTAT>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long rows_pad = (rows + kTileM - 1) / kTileM * kTileM;
  const int n_tiles = (int)(rows_pad / kTileM);
  CUtensorMap tmX, tmY, tmYsk;
  {
    int rc = make_rows_tensor_map_sw(&tmX, X - (size_t)guard * CIN, CIN, rows_pad + 2 * guard, kBoxRows);
    if (rc) return rc;
    rc = make_rows_tensor_map_sw(&tmY, Y - (size_t)guard * NT, NT, rows_pad + 2 * guard, kTileM);
    if (rc) return rc;
    tmYsk = tmY;
    if (SKIP) {
      rc = make_rows_tensor_map_sw(&tmYsk, Ysk - (size_t)guard * NT, NT, rows_pad + 2 * guard, kTileM);
      if (rc) return rc;
    }
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int pairs = sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(pairs * 2, 1, 1);
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0 && max_clusters < pairs)
    pairs = max_clusters;
  if (pairs > (n_tiles + 1) / 2) pairs = (n_tiles + 1) / 2;
  if (pairs < 1) pairs = 1;
  cfg.gridDim = dim3(pairs * 2, 1, 1);
  char tag[48];
  snprintf(tag, sizeof(tag), "conv_rows_tc_%dx%dx%d%s", CIN, NT, taps, SKIP ? "+skip" : BSTAT ? "+bnbwd" : "");
  const double valid_rows = (double)rows * P * P / ((double)(P + 1) * (P + 1));
  MivitProfScope prof(tag, 2.0 * valid_rows * (taps + (SKIP ? 1 : 0)) * CIN * NT, st);
  MIVIT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmX, tmY, tmYsk, Wp, Wsk, stats, stats_sk, rows, n_tiles, P, sh, halo, xslab_rows,
                                      ring, guard, bn_raw, bn_ss, reinterpret_cast<uint4*>(Y)));
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

}  // namespace

// C_out = 128 shapes on CTA pairs; *handled = false -> the caller uses conv_tc3.cu.
int conv_rows_forward_v4(const __nv_bfloat16* X, const __nv_bfloat16* Wp, const __nv_bfloat16* Wsk, __nv_bfloat16* Y,
                         __nv_bfloat16* Ysk, float* stats, float* stats_sk, long long rows, int P, int cin, int cout, int taps,
                         const ConvShifts& sh, cudaStream_t st, bool* handled) {
  const bool skip = Wsk != nullptr;
  *handled = true;
  bool fits = true;
  int rc = MIVIT_OK;
  if (cout == 128 && taps == 9 && cin == 128 && !skip) rc = launch4<128, false, 9, 128>(X, Wp, Wsk, Y, Ysk, stats, stats_sk, rows, P, sh, st, &fits);
  // C_out = 64 on pairs (N = 64 per 256-row MMA, 32 weight columns per CTA) was measured too: 64 -> 64 0.69 ms vs 0.39 ms on
  // conv_tc3, 128 -> 64 0.76 vs 0.79 ms -- each SM still streams 5 KB per 32 tensor cycles -- so those shapes stay on conv_tc3.
  else if (cout == 64 && taps == 9 && cin == 128 && !skip && getenv("MIVIT_PAIRS_C64") != nullptr)
    rc = launch4<128, false, 9, 64>(X, Wp, Wsk, Y, Ysk, stats, stats_sk, rows, P, sh, st, &fits);
  // (64 -> 128 + skip is epilogue-bound -- 256 staged columns per tile -- and measured 1.14 ms on pairs vs 0.99 ms on
  //  conv_tc3's independent half-column CTAs, so it stays there; launch4<64, true, 9> is kept compilable for experiments.
  //  Round 2: a SECOND epilogue group on pairs (warps 8-11, alternate tiles, own staging tiles / barrier / bulk groups) changed
  //  nothing -- 1.066 vs 1.023 ms with one group, 0.733 vs 0.734 ms without the statistics loop -- and ncu then shows both groups
  //  waiting for the accumulator 25 % of the time: the tile is bounded by shared-memory traffic (operand reads 243 KB + staging
  //  writes / TMA-store reads / statistics reads 192 KB per 128-row tile and SM), not by epilogue issue slots.)
  else if (cout == 128 && taps == 9 && cin == 64 && skip && getenv("MIVIT_PAIRS_SKIP") != nullptr)
    rc = launch4<64, true, 9, 128>(X, Wp, Wsk, Y, Ysk, stats, stats_sk, rows, P, sh, st, &fits);
  else if (cout == 128 && taps == 9 && cin == 64 && !skip) rc = launch4<64, false, 9, 128>(X, Wp, Wsk, Y, Ysk, stats, stats_sk, rows, P, sh, st, &fits);
  else fits = false;
  if (!fits) *handled = false;
  return rc;
}

// Input gradient of a 3x3 convolution whose input was act = relu(bn(raw)):  Y = g = dL/dact, and sums[0..C) += sum m*g,
// sums[C..2C) += sum m*g*raw with m = 1[raw*scale + shift > 0] (ss = scale | shift; sums pre-zeroed by the caller).
// *handled = false: shape not covered here, nothing was launched.
int conv_rows_dgrad_bnsums(const __nv_bfloat16* X, const __nv_bfloat16* Wp, __nv_bfloat16* Y, const __nv_bfloat16* raw,
                           const float* ss, float* sums, long long rows, int P, int cin, int cout, const ConvShifts& sh,
                           cudaStream_t st, bool* handled) {
  *handled = false;
  bool fits = false;
  int rc = MIVIT_OK;
  if (cin == 128 && cout == 128)
    rc = launch4<128, false, 9, 128, true>(X, Wp, nullptr, Y, nullptr, sums, nullptr, rows, P, sh, st, &fits, raw, ss);
  *handled = fits;
  return rc;
}
