// TMA-fed version of the shifted-row implicit-GEMM convolution (algorithm and layout: conv_tc.cu; the
// pipelined cp.async version: conv_tc2.cu).  What changed, and the measurement behind it (scripts/bench_conv.py,
// 6.02 M rows): v2 spent 4200-14000 cycles per 128-row tile against 1150-2560 cycles of tensor work because
// (a) only two row slabs fit the ring and each slab's HBM latency was exposed, (b) the epilogue warps did the
// BatchNorm column sums with 62 shuffles per 32 columns and wrote the tile with 16-byte stores 128-256 B apart.
//   * input: one elected thread streams 32-row TMA boxes into a ring of up to 8 swizzled [row][128 B] tiles
//     (tma.cuh); the 9 taps are row-shifted descriptor start addresses.  The pipeline unit is one 64-channel
//     region of a tile, so even 128-channel inputs have >= 3 units in flight next to 147 KB of resident weights;
//   * when the output columns are split over two CTAs (resident weights of 128 -> 128 and 64 -> 128 do not fit
//     one SM) the two CTAs form a cluster and multicast each other's half of the boxes: a row slab crosses
//     L2 -> SM once;
//   * output: the epilogue warps stage the bf16 tile in shared memory in the same swizzled layout (conflict-free
//     16-byte stores), one thread writes it out with a TMA store, and the BatchNorm statistics are column sums
//     read back from the staged tile with a (chunk, 8-row group) thread mapping -- 8 LDS.128 and 128 FMAs per
//     thread and region, no shuffles, partial sums carried in registers across the CTA's tiles.
// Warp roles (256 threads, 384 with two epilogue groups): warps 0-3 epilogue group 0 (TMEM lane quarter = warp id & 3), warps 4 / 5 MMA issuers (even /
// odd tiles, two accumulators), warp 6 TMA producer, warp 7 idle, warps 8-11 epilogue group 1.
// Two epilogue groups (even tiles -> group 0, odd tiles -> group 1, each with its own accumulator, staging buffers, named
// barrier and statistics partials) are used for C_in = 32 only.  With C_out <= 64 per CTA a tile is 1150-1300 cycles of
// tensor work but ~2450 cycles of epilogue (TMEM read, bf16 pack, staging, TMA store, BatchNorm column sums) for ONE group of
// four single-issue warps (ncu: tensor pipe 53 % active, the epilogue warps busy 92 % of the time).  Measured with two
// groups: 32 -> 64 + skip 0.533 -> 0.440 ms, but 64 -> 64 0.369 -> 0.460 ms, 64 -> 128 + skip 0.976 -> 1.212 ms and
// 64 -> 32 (dual input) 0.352 -> 0.377 ms: at C_in >= 64 the N = 64 MMAs already want more shared-memory bandwidth than the
// SM has (192 B/cycle, see conv_tc4.cu), and a second group's staging stores, statistics loads and TMA-store reads running
// concurrently with them take it away from the tensor core.
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kTileM = 128;
constexpr int kBoxRows = 32;
constexpr int kMaxRing = 8;

__device__ __forceinline__ void mbar_arrive3(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void epi_bar_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); }

__device__ __forceinline__ void tma_store_tile(const CUtensorMap* tm, const void* smem_src, int ch0, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(ch0), "r"(row),
               "r"(umma::smem_u32(smem_src))
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int CIN, int NS, bool SKIP, int TAPS, int CIN2 = 0>
struct Cfg3 {
  static constexpr int kCH = CIN / 8;
  static constexpr int kWBytes = TAPS * CIN * NS * 2;
  static constexpr int kW2Bytes = CIN2 * NS * 2;                     // second input (1x1 path into the same accumulator)
  static constexpr int kRegions2 = CIN2 / 64;
  static constexpr int kStageBufs = kWBytes + kW2Bytes > 128 * 1024 ? 1 : 2;      // 147 KB of resident weights leave room for one
  static constexpr int kEpiGroups = CIN <= 32 ? 2 : 1;      // epilogue warp groups (even / odd tiles), see the header comment
  static constexpr int kThreads = kEpiGroups == 2 ? 384 : 256;
  static constexpr int kRegions = CIN >= 64 ? CIN / 64 : 1;          // pipeline units (64-channel regions) per tile
  static constexpr int kUnits = kRegions + kRegions2;
  static constexpr int kPitch = CIN >= 64 ? 128 : 64;                // bytes per input-tile row
  static constexpr int kKPerRegion = (CIN >= 64 ? 64 : CIN) / 16;    // K = 16 MMAs per tap and region
  static constexpr int kWSkipBytes = SKIP ? CIN * NS * 2 : 0;
  static constexpr int kAccCols = NS * (SKIP ? 2 : 1);
  static constexpr int kTmemCols = 2 * kAccCols <= 32 ? 32 : 2 * kAccCols <= 64 ? 64 : 2 * kAccCols <= 128 ? 128
                                   : 2 * kAccCols <= 256 ? 256 : 512;
  static constexpr int kOutW = NS >= 64 ? 64 : NS;                   // staged output region width (channels)
  static constexpr int kOutPitch = kOutW * 2;                        // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
  static constexpr int kOutPerAcc = NS / kOutW;
  static constexpr int kOutRegions = kOutPerAcc * (SKIP ? 2 : 1);
  static constexpr int kStageBytes = kTileM * kOutPitch;
};

template <int CIN, int NS, bool SKIP, int TAPS, int CIN2>
__global__ void __launch_bounds__((Cfg3<CIN, NS, SKIP, TAPS, CIN2>::kThreads), 1)
conv_rows_tc3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                     const __grid_constant__ CUtensorMap tmYsk, const __grid_constant__ CUtensorMap tmX2,
                     const __nv_bfloat16* __restrict__ Wp, const __nv_bfloat16* __restrict__ Wsk,
                     const __nv_bfloat16* __restrict__ Wp2, float* __restrict__ stats, float* __restrict__ stats_sk, long long rows,
                     int n_tiles, int P, ConvShifts shifts, int halo, int xslab_rows, int cout_total, int nsplit, int ring, int guard,
                     __nv_bfloat16* __restrict__ Yp, __nv_bfloat16* __restrict__ Yskp, int use_tma_store) {
  using C = Cfg3<CIN, NS, SKIP, TAPS, CIN2>;
  constexpr int taps = TAPS;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_uniform(), lane = tid & 31;
  const int unit_bytes = xslab_rows * C::kPitch;
  uint8_t* wsm = smem;                                   // [taps][CH][NS][8]   K-major, no swizzle
  uint8_t* wsk = wsm + taps * CIN * NS * 2;              // [CH][NS][8]         (SKIP)
  uint8_t* w2sm = wsk + C::kWSkipBytes;                  // [CIN2/8][NS][8]     (second input, 1 tap)
  uint8_t* stage0 = w2sm + C::kW2Bytes;                  // 1-2 staged output regions (swizzled [row][kOutPitch])
  uint8_t* slab0 = stage0 + C::kEpiGroups * C::kStageBufs * C::kStageBytes;   // ring of input units
  uint8_t* tail = slab0 + (size_t)ring * unit_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);    // [kMaxRing] TMA -> MMA
  uint64_t* empty = full + kMaxRing;                     // [kMaxRing] MMA (commit) -> TMA
  uint64_t* tfull = empty + kMaxRing;                    // [2] MMA (commit) -> epilogue
  uint64_t* tempty = tfull + 2;                          // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* red = reinterpret_cast<float*>(tmem_slot + 4);  // [kEpiGroups][kOutRegions][2][kOutW] statistics of the CTA

  const int nh = blockIdx.x % nsplit;                    // output-column slice of this CTA = rank in the cluster
  const int cta_in_slice = blockIdx.x / nsplit, ctas_per_slice = gridDim.x / nsplit;
  const int col0 = nh * NS;
  const uint16_t cmask = (uint16_t)((1u << nsplit) - 1u);

  if (tid == 0) {
    for (int i = 0; i < kMaxRing; ++i) {
      umma::mbar_init(full + i, 1);
      umma::mbar_init(empty + i, nsplit);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(tfull + i, 1);
      umma::mbar_init(tempty + i, 4);
    }
    umma::mbar_fence_init();
    tma::prefetch_map(&tmX);
    tma::prefetch_map(&tmY);
    if (SKIP) tma::prefetch_map(&tmYsk);
    if (CIN2 > 0) tma::prefetch_map(&tmX2);
  }
  if (warp == 0) umma::tmem_alloc<C::kTmemCols>(tmem_slot);
  for (int i = tid; i < C::kEpiGroups * C::kOutRegions * 2 * C::kOutW; i += C::kThreads) red[i] = 0.f;
  // resident weights of this column slice
  {
    const int per_tap = C::kCH * NS;  // 16-byte units per tap in smem
    for (int i = tid; i < taps * per_tap; i += C::kThreads) {
      const int t = i / per_tap, r = i - t * per_tap, ch = r / NS, n = r - ch * NS;
      const uint4* src = reinterpret_cast<const uint4*>(Wp) + ((size_t)(t * C::kCH + ch) * cout_total + col0 + n);
      reinterpret_cast<uint4*>(wsm)[i] = __ldg(src);
    }
    if (SKIP) {
      for (int i = tid; i < per_tap; i += C::kThreads) {
        const int ch = i / NS, n = i - ch * NS;
        const uint4* src = reinterpret_cast<const uint4*>(Wsk) + ((size_t)ch * cout_total + col0 + n);
        reinterpret_cast<uint4*>(wsk)[i] = __ldg(src);
      }
    }
    if (CIN2 > 0) {
      for (int i = tid; i < (CIN2 / 8) * NS; i += C::kThreads) {
        const int ch = i / NS, n = i - ch * NS;
        const uint4* src = reinterpret_cast<const uint4*>(Wp2) + ((size_t)ch * cout_total + col0 + n);
        reinterpret_cast<uint4*>(w2sm)[i] = __ldg(src);
      }
    }
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  if (nsplit > 1) tma::cluster_sync();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 6) {
    // ===================== TMA producer =====================
    if (umma::elect_one()) {
      const int nboxes = xslab_rows / kBoxRows;
      int slot = 0;
      uint32_t ph = 0;
      for (int tile = cta_in_slice; tile < n_tiles; tile += ctas_per_slice) {
        const int row0 = guard + tile * kTileM - halo;
#pragma unroll 1
        for (int rg = 0; rg < C::kUnits; ++rg) {
          umma::mbar_wait(empty + slot, ph ^ 1);
          uint8_t* slab = slab0 + (size_t)slot * unit_bytes;
          const bool second = rg >= C::kRegions;            // a region of the second input: 128 rows, no halo
          const int nb = second ? kTileM / kBoxRows : nboxes;
          tma::expect_tx(full + slot, (uint32_t)(nb * kBoxRows * C::kPitch));
          for (int b = nh; b < nb; b += nsplit) {
            uint8_t* dst = slab + (size_t)b * kBoxRows * C::kPitch;
            const CUtensorMap* tm = second ? &tmX2 : &tmX;
            const int ch0 = second ? (rg - C::kRegions) * 64 : rg * 64;
            const int row = (second ? row0 + halo : row0) + b * kBoxRows;
            if (nsplit == 1) tma::load_tile(dst, tm, ch0, row, full + slot);
            else tma::load_tile_multicast(dst, tm, ch0, row, full + slot, cmask);
          }
          if (++slot == ring) { slot = 0; ph ^= 1; }
        }
      }
      for (int i = 0; i < ring; ++i) {   // drain: every (also remote) release of the slots has arrived before exit
        umma::mbar_wait(empty + slot, ph ^ 1);
        if (++slot == ring) { slot = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 4 || warp == 5) {
    // ===================== MMA issuers: warp 4 -> even tiles (accumulator 0), warp 5 -> odd (accumulator 1).
    // Descriptors are (constant high word | start-address low word); only the low word is advanced
    // (tap: +delta rows, K step: +32 B inside the 128-byte swizzled row, weights: +2 chunks).
    constexpr uint32_t idesc = umma::make_idesc_bf16(kTileM, NS, 0, 0);
    const uint64_t da_base = tma::make_desc_sw(umma::smem_u32(slab0) + (uint32_t)(halo * C::kPitch), 0u, (uint32_t)C::kPitch);
    const uint64_t db_base = umma::make_desc(umma::smem_u32(wsm), (uint32_t)NS * 16u, 128u);
    const uint64_t dbsk_base = umma::make_desc(umma::smem_u32(wsk), (uint32_t)NS * 16u, 128u);
    const uint64_t db2_base = umma::make_desc(umma::smem_u32(w2sm), (uint32_t)NS * 16u, 128u);
    const uint32_t a_hi = (uint32_t)(da_base >> 32), b_hi = (uint32_t)(db_base >> 32);
    const uint32_t unit_units = (uint32_t)unit_bytes >> 4;
    int dl[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) dl[t] = shifts.d[t] * (C::kPitch / 16);
    const int my_buf = warp - 4;
    int k = 0;
    for (int tile = cta_in_slice; tile < n_tiles; tile += ctas_per_slice, ++k) {
      if ((k & 1) != my_buf) continue;
      const uint32_t tph = (k >> 1) & 1;
      const uint32_t acc = tmem + (uint32_t)(my_buf * C::kAccCols);
#pragma unroll 1
      for (int rg = 0; rg < C::kUnits; ++rg) {
        const int u = k * C::kUnits + rg;
        const int slot = u % ring;
        const uint32_t ph = (uint32_t)(u / ring) & 1u;
        umma::mbar_wait(full + slot, ph);
        if (rg == 0) umma::mbar_wait(tempty + my_buf, tph ^ 1);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          const uint32_t a_lo0 = (uint32_t)da_base + (uint32_t)slot * unit_units;
          if (CIN2 == 0 || rg < C::kRegions) {
            const uint32_t b_lo0 = (uint32_t)db_base + (uint32_t)(rg * 8 * NS);
#pragma unroll
            for (int t = 0; t < TAPS; ++t) {
              const uint32_t a_t = a_lo0 + (uint32_t)dl[t];
#pragma unroll
              for (int j = 0; j < C::kKPerRegion; ++j) {
                const uint64_t da = ((uint64_t)a_hi << 32) | (a_t + (uint32_t)(2 * j));
                const uint64_t db = ((uint64_t)b_hi << 32) | (b_lo0 + (uint32_t)((t * C::kCH + 2 * j) * NS));
                umma::mma_bf16(acc, da, db, idesc, (rg > 0 || t > 0 || j > 0) ? 1u : 0u);
              }
            }
            if (SKIP) {
              const uint32_t bs_lo0 = (uint32_t)dbsk_base + (uint32_t)(rg * 8 * NS);
#pragma unroll
              for (int j = 0; j < C::kKPerRegion; ++j) {
                const uint64_t da = ((uint64_t)a_hi << 32) | (a_lo0 + (uint32_t)(2 * j));
                const uint64_t db = ((uint64_t)b_hi << 32) | (bs_lo0 + (uint32_t)((2 * j) * NS));
                umma::mma_bf16(acc + NS, da, db, idesc, (rg > 0 || j > 0) ? 1u : 0u);
              }
            }
          } else {   // second input: rows [tile, tile+128) sit at the START of the slot (no halo), one tap, same accumulator
            const uint32_t a2 = a_lo0 - (uint32_t)(halo * (C::kPitch / 16));
            const uint32_t b2 = (uint32_t)db2_base + (uint32_t)((rg - C::kRegions) * 8 * NS);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t da = ((uint64_t)a_hi << 32) | (a2 + (uint32_t)(2 * j));
              const uint64_t db = ((uint64_t)b_hi << 32) | (b2 + (uint32_t)((2 * j) * NS));
              umma::mma_bf16(acc, da, db, idesc, 1u);
            }
          }
          if (nsplit == 1) umma::commit(empty + slot);           // unit may be refilled once these MMAs have read it
          else tma::commit_multicast(empty + slot, cmask);       // ... in both CTAs of the pair
          if (rg == C::kUnits - 1) umma::commit(tfull + my_buf);  // accumulator ready for the epilogue
        }
        __syncwarp();
      }
    }
  } else if (warp < 4 || (warp >= 8 && C::kEpiGroups == 2)) {
    // ===================== epilogue: group 0 = warps 0-3 (even tiles), group 1 = warps 8-11 (odd tiles) =====================
    const int grp = warp >= 8 ? 1 : 0;
    const int ew = warp & 3;                              // TMEM lane quarter of this warp
    const int etid = ew * 32 + lane;                      // thread index inside the group
    uint8_t* gstage = stage0 + grp * C::kStageBufs * C::kStageBytes;
    float* gred = red + grp * C::kOutRegions * 2 * C::kOutW;
    constexpr int kChunks = C::kOutPitch / 16;            // 16-byte chunks per staged row
    constexpr int kRowsPerThread = 128 / (128 / kChunks); // stats: thread = (chunk, row group); rows per group
    float ssum[C::kOutRegions][8], ssq[C::kOutRegions][8];
#pragma unroll
    for (int o = 0; o < C::kOutRegions; ++o)
#pragma unroll
      for (int i = 0; i < 8; ++i) ssum[o][i] = ssq[o][i] = 0.f;
    const int my_row = ew * 32 + lane;
    const int my_swz = C::kOutPitch == 128 ? (my_row & 7) : ((my_row >> 1) & 3);
    const int sc = etid & (kChunks - 1), sg = etid / kChunks;     // statistics mapping
    uint32_t sidx = 0;
    int k = 0;
    RowWalker rw;
    rw.init((long long)cta_in_slice * kTileM + my_row, (long long)ctas_per_slice * kTileM, P);
    for (int tile = cta_in_slice; tile < n_tiles; tile += ctas_per_slice, ++k, rw.next()) {
      const int buf = k & 1;
      if (C::kEpiGroups == 2 && buf != grp) continue;
      umma::mbar_wait(tfull + buf, (k >> 1) & 1);
      umma::fence_after_sync();
      const bool valid = rw.valid(rows, P);
      const uint32_t acc = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * C::kAccCols);
#pragma unroll
      for (int o = 0; o < (SKIP ? 2 : 1); ++o) {
        const bool want_stats = (o == 0 ? stats : stats_sk) != nullptr;
#pragma unroll
        for (int q = 0; q < C::kOutPerAcc; ++q, ++sidx) {
          uint8_t* stg = gstage + (sidx % C::kStageBufs) * C::kStageBytes;
          // the TMA store that read this staging buffer two regions ago is done; everyone finished its statistics reads
          if (use_tma_store && etid == 0) bulk_wait_read<C::kStageBufs - 1>();
          epi_bar_sync(grp);
#pragma unroll
          for (int g = 0; g < C::kOutW / 32; ++g) {
            float v[32];
            umma::tmem_ld32(acc + (uint32_t)(o * NS + q * C::kOutW + g * 32), v);
            if (o == (SKIP ? 1 : 0) && q == C::kOutPerAcc - 1 && g == C::kOutW / 32 - 1) {
              umma::fence_before_sync();   // last TMEM read of this accumulator: hand it back to the MMA warp
              __syncwarp();
              if (lane == 0) mbar_arrive3(tempty + buf);
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
              *reinterpret_cast<uint4*>(stg + my_row * C::kOutPitch + ((chunk ^ my_swz) << 4)) = pk;
            }
          }
          if (use_tma_store) {
            umma::fence_proxy_async();   // generic-proxy writes of the staged tile -> visible to the TMA store
            epi_bar_sync(grp);
            if (etid == 0) tma_store_tile(o == 0 ? &tmY : &tmYsk, stg, col0 + q * C::kOutW, guard + tile * kTileM);
          } else {
            epi_bar_sync(grp);
          }
          // MIVIT_DIRECT_STORE=1 (experiment): the tile leaves through plain 16-byte stores from the loop that reads it back for the statistics (every
          // thread: one chunk of kRowsPerThread rows, 8 lanes = one 128-byte row segment).  A TMA store is one request per
          // 128-byte ROW on the unit that also feeds the operand ring (~1 row per 6-8 cycles per SM): with
          // C_out <= 64 per CTA the 128-256 stored rows per tile outnumbered the 80-160 loaded ones and set the tile time.
          if (want_stats || !use_tma_store) {
            uint4* yout = reinterpret_cast<uint4*>(o == 0 ? Yp : Yskp) +
                          (((long long)tile * kTileM + sg * kRowsPerThread) * cout_total + col0 + q * C::kOutW) / 8 + sc;
            // column sums of the STORED (bf16-rounded, pad-masked) values
#pragma unroll
            for (int i = 0; i < kRowsPerThread; ++i) {
              const int row = sg * kRowsPerThread + i;
              const int sw = C::kOutPitch == 128 ? (row & 7) : ((row >> 1) & 3);
              const uint4 u = *reinterpret_cast<const uint4*>(stg + row * C::kOutPitch + ((sc ^ sw) << 4));
              if (!use_tma_store) yout[(size_t)i * (cout_total / 8)] = u;
              if (!want_stats) continue;
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
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
    if (use_tma_store && etid == 0) bulk_wait_read<0>();   // shared memory stays valid until the last store has read it
    // CTA-level reduction of the per-thread column sums, then one atomic per channel
    if (k > 0) {
#pragma unroll
      for (int o = 0; o < (SKIP ? 2 : 1); ++o) {
        if ((o == 0 ? stats : stats_sk) == nullptr) continue;
#pragma unroll
        for (int q = 0; q < C::kOutPerAcc; ++q) {
          float* rr = gred + (o * C::kOutPerAcc + q) * 2 * C::kOutW;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            atomicAdd(rr + sc * 8 + i, ssum[o * C::kOutPerAcc + q][i]);
            atomicAdd(rr + C::kOutW + sc * 8 + i, ssq[o * C::kOutPerAcc + q][i]);
          }
        }
      }
      epi_bar_sync(grp);
#pragma unroll
      for (int o = 0; o < (SKIP ? 2 : 1); ++o) {
        float* st = o == 0 ? stats : stats_sk;
        if (st == nullptr) continue;
        for (int i = etid; i < C::kOutPerAcc * 2 * C::kOutW; i += 128) {
          const int q = i / (2 * C::kOutW), rem = i - q * 2 * C::kOutW, which = rem / C::kOutW, ch = rem - which * C::kOutW;
          atomicAdd(st + which * cout_total + col0 + q * C::kOutW + ch, gred[(o * C::kOutPerAcc + q) * 2 * C::kOutW + rem]);
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (nsplit > 1) tma::cluster_sync();
  if (warp == 0) umma::tmem_dealloc<C::kTmemCols>(tmem);
}

template <int CIN, int NS, bool SKIP, int TAPS, int CIN2 = 0>
int launch3(const __nv_bfloat16* X, const __nv_bfloat16* Wp, const __nv_bfloat16* Wsk, __nv_bfloat16* Y, __nv_bfloat16* Ysk,
            float* stats, float* stats_sk, long long rows, int P, int cout_total, const ConvShifts& sh, cudaStream_t st,
            bool* fits, const __nv_bfloat16* X2 = nullptr, const __nv_bfloat16* Wp2 = nullptr) {
  using C = Cfg3<CIN, NS, SKIP, TAPS, CIN2>;
  constexpr int taps = TAPS;
  constexpr int guard = 128;
  const int halo = taps == 1 ? 0 : P + 2;
  const int xslab_rows = (kTileM + 2 * halo + kBoxRows - 1) / kBoxRows * kBoxRows;
  *fits = halo <= guard - kBoxRows;
  if (!*fits) return MIVIT_OK;
  const int unit_bytes = xslab_rows * C::kPitch;
  const int fixed = taps * CIN * NS * 2 + C::kWSkipBytes + C::kW2Bytes + C::kEpiGroups * C::kStageBufs * C::kStageBytes;
  const int tail = (2 * kMaxRing + 4) * 8 + 16 + C::kEpiGroups * C::kOutRegions * 2 * C::kOutW * 4 + 64;
  int ring = (227 * 1024 - fixed - tail) / unit_bytes;
  if (ring > kMaxRing) ring = kMaxRing;
  *fits = ring >= 2 && ring >= C::kRegions && (CIN2 == 0 || (ring >= 3 && C::kPitch == 128));
  if (!*fits) return MIVIT_OK;
  int smem = fixed + ring * unit_bytes + tail;
  if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM: the TMEM budget assumes it
  auto kern = conv_rows_tc3_kernel<CIN, NS, SKIP, TAPS, CIN2>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long rows_pad = (rows + kTileM - 1) / kTileM * kTileM;
  const int n_tiles = (int)(rows_pad / kTileM);
  const int nsplit = cout_total / NS;
  CUtensorMap tmX, tmY, tmYsk, tmX2;
  {
    int rc = make_rows_tensor_map_sw(&tmX, X - (size_t)guard * CIN, CIN, rows_pad + 2 * guard, kBoxRows);
    if (rc) return rc;
    rc = make_rows_tensor_map_sw(&tmY, Y - (size_t)guard * cout_total, cout_total, rows_pad + 2 * guard, kTileM, C::kOutW);
    if (rc) return rc;
    tmYsk = tmY;
    if (SKIP) {
      rc = make_rows_tensor_map_sw(&tmYsk, Ysk - (size_t)guard * cout_total, cout_total, rows_pad + 2 * guard, kTileM, C::kOutW);
      if (rc) return rc;
    }
    tmX2 = tmX;
    if (CIN2 > 0) {
      rc = make_rows_tensor_map_sw(&tmX2, X2 - (size_t)guard * CIN2, CIN2, rows_pad + 2 * guard, kBoxRows);
      if (rc) return rc;
    }
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int per_slice = sms / nsplit;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(C::kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nsplit;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (nsplit > 1) {
    cfg.gridDim = dim3(per_slice * nsplit, 1, 1);
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0 && max_clusters < per_slice)
      per_slice = max_clusters;
  }
  if (per_slice > n_tiles) per_slice = n_tiles;
  if (per_slice < 1) per_slice = 1;
  cfg.gridDim = dim3(per_slice * nsplit, 1, 1);
  char tag[48];
  snprintf(tag, sizeof(tag), "conv_rows_tc_%dx%dx%d%s", CIN, cout_total, taps, SKIP ? "+skip" : CIN2 ? "+in2" : "");
  const double valid_rows = (double)rows * P * P / ((double)(P + 1) * (P + 1));
  MivitProfScope prof(tag, 2.0 * valid_rows * ((double)(taps + (SKIP ? 1 : 0)) * CIN + CIN2) * cout_total, st);
  static const int use_tma_store = getenv("MIVIT_DIRECT_STORE") != nullptr ? 0 : 1;   // A/B switch, see the epilogue
  MIVIT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmX, tmY, tmYsk, tmX2, Wp, Wsk, Wp2, stats, stats_sk, rows, n_tiles, P, sh, halo,
                                      xslab_rows, cout_total, nsplit, ring, guard, Y, Ysk, use_tma_store));
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

}  // namespace

// Returns MIVIT_OK and sets *handled = false when the configuration is not covered (the caller then uses the
// cp.async kernel of conv_tc2.cu).
int conv_rows_forward_v3(const __nv_bfloat16* X, const __nv_bfloat16* Wp, const __nv_bfloat16* Wsk, __nv_bfloat16* Y,
                         __nv_bfloat16* Ysk, float* stats, float* stats_sk, long long rows, int P, int cin, int cout, int taps,
                         const ConvShifts& sh, cudaStream_t st, bool* handled) {
  const bool skip = Wsk != nullptr;
  *handled = true;
  bool fits = true;
  int rc = MIVIT_OK;
#define V3_CASE(CI, CO, NS_, SK)                                                                                       \
  if (cin == CI && cout == CO && skip == SK) {                                                                         \
    rc = taps == 9 ? launch3<CI, NS_, SK, 9>(X, Wp, Wsk, Y, Ysk, stats, stats_sk, rows, P, cout, sh, st, &fits)        \
                   : launch3<CI, NS_, SK, 1>(X, Wp, Wsk, Y, Ysk, stats, stats_sk, rows, P, cout, sh, st, &fits);       \
    if (!fits) *handled = false;                                                                                       \
    return rc;                                                                                                         \
  }
  V3_CASE(32, 64, 64, true)
  V3_CASE(32, 64, 64, false)
  V3_CASE(64, 64, 64, false)
  V3_CASE(64, 128, 64, true)
  V3_CASE(64, 128, 64, false)
  V3_CASE(128, 128, 64, false)
  V3_CASE(64, 32, 32, false)
  V3_CASE(128, 64, 64, false)
#undef V3_CASE
  *handled = false;
  return MIVIT_OK;
}

// Y = conv3x3(X, Wp) + conv1x1(X2, Wp2): the input gradient of a ResidualBlock's input arrives through conv1 (3x3) and
// through the skip convolution (1x1) (helpers/models.py:221-226 backwards); one accumulator takes both, so the two
// partial gradients are never written to HBM and BatchNorm's backward reads one upstream tensor instead of two.
int conv_rows_forward_dual_v3(const __nv_bfloat16* X, const __nv_bfloat16* Wp, const __nv_bfloat16* X2, const __nv_bfloat16* Wp2,
                              __nv_bfloat16* Y, long long rows, int P, int cin, int cin2, int cout, const ConvShifts& sh,
                              cudaStream_t st, bool* handled) {
  *handled = true;
  bool fits = true;
  int rc = MIVIT_OK;
  // (128 + 128 -> 64 was measured too: its resident weights force 32-column slices, and N = 32 MMAs re-read the 4 KB A
  // operand from shared memory every 16 cycles -- 1.42 ms against 0.74 + 0.36 ms for the two separate kernels -- so block 2
  // keeps the two-kernel path.)
  if (cin == 64 && cin2 == 64 && cout == 32) {
    rc = launch3<64, 32, false, 9, 64>(X, Wp, nullptr, Y, nullptr, nullptr, nullptr, rows, P, cout, sh, st, &fits, X2, Wp2);
  } else {
    fits = false;
  }
  if (!fits) *handled = false;
  return rc;
}
