// TMA-fed weight-gradient kernel (formulation in conv_wgrad.cu):
//     dW[co, ci, tap] = sum_r dY[r, co] * X[r + delta_tap, ci]      (rows = K, both operands MN-major)
// v3 differences from conv_wgrad2.cu, driven by the measurement that v2 spent ~3300 cycles per 128-row
// stage against 770 cycles of tensor work (load latency exposed by a 2-deep ring, every row range re-read
// by three tap groups from L2):
//   * one elected thread fills a ring of up to 8 (dY, X-slab) stages with 32-row TMA boxes that land as
//     [row][128 B] SWIZZLE_128B tiles (tma.cuh; a first version with 16-byte-wide boxes into the no-swizzle
//     core-matrix order was TMA-issue bound at ~7 B/cycle/SM), completion by mbarrier complete_tx -- no producer
//     warps, no per-16-byte address arithmetic; taps are row-shifted descriptor start addresses into the tile;
//   * a CTA owns whole kernel ROWS (3 taps): all three for C_in = 32, one for C_in >= 64 (TMEM columns).  The
//     three taps of a kernel row are consecutive activation rows, so for C_in <= 64 they are ONE MMA of
//     N = 3*C_in whose descriptor "next channel group" stride (LBO) is one tile row: 3x fewer tcgen05.mma
//     to issue (the per-tap N = 32 / 64 chains were issue-bound at 2-3x the tensor floor);
//   * the G = 1..3 CTAs that cover the 9 taps of the SAME row range form a thread-block cluster and share
//     the loads: each CTA fetches 1/G of the 32-row boxes and the TMA multicasts them into every CTA of
//     the cluster, so a row range crosses L2 -> SM once instead of three times.  A ring slot is refilled
//     only after the MMAs of all G CTAs have released it (tcgen05.commit multicast on the empty barrier).
// Accumulators stay in TMEM over the CTA's whole row range and are flushed once with fp32 atomics.
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kStageRows = 128;
constexpr int kMaxRing = 8;
constexpr int kIssuers = 3;

constexpr int kBoxRows = 32;

template <int CIN, int COUT>
struct WgCfg3 {
  static constexpr int kTapsPerCta = CIN <= 32 ? 9 : 3;
  static constexpr int kARegions = COUT / 64, kBRegions = CIN >= 64 ? CIN / 64 : 1;
  static constexpr int kBPitch = CIN >= 64 ? 128 : 64;                 // bytes per X-tile row
  static constexpr int kARegionBytes = kStageRows * 128;
  static constexpr int kABytes = kARegions * kARegionBytes;
  static constexpr int kOverRead = COUT < 128 ? kARegionBytes : 0;     // M = 128 MMA on a 64-channel dY tile
};

// grid = (row-range CTAs, G tap groups); cluster = (1, G, 1): cluster rank = blockIdx.y = tap group.
// warp 0: TMA producer.  warps 1-3: MMA issuers (taps t = i, i+3, i+6 of the CTA's group).  warps 4-7: flush.
template <int CIN, int COUT>
__global__ void __launch_bounds__(256, 1)
conv_wgrad_tc3_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, float* __restrict__ dW,
                      int n_stages, int stages_per_cta, int taps, ConvShifts shifts, int halo, int xslab_rows, int stage_bytes,
                      int ring, int groups, int guard, const __grid_constant__ CUtensorMap tmY2, float* __restrict__ dW2) {
  using C = WgCfg3<CIN, COUT>;
  // dW2 != nullptr (9 taps, one CTA per row range): the weight gradient of the block's 1x1 SKIP convolution rides along -- same
  // input X, its own output gradient dY2 staged next to dY, one more MMA per 16 rows against the centre tap's rows, one more
  // accumulator (columns taps*CIN ...): X is read once for both and the skip's own stream kernel disappears.
  const bool has2 = dW2 != nullptr;
  const int a_bytes = C::kABytes * (has2 ? 2 : 1);
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_uniform(), lane = tid & 31;
  uint8_t* bars_base = smem + (size_t)ring * stage_bytes + C::kOverRead;   // past the M = 128 over-read of the last slot
  uint64_t* full = reinterpret_cast<uint64_t*>(bars_base);   // [kMaxRing]
  uint64_t* empty = full + kMaxRing;                         // [kMaxRing]
  uint64_t* done = empty + kMaxRing;                         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int grp = blockIdx.y;
  const int tap0 = grp * C::kTapsPerCta;
  const int ntap = min(C::kTapsPerCta, taps - tap0);
  const int s_begin = blockIdx.x * stages_per_cta;
  const int s_end = min(n_stages, s_begin + stages_per_cta);
  // issuers (one per kernel row) in this CTA and in the whole cluster (arrival count of `empty`)
  const int my_issuers = taps == 9 ? ntap / 3 : 1;
  const int cluster_issuers = taps == 9 ? 3 : 1;
  const uint16_t cmask = (uint16_t)((1u << groups) - 1u);

  if (tid == 0) {
    for (int i = 0; i < kMaxRing; ++i) {
      umma::mbar_init(full + i, 1);
      umma::mbar_init(empty + i, cluster_issuers);
    }
    umma::mbar_init(done, my_issuers);
    umma::mbar_fence_init();
    tma::prefetch_map(&tmY);
    tma::prefetch_map(&tmX);
  }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  if (groups > 1) tma::cluster_sync();   // every CTA's barriers are initialised before any remote arrive / multicast
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (umma::elect_one()) {
      const uint32_t bytes = (uint32_t)(a_bytes + C::kBRegions * xslab_rows * C::kBPitch);
      // load units = 32-row boxes: A region-major (kARegions x 4) (twice with a second dY), then B (kBRegions x xslab_rows/32);
      // CTA g takes u = g mod G
      const int a1_units = C::kARegions * (kStageRows / kBoxRows), a_units = a1_units * (has2 ? 2 : 1), xb = xslab_rows / kBoxRows;
      const int n_units = a_units + C::kBRegions * xb;
      int slot = 0;
      uint32_t ph = 0;
      for (int s = s_begin; s < s_end; ++s) {
        umma::mbar_wait(empty + slot, ph ^ 1);
        uint8_t* aslab = smem + (size_t)slot * stage_bytes;
        uint8_t* bslab = aslab + a_bytes;
        tma::expect_tx(full + slot, bytes);
        const int r0 = guard + s * kStageRows;
        for (int u = grp; u < n_units; u += groups) {
          uint8_t* dst;
          const CUtensorMap* tm;
          int ch0, row;
          if (u < a_units) {
            const int second = u >= a1_units ? 1 : 0, v = u - second * a1_units;
            const int reg = v / (kStageRows / kBoxRows), rb = v - reg * (kStageRows / kBoxRows);
            dst = aslab + (size_t)second * C::kABytes + (size_t)reg * C::kARegionBytes + (size_t)rb * kBoxRows * 128;
            tm = second ? &tmY2 : &tmY; ch0 = reg * 64; row = r0 + rb * kBoxRows;
          } else {
            const int v = u - a_units, reg = v / xb, rb = v - reg * xb;
            dst = bslab + ((size_t)reg * xslab_rows + (size_t)rb * kBoxRows) * C::kBPitch;
            tm = &tmX; ch0 = reg * 64; row = r0 - halo + rb * kBoxRows;
          }
          if (groups == 1) tma::load_tile(dst, tm, ch0, row, full + slot);
          else tma::load_tile_multicast(dst, tm, ch0, row, full + slot, cmask);
        }
        if (++slot == ring) { slot = 0; ph ^= 1; }
      }
      // drain: the last release of every slot has arrived (also the remote, multicast ones) before this CTA may exit
      for (int i = 0; i < ring; ++i) {
        umma::mbar_wait(empty + slot, ph ^ 1);
        if (++slot == ring) { slot = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp <= kIssuers) {
    // ===================== MMA issuers: one warp per kernel row (3 taps) of this CTA =====================
    const int iss = warp - 1;
    const int row_taps = taps == 9 ? 3 : 1;
    if (iss * row_taps < ntap) {
      // the taps of a kernel row as ONE MMA per 64-channel region of X: N = 3 * min(CIN, 64), LBO = one tile row.  For
      // C_in = 128 that is two N = 192 MMAs per 16 rows instead of three of N = 128: the dY operand crosses shared memory
      // twice per kernel row instead of three times (20 KB instead of 24 KB per 192 tensor cycles).
      constexpr int kRegionCh = CIN < 64 ? CIN : 64;
      const uint32_t idesc = umma::make_idesc_bf16(128, row_taps * kRegionCh, 1, 1);
      // MN-major swizzled row tiles: LBO = next channel group, SBO = 8 rows; one row = pitch/16 address units
      const uint64_t da0 = tma::make_desc_sw(umma::smem_u32(smem), (uint32_t)C::kARegionBytes, 128u);
      const uint64_t db0 = tma::make_desc_sw(umma::smem_u32(smem) + (uint32_t)a_bytes + (uint32_t)(halo * C::kBPitch), (uint32_t)C::kBPitch,
                                             (uint32_t)C::kBPitch);
      // skip product (issued with the CENTRE kernel row, whose middle tap has shift 0): M = 128 (dY2), N = min(CIN, 64) per region
      const bool skip_here = has2 && taps == 9 && iss == 1;
      const uint32_t idesc2 = umma::make_idesc_bf16(128, kRegionCh, 1, 1);
      const uint32_t acc2 = tmem + (uint32_t)(taps * CIN);
      const uint32_t region_units = (uint32_t)(xslab_rows * C::kBPitch) >> 4;   // next 64-channel region of the X slab
      const uint32_t a_hi = (uint32_t)(da0 >> 32), b_hi = (uint32_t)(db0 >> 32);
      const uint32_t stage_units = (uint32_t)stage_bytes >> 4;
      const int t_first = iss * row_taps;                                  // first tap of this issuer's kernel row (CTA-local)
      const int d_first = shifts.d[tap0 + t_first];
      const uint32_t acc0 = tmem + (uint32_t)(t_first * CIN);
      int slot = 0;
      uint32_t ph = 0;
      bool first = true;
      for (int s = s_begin; s < s_end; ++s) {
        umma::mbar_wait(full + slot, ph);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          const uint32_t a_lo0 = (uint32_t)da0 + (uint32_t)slot * stage_units;
          const uint32_t b_lo0 = (uint32_t)db0 + (uint32_t)slot * stage_units + (uint32_t)(d_first * (C::kBPitch / 16));
#pragma unroll
          for (int kk = 0; kk < kStageRows / 16; ++kk) {
            const uint64_t da = ((uint64_t)a_hi << 32) | (a_lo0 + (uint32_t)(kk * 16 * 8));                    // 16 rows x 128 B
            const uint32_t b_lo = b_lo0 + (uint32_t)(kk * 16 * (C::kBPitch / 16));
#pragma unroll
            for (int reg = 0; reg < C::kBRegions; ++reg)
              umma::mma_bf16(acc0 + (uint32_t)(reg * row_taps * kRegionCh), da, ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)reg * region_units),
                             idesc, (!first || kk > 0) ? 1u : 0u);
            if (skip_here) {
              const uint64_t da2 = da + (uint64_t)(C::kABytes >> 4);                        // the second dY stage
              const uint32_t b_mid = b_lo + (uint32_t)(C::kBPitch / 16);                    // centre tap = first tap of the row + 1
#pragma unroll
              for (int reg = 0; reg < C::kBRegions; ++reg)
                umma::mma_bf16(acc2 + (uint32_t)(reg * kRegionCh), da2, ((uint64_t)b_hi << 32) | (b_mid + (uint32_t)reg * region_units), idesc2,
                               (!first || kk > 0) ? 1u : 0u);
            }
          }
          if (groups == 1) umma::commit(empty + slot);
          else tma::commit_multicast(empty + slot, cmask);
          if (s == s_end - 1) umma::commit(done);
        }
        __syncwarp();
        first = false;
        if (++slot == ring) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4 && s_end > s_begin) {
    // ===================== flush (warps 4-7): lane = co, columns = (tap, ci) =====================
    const int q = warp - 4;
    umma::mbar_wait(done, 0);
    umma::fence_after_sync();
    if (q * 32 < COUT) {
      const int co = q * 32 + lane;
      // accumulator columns: (issuer's kernel row, 64-channel region, tap of the row, channel of the region)
      const int row_taps = taps == 9 ? 3 : 1;
      constexpr int kRegionCh = CIN < 64 ? CIN : 64;
      for (int t = 0; t < ntap; ++t) {
        const int kr = t / row_taps, tr = t - kr * row_taps;
#pragma unroll
        for (int cg = 0; cg < CIN / 32; ++cg) {
          float v[32];
          const int reg = cg * 32 / 64, within = cg * 32 - reg * 64;
          umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) +
                              (uint32_t)(kr * row_taps * CIN + reg * row_taps * kRegionCh + tr * kRegionCh + within), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(dW + ((size_t)co * CIN + cg * 32 + i) * taps + tap0 + t, v[i]);
        }
      }
      if (has2) {   // the skip convolution's gradient: columns (64-channel region, channel of the region) behind the taps
#pragma unroll
        for (int cg = 0; cg < CIN / 32; ++cg) {
          float v[32];
          umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(taps * CIN + cg * 32), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(dW2 + (size_t)co * CIN + cg * 32 + i, v[i]);
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (groups > 1) tma::cluster_sync();   // no CTA exits while a peer may still multicast into it / arrive on its barriers
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

template <int CIN, int COUT>
int launch_wgrad3(const __nv_bfloat16* X, const __nv_bfloat16* dY, float* dW, long long rows, int P, int taps,
                  const ConvShifts& sh, cudaStream_t st, bool* fits, const __nv_bfloat16* dY2 = nullptr, float* dW2 = nullptr) {
  using C = WgCfg3<CIN, COUT>;
  constexpr int guard = 128;
  const int halo = taps == 1 ? 0 : P + 2;
  const int xslab_rows = (kStageRows + 2 * halo + kBoxRows - 1) / kBoxRows * kBoxRows;   // whole 32-row TMA boxes
  *fits = halo <= guard - kBoxRows;
  if (taps == 9)   // the taps of a kernel row must be consecutive activation rows (true for stride-1 / pad-1 shifts)
    for (int kh = 0; kh < 3; ++kh) *fits = *fits && sh.d[kh * 3 + 1] == sh.d[kh * 3] + 1 && sh.d[kh * 3 + 2] == sh.d[kh * 3] + 2;
  if (!*fits) return MIVIT_OK;
  const bool has2 = dW2 != nullptr;
  if (has2 && (taps != 9 || C::kTapsPerCta != 9 || (taps + 1) * CIN > 512)) { *fits = false; return MIVIT_OK; }
  int stage_bytes = C::kABytes * (has2 ? 2 : 1) + C::kBRegions * xslab_rows * C::kBPitch;
  stage_bytes = (stage_bytes + 1023) & ~1023;   // swizzle atoms are 1024-byte aligned
  const int tail = C::kOverRead + (2 * kMaxRing + 1) * 8 + 64;
  int ring = (227 * 1024 - tail) / stage_bytes;
  if (ring > kMaxRing) ring = kMaxRing;
  *fits = ring >= 2;
  if (!*fits) return MIVIT_OK;
  int smem = ring * stage_bytes + tail;
  if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM (512 TMEM columns)
  auto kern = conv_wgrad_tc3_kernel<CIN, COUT>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long rows_pad = (rows + kStageRows - 1) / kStageRows * kStageRows;
  const int n_stages = (int)(rows_pad / kStageRows);
  const int groups = (taps + C::kTapsPerCta - 1) / C::kTapsPerCta;
  CUtensorMap tmY, tmX, tmY2;
  {
    int rc = make_rows_tensor_map_sw(&tmY, dY - (size_t)guard * COUT, COUT, rows_pad + 2 * guard, kBoxRows);
    if (rc) return rc;
    tmY2 = tmY;
    if (has2) {
      rc = make_rows_tensor_map_sw(&tmY2, dY2 - (size_t)guard * COUT, COUT, rows_pad + 2 * guard, kBoxRows);
      if (rc) return rc;
    }
    rc = make_rows_tensor_map_sw(&tmX, X - (size_t)guard * CIN, CIN, rows_pad + 2 * guard, kBoxRows);
    if (rc) return rc;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int ctas_x = sms / groups;
  if (groups > 1) {   // only as many clusters as can be co-resident (GPC granularity): one wave, balanced row ranges
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(sms / groups, groups, 1);
    q.blockDim = dim3(256, 1, 1);
    q.dynamicSmemBytes = smem;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 1; qa[0].val.clusterDim.y = groups; qa[0].val.clusterDim.z = 1;
    q.attrs = qa; q.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &q) == cudaSuccess && max_clusters > 0 && max_clusters < ctas_x)
      ctas_x = max_clusters;
  }
  if (ctas_x < 1) ctas_x = 1;
  if (ctas_x > n_stages) ctas_x = n_stages;
  const int spc = (n_stages + ctas_x - 1) / ctas_x;
  ctas_x = (n_stages + spc - 1) / spc;
  char tag[48];
  snprintf(tag, sizeof(tag), "conv_wgrad_tc_%dx%dx%d%s", CIN, COUT, taps, has2 ? "+skip" : "");
  const double valid_rows = (double)rows * P * P / ((double)(P + 1) * (P + 1));
  MivitProfScope prof(tag, 2.0 * valid_rows * (taps + (has2 ? 1 : 0)) * CIN * COUT, st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas_x, groups, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = groups;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MIVIT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmY, tmX, dW, n_stages, spc, taps, sh, halo, xslab_rows, stage_bytes, ring, groups, guard,
                                      tmY2, dW2));
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

}  // namespace

int conv_rows_wgrad_v3(const __nv_bfloat16* X, const __nv_bfloat16* dY, float* dW, long long rows, int P, int cin, int cout,
                       int taps, const ConvShifts& sh, cudaStream_t st, bool* handled) {
  *handled = true;
  bool fits = true;
  int rc;
#define WG3_CASE(CI, CO)                                                        \
  if (cin == CI && cout == CO) {                                                \
    rc = launch_wgrad3<CI, CO>(X, dY, dW, rows, P, taps, sh, st, &fits);        \
    if (!fits) *handled = false;                                                \
    return rc;                                                                  \
  }
  WG3_CASE(32, 64)
  WG3_CASE(64, 64)
  WG3_CASE(64, 128)
  WG3_CASE(128, 128)
#undef WG3_CASE
  *handled = false;
  return MIVIT_OK;
}

// 3x3 weight gradient + the 1x1 skip convolution's weight gradient of the same input in one launch (C_in = 32: all nine taps
// and the skip accumulator fit one CTA's tensor memory).  *handled = false: not covered, nothing was launched.
int conv_rows_wgrad_skip_v3(const __nv_bfloat16* X, const __nv_bfloat16* dY, const __nv_bfloat16* dYskip, float* dW, float* dWskip,
                            long long rows, int P, int cin, int cout, const ConvShifts& sh, cudaStream_t st, bool* handled) {
  static const bool off = getenv("MIVIT_NO_WGRAD_SKIP") != nullptr;   // A/B switch
  *handled = false;
  if (off || cin != 32 || cout != 64) return MIVIT_OK;
  bool fits = true;
  const int rc = launch_wgrad3<32, 64>(X, dY, dW, rows, P, 9, sh, st, &fits, dYskip, dWskip);
  *handled = fits;
  return rc;
}
