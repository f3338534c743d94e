// Backward of a 64 -> 64 channel 3x3 convolution (ResidualBlock 1's conv2, helpers/models.py:221-226) in ONE pass over its
// output gradient dY [rows, 64], input X [rows, 64]:
//     dW[co, ci, tap] = sum_r dY[r, co] * X[r + delta_tap, ci]
//     dX[r, ci]       = sum_tap sum_co dY[r - delta_tap, co] * W[co, ci, tap]
// Same idea as conv_block1_bwd.cu (both products are bound by TMA row requests, so the rows are shared), plus a TRANSPOSED
// weight-gradient product: the clustered kernel of conv_wgrad3.cu computes D[co][(tap, ci)] with M = 128 MMAs of which 64 lanes
// (C_out = 64) are wasted and needs 9 x 64 = 576 TMEM columns, i.e. three CTAs per row range.  Here
//     D[(tap pair, ci)][co] += X(shifted)^T dY        A = the input slab, MN-major, M = two taps x 64 channels: the descriptor's
//                                                      group stride is the row distance between the two taps; B = dY, N = 64
// so the nine taps are five M = 128 MMAs per 16 rows (the last one half used) and 5 x 64 = 320 TMEM columns in ONE CTA, next to
// two 64-column dgrad accumulators.
//   warp 0 producer | warp 1 weight-gradient issuer | warp 2 dgrad issuer | warps 4-7 dgrad epilogue, then the weight-gradient flush.
// dX's pad rows are written with whatever the taps produce there: its only consumer, bn1's backward, reads valid pixels only.
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kStageRows = 128;
constexpr int kBoxRows = 32;
constexpr int kMaxRing = 4;
constexpr int CH = 64;                              // C_in = C_out
constexpr int kWBytes = 9 * CH * CH * 2;            // packed dgrad weights [9][CH/8][CH][8]
constexpr int kPairs = 5;
constexpr int kAccDx = kPairs * CH;                 // TMEM columns: [0, 320) tap pairs, [320, 448) two dgrad accumulators

__device__ __forceinline__ void mbar_arrive_cta64(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(mbar)) : "memory");
}

__global__ void __launch_bounds__(256, 1)
conv_layer64_bwd_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                        const __nv_bfloat16* __restrict__ Wd, float* __restrict__ dW, __nv_bfloat16* __restrict__ dX, int n_stages,
                        int stages_per_cta, ConvShifts sh_w, ConvShifts sh_d, int halo, int slab_rows, int ring, int guard) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_uniform(), lane = tid & 31;
  const int slab_bytes = slab_rows * 128;
  const int stage_bytes = 2 * slab_bytes;                          // [dY slab][X slab]
  uint8_t* wsm = smem + (size_t)ring * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(wsm + kWBytes);    // [kMaxRing]
  uint64_t* empty = full + kMaxRing;                               // [kMaxRing] both issuers
  uint64_t* tfull = empty + kMaxRing;                              // [2]
  uint64_t* tempty = tfull + 2;                                    // [2]
  uint64_t* done = tempty + 2;                                     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int s_begin = blockIdx.x * stages_per_cta;
  const int s_end = min(n_stages, s_begin + stages_per_cta);

  if (tid == 0) {
    for (int i = 0; i < kMaxRing; ++i) {
      umma::mbar_init(full + i, 1);
      umma::mbar_init(empty + i, 2);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(tfull + i, 1);
      umma::mbar_init(tempty + i, 4);
    }
    umma::mbar_init(done, 1);
    umma::mbar_fence_init();
    tma::prefetch_map(&tmY);
    tma::prefetch_map(&tmX);
  }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  for (int i = tid; i < kWBytes / 16; i += 256) reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(Wd) + i);
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t stage_units = (uint32_t)stage_bytes >> 4;
  const uint32_t smem0 = umma::smem_u32(smem);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (umma::elect_one()) {
      const int nb = slab_rows / kBoxRows;
      int slot = 0;
      uint32_t ph = 0;
      for (int s = s_begin; s < s_end; ++s) {
        umma::mbar_wait(empty + slot, ph ^ 1);
        uint8_t* ys = smem + (size_t)slot * stage_bytes;
        uint8_t* xs = ys + slab_bytes;
        tma::expect_tx(full + slot, (uint32_t)stage_bytes);
        const int r0 = guard + s * kStageRows - halo;
        for (int b = 0; b < nb; ++b) tma::load_tile(ys + (size_t)b * kBoxRows * 128, &tmY, 0, r0 + b * kBoxRows, full + slot);
        for (int b = 0; b < nb; ++b) tma::load_tile(xs + (size_t)b * kBoxRows * 128, &tmX, 0, r0 + b * kBoxRows, full + slot);
        if (++slot == ring) { slot = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== weight gradient, transposed: D[(tap pair, ci)][co] += X(shifted)^T dY, both MN-major =====================
    const uint32_t idesc = umma::make_idesc_bf16(128, CH, 1, 1);
    // B = the centre 128 rows of the dY slab (N = 64 output channels = one group)
    const uint64_t db0 = tma::make_desc_sw(smem0 + (uint32_t)(halo * 128), 128u, 128u);
    const uint32_t b_hi = (uint32_t)(db0 >> 32);
    // A = the X slab at the first tap of the pair; group stride (LBO) = row distance to the second tap (any positive distance
    // for the lone ninth tap: its second group lands in accumulator lanes 64-127, which the flush skips)
    uint32_t a_hi[kPairs], a_lo[kPairs];     // (the group stride lives in the descriptor's LOW word, next to the start address)
#pragma unroll
    for (int p = 0; p < kPairs; ++p) {
      const int t0 = 2 * p, t1 = 2 * p + 1 < 9 ? 2 * p + 1 : -1;
      const int dist = t1 >= 0 ? sh_w.d[t1] - sh_w.d[t0] : 1;
      const uint64_t d = tma::make_desc_sw(smem0 + (uint32_t)slab_bytes + (uint32_t)((halo + sh_w.d[t0]) * 128), (uint32_t)(dist * 128), 128u);
      a_hi[p] = (uint32_t)(d >> 32);
      a_lo[p] = (uint32_t)d;
    }
    int slot = 0;
    uint32_t ph = 0;
    bool first = true;
    for (int s = s_begin; s < s_end; ++s) {
      umma::mbar_wait(full + slot, ph);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t b_lo0 = (uint32_t)db0 + (uint32_t)slot * stage_units;
        const uint32_t soff = (uint32_t)slot * stage_units;
#pragma unroll
        for (int kk = 0; kk < kStageRows / 16; ++kk)
#pragma unroll
          for (int p = 0; p < kPairs; ++p)
            umma::mma_bf16(tmem + (uint32_t)(p * CH), ((uint64_t)a_hi[p] << 32) | (a_lo[p] + soff + (uint32_t)(kk * 128)),
                           ((uint64_t)b_hi << 32) | (b_lo0 + (uint32_t)(kk * 128)), idesc, (!first || kk > 0) ? 1u : 0u);
        umma::commit(empty + slot);
        if (s == s_end - 1) umma::commit(done);
      }
      __syncwarp();
      first = false;
      if (++slot == ring) { slot = 0; ph ^= 1; }
    }
  } else if (warp == 2) {
    // ===================== input gradient: D[r][ci] = sum_tap dY[r + shift] W[tap], dY slab K-major (rows = M) =====================
    constexpr uint32_t idesc = umma::make_idesc_bf16(128, CH, 0, 0);
    const uint64_t da0 = tma::make_desc_sw(smem0 + (uint32_t)(halo * 128), 0u, 128u);
    const uint64_t db0 = umma::make_desc(umma::smem_u32(wsm), (uint32_t)CH * 16u, 128u);
    const uint32_t a_hi = (uint32_t)(da0 >> 32), b_hi = (uint32_t)(db0 >> 32);
    int dl[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) dl[t] = sh_d.d[t] * 8;
    int slot = 0;
    uint32_t ph = 0;
    int k = 0;
    for (int s = s_begin; s < s_end; ++s, ++k) {
      const int buf = k & 1;
      umma::mbar_wait(full + slot, ph);
      umma::mbar_wait(tempty + buf, ((k >> 1) & 1) ^ 1);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t acc = tmem + (uint32_t)(kAccDx + buf * CH);
        const uint32_t a_lo0 = (uint32_t)da0 + (uint32_t)slot * stage_units;
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int j = 0; j < CH / 16; ++j)
            umma::mma_bf16(acc, ((uint64_t)a_hi << 32) | (a_lo0 + (uint32_t)dl[t] + (uint32_t)(2 * j)),
                           ((uint64_t)b_hi << 32) | ((uint32_t)db0 + (uint32_t)((t * (CH / 8) + 2 * j) * CH)), idesc, (t > 0 || j > 0) ? 1u : 0u);
        umma::commit(empty + slot);
        umma::commit(tfull + buf);
      }
      __syncwarp();
      if (++slot == ring) { slot = 0; ph ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== dgrad epilogue, then the weight-gradient flush =====================
    const int q = warp - 4;
    int k = 0;
    for (int s = s_begin; s < s_end; ++s, ++k) {
      const int buf = k & 1;
      umma::mbar_wait(tfull + buf, (k >> 1) & 1);
      umma::fence_after_sync();
      uint4* out = reinterpret_cast<uint4*>(dX + ((size_t)s * kStageRows + q * 32 + lane) * CH);
#pragma unroll
      for (int g = 0; g < CH / 32; ++g) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(kAccDx + buf * CH + g * 32), v);
        if (g == CH / 32 - 1) {
          umma::fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta64(tempty + buf);
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint4 pk;
          uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[c4 * 8 + 2 * e], v[c4 * 8 + 2 * e + 1]);
            pw[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          out[g * 4 + c4] = pk;
        }
      }
    }
    if (s_end > s_begin) {
      umma::mbar_wait(done, 0);
      umma::fence_after_sync();
      // accumulator lane = (tap of the pair, ci), column = co
      const int l = q * 32 + lane, g = l >> 6, ci = l & 63;
      for (int p = 0; p < kPairs; ++p) {
        const int tap = 2 * p + g;
#pragma unroll
        for (int cg = 0; cg < CH / 32; ++cg) {
          float v[32];
          umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * CH + cg * 32), v);
          if (tap < 9) {
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(dW + ((size_t)(cg * 32 + i) * CH + ci) * 9 + tap, v[i]);
          }
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

}  // namespace

// X [rows, 64]; dY [rows, 64]; Wd: the weights packed for the input-gradient direction (pack_conv_weights(..., 1));
// dW [64][64][9] is ACCUMULATED into; dX [rows_pad, 64] is written.  sh_w: tap shifts of the weight gradient
// (make_shifts(P, 9, false)), sh_d: of the input gradient (mirrored).  *handled = false: not covered / switched off.
int conv_layer64_backward_fused(const __nv_bfloat16* X, const __nv_bfloat16* dY, const __nv_bfloat16* Wd, float* dW, __nv_bfloat16* dX,
                                long long rows, int P, int cin, int cout, const ConvShifts& sh_w, const ConvShifts& sh_d,
                                cudaStream_t st, bool* handled) {
  static const bool off = getenv("MIVIT_NO_LAYER64_FUSED") != nullptr;   // A/B switch
  *handled = false;
  if (off || cin != CH || cout != CH) return MIVIT_OK;
  constexpr int guard = 128;
  const int halo = P + 2;
  if (halo > guard - kBoxRows) return MIVIT_OK;
  const int slab_rows = (kStageRows + 2 * halo + kBoxRows - 1) / kBoxRows * kBoxRows;
  const int stage_bytes = 2 * slab_rows * 128;
  const int tail = kWBytes + (2 * kMaxRing + 5) * 8 + 16 + 64;
  int ring = (227 * 1024 - tail) / stage_bytes;
  if (ring > kMaxRing) ring = kMaxRing;
  if (ring < 2) return MIVIT_OK;
  const int smem = ring * stage_bytes + tail;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(conv_layer64_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long rows_pad = (rows + kStageRows - 1) / kStageRows * kStageRows;
  const int n_stages = (int)(rows_pad / kStageRows);
  CUtensorMap tmY, tmX;
  int rc = make_rows_tensor_map_sw(&tmY, dY - (size_t)guard * CH, CH, rows_pad + 2 * guard, kBoxRows);
  if (!rc) rc = make_rows_tensor_map_sw(&tmX, X - (size_t)guard * CH, CH, rows_pad + 2 * guard, kBoxRows);
  if (rc) return rc;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int ctas = sms < n_stages ? sms : n_stages;
  if (ctas < 1) { *handled = true; return MIVIT_OK; }
  const int spc = (n_stages + ctas - 1) / ctas;
  ctas = (n_stages + spc - 1) / spc;
  const double valid_rows = (double)rows * P * P / ((double)(P + 1) * (P + 1));
  MivitProfScope prof("conv_layer_bwd_64x64x9", 2.0 * 2.0 * valid_rows * 9 * CH * CH, st);
  conv_layer64_bwd_kernel<<<ctas, 256, smem, st>>>(tmY, tmX, Wd, dW, dX, n_stages, spc, sh_w, sh_d, halo, slab_rows, ring, guard);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  *handled = true;
  return MIVIT_OK;
}
