// Backward of ResidualBlock 1's first convolution and skip convolution (helpers/models.py:221-226) in ONE pass over their two
// output gradients dY1 = d raw1 (3x3 conv1) and dY2 = d raw_skip (1x1 skip), both [rows, 64], input X = act0 [rows, 32]:
//     dW1[co, ci, tap] = sum_r dY1[r, co] * X[r + delta_tap, ci]           (conv_wgrad3.cu)
//     dW2[co, ci]      = sum_r dY2[r, co] * X[r, ci]                       (conv_wgrad3.cu, ride-along)
//     dX[r, ci]        = sum_tap sum_co dY1[r - delta_tap, co] * W1[co, ci, tap] + sum_co dY2[r, co] * W2[co, ci]   (conv_tc3.cu, dual input)
// As two launches (weight gradients, then the dual-input dgrad) the two gradient tensors cross the TMA units twice; both kernels
// are bound by TMA row requests (~8.6 cycles per 64- or 128-byte row and SM), not by the tensor pipe, so
// sharing the rows is what pays: per 128-row stage 160 + 128 + 160 rows instead of (128 + 128 + 160) + (160 + 128).
// A stage holds the dY1 slab WITH its halo (the dgrad's taps are row shifts of it; the weight gradient reads its centre 128
// rows), the dY2 tile and the X slab with halo, each as the swizzled [row][128 B / 64 B] TMA tile that tcgen05 reads both as a
// K-major (rows = M: dgrad) and as an MN-major (rows = K: weight gradients) operand.
//   warp 0 producer | warps 1-3 weight-gradient issuers, one kernel row each (N = 3 taps x 32; warp 2 also the skip product) |
//   warp 8 dgrad issuer (two accumulators, even / odd stages) | warps 4-7 dgrad epilogue, then the flush of the ten weight-gradient
//   accumulators that live in TMEM over the CTA's whole row range.
// dX's pad rows are written with whatever the taps produce there: its only consumer, the stem BatchNorm's backward, reads valid
// pixels only.
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kStageRows = 128;
constexpr int kBoxRows = 32;
constexpr int kMaxRing = 4;
constexpr int CIN = 32, COUT = 64;
constexpr int kA2Bytes = kStageRows * 128;          // dY2 tile
constexpr int kW1Bytes = 9 * COUT * CIN * 2;        // packed dgrad weights [9][COUT/8][CIN][8]
constexpr int kW2Bytes = COUT * CIN * 2;
constexpr int kAccSkip = 9 * CIN;                   // TMEM columns: [0, 288) taps, [288, 320) skip, [320, 384) two dgrad accumulators
constexpr int kAccDx = kAccSkip + CIN;

__device__ __forceinline__ void mbar_arrive_cta(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(mbar)) : "memory");
}

__global__ void __launch_bounds__(288, 1)
conv_block1_bwd_kernel(const __grid_constant__ CUtensorMap tmY1, const __grid_constant__ CUtensorMap tmY2,
                       const __grid_constant__ CUtensorMap tmX, const __nv_bfloat16* __restrict__ Wd1, const __nv_bfloat16* __restrict__ Wd2,
                       float* __restrict__ dW1, float* __restrict__ dW2, __nv_bfloat16* __restrict__ dX, int n_stages, int stages_per_cta,
                       ConvShifts sh_w, ConvShifts sh_d, int halo, int slab_rows, int ring, int guard) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_uniform(), lane = tid & 31;
  const int a1_bytes = slab_rows * 128, b_bytes = slab_rows * 64;
  const int stage_bytes = a1_bytes + kA2Bytes + b_bytes;
  uint8_t* w1sm = smem + (size_t)ring * stage_bytes;       // behind the ring: also absorbs the M = 128 over-read of the last slot
  uint8_t* w2sm = w1sm + kW1Bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(w2sm + kW2Bytes);   // [kMaxRing]
  uint64_t* empty = full + kMaxRing;                               // [kMaxRing] three weight-gradient issuers + the dgrad issuer
  uint64_t* tfull = empty + kMaxRing;                              // [2]
  uint64_t* tempty = tfull + 2;                                    // [2]
  uint64_t* done = tempty + 2;                                     // [1] the three weight-gradient issuers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int s_begin = blockIdx.x * stages_per_cta;
  const int s_end = min(n_stages, s_begin + stages_per_cta);

  if (tid == 0) {
    for (int i = 0; i < kMaxRing; ++i) {
      umma::mbar_init(full + i, 1);
      umma::mbar_init(empty + i, 4);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(tfull + i, 1);
      umma::mbar_init(tempty + i, 4);
    }
    umma::mbar_init(done, 3);
    umma::mbar_fence_init();
    tma::prefetch_map(&tmY1);
    tma::prefetch_map(&tmY2);
    tma::prefetch_map(&tmX);
  }
  if (warp == 0) umma::tmem_alloc<512>(tmem_slot);
  for (int i = tid; i < kW1Bytes / 16; i += 288) reinterpret_cast<uint4*>(w1sm)[i] = __ldg(reinterpret_cast<const uint4*>(Wd1) + i);
  for (int i = tid; i < kW2Bytes / 16; i += 288) reinterpret_cast<uint4*>(w2sm)[i] = __ldg(reinterpret_cast<const uint4*>(Wd2) + i);
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
        uint8_t* a1 = smem + (size_t)slot * stage_bytes;
        uint8_t* a2 = a1 + a1_bytes;
        uint8_t* bs = a2 + kA2Bytes;
        tma::expect_tx(full + slot, (uint32_t)stage_bytes);
        const int r0 = guard + s * kStageRows;
        for (int b = 0; b < nb; ++b) tma::load_tile(a1 + (size_t)b * kBoxRows * 128, &tmY1, 0, r0 - halo + b * kBoxRows, full + slot);
#pragma unroll
        for (int b = 0; b < kStageRows / kBoxRows; ++b) tma::load_tile(a2 + (size_t)b * kBoxRows * 128, &tmY2, 0, r0 + b * kBoxRows, full + slot);
        for (int b = 0; b < nb; ++b) tma::load_tile(bs + (size_t)b * kBoxRows * 64, &tmX, 0, r0 - halo + b * kBoxRows, full + slot);
        if (++slot == ring) { slot = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 1 && warp <= 3) {
    // ===================== weight gradients: kernel row `iss`, D[co][(tap, ci)] += dY1^T X(shifted), both MN-major =====================
    const int iss = warp - 1;
    const uint32_t idesc = umma::make_idesc_bf16(128, 3 * CIN, 1, 1);
    const uint32_t idesc2 = umma::make_idesc_bf16(128, CIN, 1, 1);
    // dY tiles: M groups of 64 channels (the second group is the over-read of a 64-channel tile: accumulator lanes 64-127, never read)
    const uint64_t da0 = tma::make_desc_sw(smem0 + (uint32_t)(halo * 128), (uint32_t)kA2Bytes, 128u);
    const uint64_t da2 = tma::make_desc_sw(smem0 + (uint32_t)a1_bytes, (uint32_t)kA2Bytes, 128u);
    // X slab (64-byte rows, SWIZZLE_64B): LBO = one row = the next tap of the kernel row
    const uint64_t db0 = tma::make_desc_sw(smem0 + (uint32_t)(a1_bytes + kA2Bytes + halo * 64), 64u, 64u);
    const uint32_t a_hi = (uint32_t)(da0 >> 32), a2_hi = (uint32_t)(da2 >> 32), b_hi = (uint32_t)(db0 >> 32);
    const int d_first = sh_w.d[iss * 3];
    const uint32_t acc = tmem + (uint32_t)(iss * 3 * CIN);
    int slot = 0;
    uint32_t ph = 0;
    bool first = true;
    for (int s = s_begin; s < s_end; ++s) {
      umma::mbar_wait(full + slot, ph);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t a_lo0 = (uint32_t)da0 + (uint32_t)slot * stage_units;
        const uint32_t a2_lo0 = (uint32_t)da2 + (uint32_t)slot * stage_units;
        const uint32_t b_lo0 = (uint32_t)db0 + (uint32_t)slot * stage_units;
#pragma unroll
        for (int kk = 0; kk < kStageRows / 16; ++kk) {
          const uint32_t b_kk = b_lo0 + (uint32_t)(kk * 16 * 4);                      // 16 rows x 64 B
          umma::mma_bf16(acc, ((uint64_t)a_hi << 32) | (a_lo0 + (uint32_t)(kk * 128)), ((uint64_t)b_hi << 32) | (b_kk + (uint32_t)(d_first * 4)),
                         idesc, (!first || kk > 0) ? 1u : 0u);
          if (iss == 1)                                                               // skip product: dY2 against the unshifted rows
            umma::mma_bf16(tmem + (uint32_t)kAccSkip, ((uint64_t)a2_hi << 32) | (a2_lo0 + (uint32_t)(kk * 128)), ((uint64_t)b_hi << 32) | b_kk,
                           idesc2, (!first || kk > 0) ? 1u : 0u);
        }
        umma::commit(empty + slot);
        if (s == s_end - 1) umma::commit(done);
      }
      __syncwarp();
      first = false;
      if (++slot == ring) { slot = 0; ph ^= 1; }
    }
  } else if (warp == 8) {
    // ===================== input gradient: D[r][ci] = sum_tap dY1[r + shift] W1[tap] + dY2[r] W2, dY tiles K-major (rows = M) =====================
    constexpr uint32_t idesc = umma::make_idesc_bf16(128, CIN, 0, 0);
    const uint64_t da0 = tma::make_desc_sw(smem0 + (uint32_t)(halo * 128), 0u, 128u);
    const uint64_t da2 = tma::make_desc_sw(smem0 + (uint32_t)a1_bytes, 0u, 128u);
    const uint64_t db1 = umma::make_desc(umma::smem_u32(w1sm), (uint32_t)CIN * 16u, 128u);
    const uint64_t db2 = umma::make_desc(umma::smem_u32(w2sm), (uint32_t)CIN * 16u, 128u);
    const uint32_t a_hi = (uint32_t)(da0 >> 32), a2_hi = (uint32_t)(da2 >> 32), b1_hi = (uint32_t)(db1 >> 32), b2_hi = (uint32_t)(db2 >> 32);
    int dl[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) dl[t] = sh_d.d[t] * 8;       // rows -> 16-byte units of a 128-byte row
    int slot = 0;
    uint32_t ph = 0;
    int k = 0;
    for (int s = s_begin; s < s_end; ++s, ++k) {
      const int buf = k & 1;
      umma::mbar_wait(full + slot, ph);
      umma::mbar_wait(tempty + buf, ((k >> 1) & 1) ^ 1);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t acc = tmem + (uint32_t)(kAccDx + buf * CIN);
        const uint32_t a_lo0 = (uint32_t)da0 + (uint32_t)slot * stage_units;
        const uint32_t a2_lo0 = (uint32_t)da2 + (uint32_t)slot * stage_units;
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int j = 0; j < COUT / 16; ++j)
            umma::mma_bf16(acc, ((uint64_t)a_hi << 32) | (a_lo0 + (uint32_t)dl[t] + (uint32_t)(2 * j)),
                           ((uint64_t)b1_hi << 32) | ((uint32_t)db1 + (uint32_t)((t * (COUT / 8) + 2 * j) * CIN)), idesc, (t > 0 || j > 0) ? 1u : 0u);
#pragma unroll
        for (int j = 0; j < COUT / 16; ++j)
          umma::mma_bf16(acc, ((uint64_t)a2_hi << 32) | (a2_lo0 + (uint32_t)(2 * j)), ((uint64_t)b2_hi << 32) | ((uint32_t)db2 + (uint32_t)(2 * j * CIN)),
                         idesc, 1u);
        umma::commit(empty + slot);
        umma::commit(tfull + buf);
      }
      __syncwarp();
      if (++slot == ring) { slot = 0; ph ^= 1; }
    }
  } else if (warp >= 4 && warp <= 7) {
    // ===================== dgrad epilogue, then the weight-gradient flush =====================
    const int q = warp - 4;
    int k = 0;
    for (int s = s_begin; s < s_end; ++s, ++k) {
      const int buf = k & 1;
      umma::mbar_wait(tfull + buf, (k >> 1) & 1);
      umma::fence_after_sync();
      float v[32];
      umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(kAccDx + buf * CIN), v);
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(tempty + buf);
      uint4* out = reinterpret_cast<uint4*>(dX + ((size_t)s * kStageRows + q * 32 + lane) * CIN);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        uint4 pk;
        uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(v[c4 * 8 + 2 * e], v[c4 * 8 + 2 * e + 1]);
          pw[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        out[c4] = pk;
      }
    }
    if (s_end > s_begin) {
      umma::mbar_wait(done, 0);
      umma::fence_after_sync();
      if (q * 32 < COUT) {
        const int co = q * 32 + lane;
        for (int t = 0; t < 9; ++t) {
          float v[32];
          umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * CIN), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(dW1 + ((size_t)co * CIN + i) * 9 + t, v[i]);
        }
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)kAccSkip, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(dW2 + (size_t)co * CIN + i, v[i]);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tmem);
}

}  // namespace

// X = act0 [rows, 32]; dY1 / dY2 [rows, 64]; Wd1 / Wd2: conv1 / skip weights packed for the input-gradient direction
// (pack_conv_weights(..., 1)); dW1 [64][32][9] and dW2 [64][32] are ACCUMULATED into; dX [rows_pad, 32] is written.
// sh_w: tap shifts of the weight gradient (make_shifts(P, 9, false)), sh_d: of the input gradient (mirrored).
// *handled = false: shape not covered or switched off, nothing was launched.
int conv_block1_backward_fused(const __nv_bfloat16* X, const __nv_bfloat16* dY1, const __nv_bfloat16* dY2, const __nv_bfloat16* Wd1,
                               const __nv_bfloat16* Wd2, float* dW1, float* dW2, __nv_bfloat16* dX, long long rows, int P, int cin,
                               int cout, const ConvShifts& sh_w, const ConvShifts& sh_d, cudaStream_t st, bool* handled) {
  static const bool off = getenv("MIVIT_NO_BLOCK1_FUSED") != nullptr;   // A/B switch
  *handled = false;
  if (off || cin != CIN || cout != COUT) return MIVIT_OK;
  constexpr int guard = 128;
  const int halo = P + 2;
  if (halo > guard - kBoxRows) return MIVIT_OK;
  const int slab_rows = (kStageRows + 2 * halo + kBoxRows - 1) / kBoxRows * kBoxRows;
  const int stage_bytes = slab_rows * 128 + kA2Bytes + slab_rows * 64;
  const int tail = kW1Bytes + kW2Bytes + (2 * kMaxRing + 5) * 8 + 16 + 64;
  int ring = (227 * 1024 - tail) / stage_bytes;
  if (ring > kMaxRing) ring = kMaxRing;
  if (ring < 2 || kW1Bytes < kA2Bytes) return MIVIT_OK;   // (the weights behind the ring absorb the last slot's 16 KB over-read)
  const int smem = ring * stage_bytes + tail;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(conv_block1_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long rows_pad = (rows + kStageRows - 1) / kStageRows * kStageRows;
  const int n_stages = (int)(rows_pad / kStageRows);
  CUtensorMap tmY1, tmY2, tmX;
  int rc = make_rows_tensor_map_sw(&tmY1, dY1 - (size_t)guard * COUT, COUT, rows_pad + 2 * guard, kBoxRows);
  if (!rc) rc = make_rows_tensor_map_sw(&tmY2, dY2 - (size_t)guard * COUT, COUT, rows_pad + 2 * guard, kBoxRows);
  if (!rc) rc = make_rows_tensor_map_sw(&tmX, X - (size_t)guard * CIN, CIN, rows_pad + 2 * guard, kBoxRows);
  if (rc) return rc;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int ctas = sms < n_stages ? sms : n_stages;
  if (ctas < 1) { *handled = true; return MIVIT_OK; }
  const int spc = (n_stages + ctas - 1) / ctas;
  ctas = (n_stages + spc - 1) / spc;
  const double valid_rows = (double)rows * P * P / ((double)(P + 1) * (P + 1));
  MivitProfScope prof("conv_block1_bwd_32x64", 2.0 * 2.0 * valid_rows * 10 * CIN * COUT, st);
  conv_block1_bwd_kernel<<<ctas, 288, smem, st>>>(tmY1, tmY2, tmX, Wd1, Wd2, dW1, dW2, dX, n_stages, spc, sh_w, sh_d, halo, slab_rows, ring,
                                                  guard);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  *handled = true;
  return MIVIT_OK;
}
