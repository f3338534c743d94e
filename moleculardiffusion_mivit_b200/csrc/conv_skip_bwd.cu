// Backward of a ResidualBlock's 1x1 skip convolution (helpers/models.py:221-226: identity = skip_bn(skip_conv(x))) in ONE pass
// over its output gradient dY [rows, C_out]:
//     dW[co, ci]  = sum_r dY[r, co] * X[r, ci]          (weight gradient, conv_wgrad3.cu with taps = 1)
//     dX[r, ci]   = sum_co dY[r, co] * W[co, ci]        (input gradient, conv_tc3.cu with taps = 1)
// Both products are pure streams over dY (1.5 GB at 128 channels and 1024 x 30 frames of 13 x 13): as two launches they read it
// twice (0.54 ms at 4.3 TB/s + 0.37 ms at 6.3 TB/s).  Here a stage of 128 rows of dY lands ONCE in shared memory as the swizzled
// [row][128 B] TMA tile that tcgen05 reads both ways: as an MN-major operand (rows = K) for dW and as a
// K-major operand (rows = M) for dX.
//   warp 0: TMA producer (ring of (dY stage, X stage) pairs);  warp 1: dW issuer (accumulator lives in TMEM over the CTA's whole
//   row range, flushed once with fp32 atomics);  warp 2: dX issuer (two accumulators, even / odd stages);  warps 4-7: dX epilogue
//   (TMEM -> bf16 -> 128 contiguous bytes per row and thread) and the final dW flush.
// Pad rows of dY are zeros (written by the BatchNorm backward), so dX's pad rows come out as zeros without masking.
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kStageRows = 128;
constexpr int kBoxRows = 32;
constexpr int kMaxRing = 8;

__device__ __forceinline__ void mbar_arrive_local(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(mbar)) : "memory");
}

template <int CIN, int COUT>
struct SkipCfg {
  static_assert(CIN == 64 && (COUT == 64 || COUT == 128), "shapes of the reference's residual blocks");
  static constexpr int kARegions = COUT / 64;
  static constexpr int kARegionBytes = kStageRows * 128;
  static constexpr int kABytes = kARegions * kARegionBytes;     // dY stage
  static constexpr int kBBytes = kStageRows * 128;              // X stage (64 channels)
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kWBytes = COUT * CIN * 2;                // packed dgrad weights [COUT/8][CIN][8]
  static constexpr int kOverRead = COUT < 128 ? kARegionBytes : 0;   // M = 128 MMA on a 64-channel dY tile (dW product)
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(256, 1)
conv_skip_bwd_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                     const __nv_bfloat16* __restrict__ Wp, float* __restrict__ dW, __nv_bfloat16* __restrict__ dX, int n_stages,
                     int stages_per_cta, int ring, int guard) {
  using C = SkipCfg<CIN, COUT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_uniform(), lane = tid & 31;
  uint8_t* wsm = smem + (size_t)ring * C::kStageBytes + C::kOverRead;   // K-major, no swizzle: [COUT/8][CIN][8]
  uint64_t* full = reinterpret_cast<uint64_t*>(wsm + C::kWBytes);      // [kMaxRing] TMA -> both issuers
  uint64_t* empty = full + kMaxRing;                                    // [kMaxRing] both issuers (commit) -> TMA
  uint64_t* tfull = empty + kMaxRing;                                   // [2] dX accumulator ready
  uint64_t* tempty = tfull + 2;                                         // [2] dX accumulator read
  uint64_t* done = tempty + 2;                                          // [1] dW accumulator complete
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
  if (warp == 0) umma::tmem_alloc<256>(tmem_slot);
  for (int i = tid; i < C::kWBytes / 16; i += 256) reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(Wp) + i);
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t stage_units = (uint32_t)C::kStageBytes >> 4;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (umma::elect_one()) {
      int slot = 0;
      uint32_t ph = 0;
      for (int s = s_begin; s < s_end; ++s) {
        umma::mbar_wait(empty + slot, ph ^ 1);
        uint8_t* aslab = smem + (size_t)slot * C::kStageBytes;
        uint8_t* bslab = aslab + C::kABytes;
        tma::expect_tx(full + slot, (uint32_t)C::kStageBytes);
        const int r0 = guard + s * kStageRows;
#pragma unroll
        for (int reg = 0; reg < C::kARegions; ++reg)
#pragma unroll
          for (int rb = 0; rb < kStageRows / kBoxRows; ++rb)
            tma::load_tile(aslab + (size_t)reg * C::kARegionBytes + (size_t)rb * kBoxRows * 128, &tmY, reg * 64, r0 + rb * kBoxRows,
                           full + slot);
#pragma unroll
        for (int rb = 0; rb < kStageRows / kBoxRows; ++rb)
          tma::load_tile(bslab + (size_t)rb * kBoxRows * 128, &tmX, 0, r0 + rb * kBoxRows, full + slot);
        if (++slot == ring) { slot = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== dW issuer: D[co][ci] += dY^T X, both operands MN-major (rows = K) =====================
    const uint32_t idesc = umma::make_idesc_bf16(128, CIN, 1, 1);
    const uint64_t da0 = tma::make_desc_sw(umma::smem_u32(smem), (uint32_t)C::kARegionBytes, 128u);
    const uint64_t db0 = tma::make_desc_sw(umma::smem_u32(smem) + C::kABytes, 128u, 128u);
    const uint32_t a_hi = (uint32_t)(da0 >> 32), b_hi = (uint32_t)(db0 >> 32);
    int slot = 0;
    uint32_t ph = 0;
    bool first = true;
    for (int s = s_begin; s < s_end; ++s) {
      umma::mbar_wait(full + slot, ph);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t a_lo0 = (uint32_t)da0 + (uint32_t)slot * stage_units;
        const uint32_t b_lo0 = (uint32_t)db0 + (uint32_t)slot * stage_units;
#pragma unroll
        for (int kk = 0; kk < kStageRows / 16; ++kk)      // 16 rows x 128 B per K step
          umma::mma_bf16(tmem, ((uint64_t)a_hi << 32) | (a_lo0 + (uint32_t)(kk * 128)), ((uint64_t)b_hi << 32) | (b_lo0 + (uint32_t)(kk * 128)),
                         idesc, (!first || kk > 0) ? 1u : 0u);
        umma::commit(empty + slot);
        if (s == s_end - 1) umma::commit(done);
      }
      __syncwarp();
      first = false;
      if (++slot == ring) { slot = 0; ph ^= 1; }
    }
  } else if (warp == 2) {
    // ===================== dX issuer: D[r][ci] = dY W, dY tile as the K-major operand (rows = M), weights K-major =====================
    constexpr uint32_t idesc = umma::make_idesc_bf16(128, CIN, 0, 0);
    const uint64_t da0 = tma::make_desc_sw(umma::smem_u32(smem), 0u, 128u);
    const uint64_t db0 = umma::make_desc(umma::smem_u32(wsm), (uint32_t)CIN * 16u, 128u);
    const uint32_t a_hi = (uint32_t)(da0 >> 32), b_hi = (uint32_t)(db0 >> 32);
    int slot = 0;
    uint32_t ph = 0;
    int k = 0;
    for (int s = s_begin; s < s_end; ++s, ++k) {
      const int buf = k & 1;
      umma::mbar_wait(full + slot, ph);
      umma::mbar_wait(tempty + buf, ((k >> 1) & 1) ^ 1);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t acc = tmem + 64u + (uint32_t)(buf * 64);
        const uint32_t a_lo0 = (uint32_t)da0 + (uint32_t)slot * stage_units;
#pragma unroll
        for (int reg = 0; reg < C::kARegions; ++reg)
#pragma unroll
          for (int j = 0; j < 4; ++j) {                 // 16 of the region's 64 channels per MMA: +32 B inside the swizzled row
            const uint64_t da = ((uint64_t)a_hi << 32) | (a_lo0 + (uint32_t)(reg * (C::kARegionBytes >> 4)) + (uint32_t)(2 * j));
            const uint64_t db = ((uint64_t)b_hi << 32) | ((uint32_t)db0 + (uint32_t)((reg * 8 + 2 * j) * CIN));
            umma::mma_bf16(acc, da, db, idesc, (reg > 0 || j > 0) ? 1u : 0u);
          }
        umma::commit(empty + slot);
        umma::commit(tfull + buf);
      }
      __syncwarp();
      if (++slot == ring) { slot = 0; ph ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== dX epilogue, then the dW flush =====================
    const int q = warp - 4;
    int k = 0;
    for (int s = s_begin; s < s_end; ++s, ++k) {
      const int buf = k & 1;
      umma::mbar_wait(tfull + buf, (k >> 1) & 1);
      umma::fence_after_sync();
      const uint32_t acc = tmem + ((uint32_t)(q * 32) << 16) + 64u + (uint32_t)(buf * 64);
      uint4* out = reinterpret_cast<uint4*>(dX + ((size_t)s * kStageRows + q * 32 + lane) * CIN);
#pragma unroll
      for (int g = 0; g < CIN / 32; ++g) {
        float v[32];
        umma::tmem_ld32(acc + (uint32_t)(g * 32), v);
        if (g == CIN / 32 - 1) {      // last TMEM read of this accumulator: hand it back to the issuer
          umma::fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_local(tempty + buf);
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
      if (q * 32 < COUT) {
        const int co = q * 32 + lane;
#pragma unroll
        for (int cg = 0; cg < CIN / 32; ++cg) {
          float v[32];
          umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 32), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(dW + (size_t)co * CIN + cg * 32 + i, v[i]);
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<256>(tmem);
}

template <int CIN, int COUT>
int launch_skip_bwd(const __nv_bfloat16* X, const __nv_bfloat16* dY, const __nv_bfloat16* Wp, float* dW, __nv_bfloat16* dX,
                    long long rows, int P, cudaStream_t st) {
  using C = SkipCfg<CIN, COUT>;
  constexpr int guard = 128;
  const int tail = C::kOverRead + C::kWBytes + (2 * kMaxRing + 5) * 8 + 16 + 64;
  int ring = (227 * 1024 - tail) / C::kStageBytes;
  if (ring > kMaxRing) ring = kMaxRing;
  const int smem = ring * C::kStageBytes + tail;
  auto kern = conv_skip_bwd_kernel<CIN, COUT>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long rows_pad = (rows + kStageRows - 1) / kStageRows * kStageRows;
  const int n_stages = (int)(rows_pad / kStageRows);
  CUtensorMap tmY, tmX;
  int rc = make_rows_tensor_map_sw(&tmY, dY - (size_t)guard * COUT, COUT, rows_pad + 2 * guard, kBoxRows);
  if (rc) return rc;
  rc = make_rows_tensor_map_sw(&tmX, X - (size_t)guard * CIN, CIN, rows_pad + 2 * guard, kBoxRows);
  if (rc) return rc;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int ctas = sms < n_stages ? sms : n_stages;
  if (ctas < 1) return MIVIT_OK;
  const int spc = (n_stages + ctas - 1) / ctas;
  ctas = (n_stages + spc - 1) / spc;
  char tag[48];
  snprintf(tag, sizeof(tag), "conv_skip_bwd_%dx%dx1", CIN, COUT);
  const double valid_rows = (double)rows * P * P / ((double)(P + 1) * (P + 1));
  MivitProfScope prof(tag, 2.0 * 2.0 * valid_rows * CIN * COUT, st);
  kern<<<ctas, 256, smem, st>>>(tmY, tmX, Wp, dW, dX, n_stages, spc, ring, guard);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

}  // namespace

// dW [cout][cin] is ACCUMULATED into (the caller zeroes the gradient buffer); dX rows [0, rows_pad) are written.
// Wp: the skip weights packed for the input-gradient direction (pack_conv_weights(..., transpose = 1): [cout/8][cin][8] bf16).
// *handled = false: shape not covered here, nothing was launched.
int conv_skip_backward_fused(const __nv_bfloat16* X, const __nv_bfloat16* dY, const __nv_bfloat16* Wp, float* dW, __nv_bfloat16* dX,
                             long long rows, int P, int cin, int cout, cudaStream_t st, bool* handled) {
  static const bool off = getenv("MIVIT_NO_SKIP_FUSED") != nullptr;   // A/B switch
  *handled = false;
  if (off) return MIVIT_OK;
  if (cin == 64 && cout == 128) {
    *handled = true;
    return launch_skip_bwd<64, 128>(X, dY, Wp, dW, dX, rows, P, st);
  }
  return MIVIT_OK;
}
