// 3x3 / 1x1 convolutions of DeepResNetEmbedding (reference helpers/models.py:202-257) as
// shifted-row implicit GEMMs on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators).
//
// Activation layout ("pitched rows", bf16, channels last): frame f, pixel (y,x) lives in row
//     r = f*(P+1)^2 + y*(P+1) + x
// of a [rows, C] matrix; column x = P of every line and line y = P of every frame are zero,
// and there are >= 128 zero guard rows on both ends.  A tap (dy,dx) of a stride-1 / pad-1
// convolution is then the pure row shift  delta = dy*(P+1) + dx  -- the zero column / line
// supply the padding -- so
//     Y[r, :] = sum_tap  X[r + delta_tap, :] * W_tap^T          (forward and, with the mirrored
//                                                               taps + transposed W, dgrad)
//     dW_tap  = sum_r    dY[r, :]^T * X[r + delta_tap, :]       (wgrad: rows are the K dimension)
// The row slab needed by one 128-row tile (128 + 2*halo rows) is staged in shared memory ONCE,
// in the no-swizzle core-matrix order [channel chunk of 8][row][8], and every tap reads it
// through a UMMA descriptor whose start address is advanced by delta*16 bytes: 9 taps reuse
// one slab (no im2col, no 9x re-read).  The same slab order is K-major for forward/dgrad
// (rows = M) and MN-major for wgrad (rows = K).
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kTileM = 128;

__device__ __forceinline__ bool row_valid(long long r, long long rows, int P) {
  if (r < 0 || r >= rows) return false;
  const int pitch = P + 1;
  const int q = (int)(r % (long long)(pitch * pitch));
  const int y = q / pitch, x = q - y * pitch;
  return y < P && x < P;
}

// -------------------------------------------------------------------------------------------
// forward / dgrad:  Y[rows, COUT] = sum_tap X[rows + delta, CIN] * Wp[tap]   (+ per-channel stats)
//   Wp: bf16 [taps][CIN/8][COUT][8]  (packed by pack kernels; K-major core-matrix order)
// One CTA = 128 threads, persistent over 128-row tiles.
// -------------------------------------------------------------------------------------------
template <int CIN, int COUT>
struct FwdCfg {
  static constexpr int kChunks = CIN / 8;
  static constexpr int kTapBytes = CIN * COUT * 2;
  static constexpr int kTapsPerStage = (96 * 1024) / kTapBytes >= 9 ? 9 : (96 * 1024) / kTapBytes;
  static constexpr int kStageBytes = kTapsPerStage * kTapBytes;
  static constexpr int kStagingBytes = kTileM * COUT * 2;
  static constexpr int kTmemCols = COUT <= 32 ? 32 : COUT <= 64 ? 64 : COUT <= 128 ? 128 : 256;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(128, 1)
conv_rows_tc_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ Wp,
                    __nv_bfloat16* __restrict__ Y, float* __restrict__ stats, long long rows, int n_tiles, int P,
                    int taps, ConvShifts shifts, int halo, int slab_rows) {
  using Cfg = FwdCfg<CIN, COUT>;
  constexpr int kSwz = (COUT / 8 - 1) < 7 ? (COUT / 8 - 1) : 7;  // XOR swizzle of the 16-byte chunks of a staging row
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slab_bytes = Cfg::kChunks * slab_rows * 16;
  uint8_t* slab = smem;
  uint8_t* wst = smem + ((slab_bytes + 127) & ~127);
  uint8_t* staging = wst + Cfg::kStageBytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  if (tid == 0) {
    umma::mbar_init(mbar, 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t idesc = umma::make_idesc_bf16(kTileM, COUT, 0, 0);
  const bool single_stage = taps <= Cfg::kTapsPerStage;
  uint32_t parity = 0;

  auto load_w_stage = [&](int tap0, int ntap) {
    const uint4* src = reinterpret_cast<const uint4*>(Wp + (size_t)tap0 * CIN * COUT);
    uint4* dst = reinterpret_cast<uint4*>(wst);
    const int n16 = ntap * Cfg::kTapBytes / 16;
    for (int i = tid; i < n16; i += 128) dst[i] = __ldg(src + i);
  };
  if (single_stage) load_w_stage(0, taps);

  float s_sum = 0.f, s_sq = 0.f;  // per-thread running channel statistics (thread -> column)

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long m0 = (long long)tile * kTileM;
    // ---- stage the row slab [m0-halo, m0-halo+slab_rows) x CIN as [chunk][row][8]
    {
      const uint4* src = reinterpret_cast<const uint4*>(X + (m0 - halo) * CIN);
      const int n16 = slab_rows * Cfg::kChunks;
      for (int i = tid; i < n16; i += 128) {
        const int r = i / Cfg::kChunks, c = i - r * Cfg::kChunks;
        *reinterpret_cast<uint4*>(slab + ((size_t)c * slab_rows + r) * 16) = __ldg(src + i);
      }
    }
    for (int tap0 = 0; tap0 < taps; tap0 += Cfg::kTapsPerStage) {
      const int ntap = min(Cfg::kTapsPerStage, taps - tap0);
      if (!single_stage) load_w_stage(tap0, ntap);
      umma::fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        umma::fence_after_sync();
        const uint32_t slab_addr = umma::smem_u32(slab), w_addr = umma::smem_u32(wst);
        for (int t = 0; t < ntap; ++t) {
          const int delta = shifts.d[tap0 + t];
#pragma unroll
          for (int j = 0; j < CIN / 16; ++j) {
            const uint64_t da = umma::make_desc(slab_addr + (uint32_t)((2 * j) * slab_rows + halo + delta) * 16u,
                                                (uint32_t)slab_rows * 16u, 128u);
            const uint64_t db = umma::make_desc(w_addr + (uint32_t)((t * Cfg::kChunks + 2 * j) * COUT) * 16u,
                                                (uint32_t)COUT * 16u, 128u);
            umma::mma_bf16(tmem, da, db, idesc, (tap0 + t > 0 || j > 0) ? 1u : 0u);
          }
        }
        umma::commit(mbar);
      }
      umma::mbar_wait(mbar, parity);
      parity ^= 1u;
    }
    umma::fence_after_sync();
    // ---- epilogue: TMEM -> registers -> (mask pads, bf16) -> swizzled staging tile
    {
      const long long r = m0 + warp * 32 + lane;
      const bool valid = row_valid(r, rows, P);
      const int rl = warp * 32 + lane;
#pragma unroll
      for (int cg = 0; cg < COUT / 32; ++cg) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cg * 32), v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 pk;
          uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = valid ? v[q * 8 + 2 * e] : 0.f, b = valid ? v[q * 8 + 2 * e + 1] : 0.f;
            __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
            pw[e] = *reinterpret_cast<uint32_t*>(&h);
          }
          const int chunk = cg * 4 + q;
          *reinterpret_cast<uint4*>(staging + ((size_t)rl * (COUT / 8) + (chunk ^ (rl & kSwz))) * 16) = pk;
        }
      }
    }
    umma::fence_before_sync();
    __syncthreads();
    // ---- per-channel sum / sum of squares of the stored (bf16-rounded) values
    if (stats != nullptr) {
      constexpr int kSplit = 128 / COUT >= 1 ? 128 / COUT : 1;  // threads per column
      constexpr int kColsPerThread = COUT > 128 ? COUT / 128 : 1;
      static_assert(kColsPerThread == 1, "COUT <= 128");
      const int col = tid % COUT, part = tid / COUT;
      if (part < kSplit) {
        const int r0 = part * (kTileM / kSplit), r1 = r0 + kTileM / kSplit;
        float a = 0.f, b = 0.f;
        for (int rr = r0; rr < r1; ++rr) {
          const __nv_bfloat16 h = *reinterpret_cast<const __nv_bfloat16*>(
              staging + ((size_t)rr * (COUT / 8) + ((col >> 3) ^ (rr & kSwz))) * 16 + (col & 7) * 2);
          const float f = __bfloat162float(h);
          a += f;
          b = fmaf(f, f, b);
        }
        s_sum += a;
        s_sq += b;
      }
    }
    // ---- coalesced copy of the tile (contiguous 128*COUT*2 bytes in Y)
    {
      uint4* dst = reinterpret_cast<uint4*>(Y + m0 * COUT);
      constexpr int n16 = kTileM * COUT / 8;
      for (int i = tid; i < n16; i += 128) {
        const int rr = i / (COUT / 8), c = i - rr * (COUT / 8);
        dst[i] = *reinterpret_cast<const uint4*>(staging + ((size_t)rr * (COUT / 8) + (c ^ (rr & kSwz))) * 16);
      }
    }
    __syncthreads();  // staging + slab are reused by the next tile
  }
  if (stats != nullptr && tid / COUT < (128 / COUT >= 1 ? 128 / COUT : 1)) {
    atomicAdd(stats + (tid % COUT), s_sum);
    atomicAdd(stats + COUT + (tid % COUT), s_sq);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<Cfg::kTmemCols>(tmem);
}

template <int CIN, int COUT>
int launch_fwd(const __nv_bfloat16* X, const __nv_bfloat16* Wp, __nv_bfloat16* Y, float* stats, long long rows, int P,
               int taps, const ConvShifts& sh, cudaStream_t st) {
  using Cfg = FwdCfg<CIN, COUT>;
  const int halo = P + 2;
  int slab_rows = kTileM + 2 * halo;
  if ((slab_rows & 1) == 0) ++slab_rows;  // odd row count -> conflict-free slab stores
  const int slab_bytes = ((Cfg::kChunks * slab_rows * 16) + 127) & ~127;
  const int smem = slab_bytes + Cfg::kStageBytes + Cfg::kStagingBytes + 64;
  MIVIT_CHECK_ARG(smem <= 227 * 1024, "conv tile needs %d bytes of shared memory", smem);
  auto kern = conv_rows_tc_kernel<CIN, COUT>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_tiles = (int)((rows + kTileM - 1) / kTileM);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = n_tiles < sms ? n_tiles : sms;
  char tag[48];
  snprintf(tag, sizeof(tag), "conv_rows_tc_%dx%dx%d", CIN, COUT, taps);
  const double valid_rows = (double)rows * P * P / ((double)(P + 1) * (P + 1));
  MivitProfScope prof(tag, 2.0 * valid_rows * taps * CIN * COUT, st);
  kern<<<grid, 128, smem, st>>>(X, Wp, Y, stats, rows, n_tiles, P, taps, sh, halo, slab_rows);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

// -------------------------------------------------------------------------------------------
// SIMT reference of the same contract (debug / cross-check only; selected with impl = 0)
// -------------------------------------------------------------------------------------------
__global__ void conv_rows_simt_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ Wp,
                                      __nv_bfloat16* __restrict__ Y, float* __restrict__ stats, long long rows,
                                      long long rows_pad, int P, int taps, ConvShifts sh, int CIN, int COUT) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_pad * COUT) return;
  const long long r = idx / COUT;
  const int co = (int)(idx - r * COUT);
  float acc = 0.f;
  if (row_valid(r, rows, P)) {
    for (int t = 0; t < taps; ++t) {
      const __nv_bfloat16* x = X + (r + sh.d[t]) * CIN;
      for (int ci = 0; ci < CIN; ++ci) {
        const float w = __bfloat162float(Wp[(((size_t)t * (CIN / 8) + (ci >> 3)) * COUT + co) * 8 + (ci & 7)]);
        acc = fmaf(__bfloat162float(x[ci]), w, acc);
      }
    }
  }
  const __nv_bfloat16 h = __float2bfloat16_rn(acc);
  Y[idx] = h;
  if (stats != nullptr && acc != 0.f) {
    const float f = __bfloat162float(h);
    atomicAdd(stats + co, f);
    atomicAdd(stats + COUT + co, f * f);
  }
}

}  // namespace

// MIVIT_NO_CTA_PAIRS=1 keeps the single-CTA TMA kernels for C_out = 128 (A/B measurements)
static bool use_cta_pairs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MIVIT_NO_CTA_PAIRS");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

int conv_rows_forward(const __nv_bfloat16* X, const __nv_bfloat16* Wp, __nv_bfloat16* Y, float* stats, long long rows,
                      int P, int cin, int cout, int taps, const ConvShifts& sh, int impl, cudaStream_t st) {
  MIVIT_CHECK_ARG(taps == 9 || taps == 1, "taps must be 1 or 9");
  MIVIT_CHECK_ARG(P + 2 <= 120, "patch size too large for the row-slab halo");
  if (impl == 0) {
    const long long rows_pad = (rows + kTileM - 1) / kTileM * kTileM;
    const long long n = rows_pad * cout;
    conv_rows_simt_kernel<<<mivit_ceil_div(n, 256), 256, 0, st>>>(X, Wp, Y, stats, rows, rows_pad, P, taps, sh, cin, cout);
    mivit_count_launch();
    MIVIT_LAUNCH_CHECK();
    return MIVIT_OK;
  }
  if (impl == 1) {  // pipelined kernel; falls through to the serial one if the slab does not fit
    bool handled = false;
    int rc = MIVIT_OK;
    if (use_cta_pairs()) {
      rc = conv_rows_forward_v4(X, Wp, nullptr, Y, nullptr, stats, nullptr, rows, P, cin, cout, taps, sh, st, &handled);   // CTA pairs
      if (rc || handled) return rc;
    }
    rc = conv_rows_forward_v3(X, Wp, nullptr, Y, nullptr, stats, nullptr, rows, P, cin, cout, taps, sh, st, &handled);   // TMA
    if (rc || handled) return rc;
  }
#define MIVIT_FWD_CASE(CI, CO) \
  if (cin == CI && cout == CO) return launch_fwd<CI, CO>(X, Wp, Y, stats, rows, P, taps, sh, st);
  MIVIT_FWD_CASE(32, 64)
  MIVIT_FWD_CASE(64, 64)
  MIVIT_FWD_CASE(64, 128)
  MIVIT_FWD_CASE(128, 128)
  MIVIT_FWD_CASE(64, 32)   // dgrad of 32->64
  MIVIT_FWD_CASE(128, 64)  // dgrad of 64->128
#undef MIVIT_FWD_CASE
  mivit_set_error("conv_rows_forward: unsupported channel pair %d -> %d", cin, cout);
  return MIVIT_ERR_INVALID;
}

// conv (taps) + the 1x1 skip convolution of the same input in one launch when the pipelined kernel
// covers the shape; otherwise two launches.
int conv_rows_forward_fused(const __nv_bfloat16* X, const __nv_bfloat16* Wp, const __nv_bfloat16* Wsk, __nv_bfloat16* Y,
                            __nv_bfloat16* Ysk, float* stats, float* stats_sk, long long rows, int P, int cin, int cout,
                            int taps, const ConvShifts& sh, int impl, cudaStream_t st) {
  if (impl == 1) {
    bool handled = false;
    int rc = MIVIT_OK;
    if (use_cta_pairs()) {
      rc = conv_rows_forward_v4(X, Wp, Wsk, Y, Ysk, stats, stats_sk, rows, P, cin, cout, taps, sh, st, &handled);   // CTA pairs
      if (rc || handled) return rc;
    }
    rc = conv_rows_forward_v3(X, Wp, Wsk, Y, Ysk, stats, stats_sk, rows, P, cin, cout, taps, sh, st, &handled);   // TMA
    if (rc || handled) return rc;
  }
  int rc = conv_rows_forward(X, Wp, Y, stats, rows, P, cin, cout, taps, sh, impl, st);
  if (rc) return rc;
  const ConvShifts s1 = make_shifts(P, 1, false);
  return conv_rows_forward(X, Wsk, Ysk, stats_sk, rows, P, cin, cout, 1, s1, impl, st);
}
