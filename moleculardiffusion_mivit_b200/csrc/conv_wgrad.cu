// Weight gradient of the shifted-row convolutions on tcgen05:
//     dW[co, ci, tap] = sum_r dY[r, co] * X[r + delta_tap, ci]
// The reduction dimension is the ROW index, so both operands are read "MN-major" from the same
// [channel chunk][row][8] shared-memory slabs the forward kernel uses (rows 16 B apart inside a
// chunk = the K direction of an MN-major core matrix) -- no transposes are materialised.
//   A = dY^T : M = co (128 lanes; for C_out = 64 the upper 64 lanes accumulate garbage that is
//              never read), B = X shifted by the tap : N = ci.
// Each CTA owns one group of <= 3 taps (3*N <= 384 TMEM columns) and a contiguous range of
// 128-row stages, accumulates in TMEM over the whole range and flushes once with fp32 atomics
// straight into the nn.Conv2d weight-gradient layout [co][ci][kh][kw].
#include "common.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kStageRows = 128;

template <int CIN, int COUT>
struct WgCfg {
  static constexpr int kTapsPerCta = 3;
  static constexpr int kCols = kTapsPerCta * CIN;
  static constexpr int kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128 : kCols <= 256 ? 256 : 512;
  static constexpr int kAChunks = COUT / 8, kBChunks = CIN / 8;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(128, 1)
conv_wgrad_tc_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ dY, float* __restrict__ dW,
                     long long rows_pad, int n_stages, int stages_per_cta, int taps, ConvShifts shifts, int halo,
                     int xslab_rows) {
  using Cfg = WgCfg<CIN, COUT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A slab first; it is over-read up to 16 chunks (M = 128) when COUT = 64, which lands in the B slab.
  uint8_t* aslab = smem;                                              // [COUT/8][128][16 B]
  uint8_t* bslab = smem + Cfg::kAChunks * kStageRows * 16;            // [CIN/8][xslab_rows][16 B]
  const int bbytes = Cfg::kBChunks * xslab_rows * 16;
  const int tail = ((Cfg::kAChunks * kStageRows * 16 + bbytes + 127) & ~127);
  // (allocation is padded on the host so that aslab + 16 chunks stays in bounds)
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + tail);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int tap0 = blockIdx.y * Cfg::kTapsPerCta;
  const int ntap = min(Cfg::kTapsPerCta, taps - tap0);
  const int s_begin = blockIdx.x * stages_per_cta;
  const int s_end = min(n_stages, s_begin + stages_per_cta);

  if (tid == 0) {
    umma::mbar_init(mbar, 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t idesc = umma::make_idesc_bf16(128, CIN, 1, 1);  // both operands MN-major
  uint32_t parity = 0;

  for (int s = s_begin; s < s_end; ++s) {
    const long long r0 = (long long)s * kStageRows;
    {  // dY rows [r0, r0+128)
      const uint4* src = reinterpret_cast<const uint4*>(dY + r0 * COUT);
      for (int i = tid; i < kStageRows * Cfg::kAChunks; i += 128) {
        const int r = i / Cfg::kAChunks, c = i - r * Cfg::kAChunks;
        *reinterpret_cast<uint4*>(aslab + ((size_t)c * kStageRows + r) * 16) = __ldg(src + i);
      }
    }
    {  // X rows [r0-halo, r0-halo+xslab_rows)
      const uint4* src = reinterpret_cast<const uint4*>(X + (r0 - halo) * CIN);
      for (int i = tid; i < xslab_rows * Cfg::kBChunks; i += 128) {
        const int r = i / Cfg::kBChunks, c = i - r * Cfg::kBChunks;
        *reinterpret_cast<uint4*>(bslab + ((size_t)c * xslab_rows + r) * 16) = __ldg(src + i);
      }
    }
    umma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      const uint32_t a_addr = umma::smem_u32(aslab), b_addr = umma::smem_u32(bslab);
      for (int t = 0; t < ntap; ++t) {
        const int delta = shifts.d[tap0 + t];
#pragma unroll
        for (int kk = 0; kk < kStageRows / 16; ++kk) {
          // MN-major: LBO = stride between 8-row K groups (128 B), SBO = stride between 8-channel chunks
          const uint64_t da = umma::make_desc(a_addr + (uint32_t)(kk * 16) * 16u, 128u, (uint32_t)kStageRows * 16u);
          const uint64_t db = umma::make_desc(b_addr + (uint32_t)(halo + delta + kk * 16) * 16u, 128u,
                                              (uint32_t)xslab_rows * 16u);
          umma::mma_bf16(tmem + (uint32_t)(t * CIN), da, db, idesc, (s > s_begin || kk > 0) ? 1u : 0u);
        }
      }
      umma::commit(mbar);
    }
    umma::mbar_wait(mbar, parity);
    parity ^= 1u;
  }
  umma::fence_after_sync();
  // flush: lane (row of D) = co, columns = (tap, ci)
  if (s_end > s_begin && warp * 32 < COUT) {
    const int co = warp * 32 + lane;
    for (int t = 0; t < ntap; ++t) {
#pragma unroll
      for (int cg = 0; cg < CIN / 32; ++cg) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * CIN + cg * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(dW + ((size_t)co * CIN + cg * 32 + i) * taps + tap0 + t, v[i]);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<Cfg::kTmemCols>(tmem);
}

template <int CIN, int COUT>
int launch_wgrad(const __nv_bfloat16* X, const __nv_bfloat16* dY, float* dW, long long rows, int P, int taps,
                 const ConvShifts& sh, cudaStream_t st) {
  using Cfg = WgCfg<CIN, COUT>;
  const int halo = taps == 1 ? 0 : P + 2;
  int xslab_rows = kStageRows + 2 * halo;
  if ((xslab_rows & 1) == 0) ++xslab_rows;
  const int abytes = Cfg::kAChunks * kStageRows * 16, bbytes = Cfg::kBChunks * xslab_rows * 16;
  int smem = ((abytes + bbytes + 127) & ~127) + 64;
  const int need_a = 16 * kStageRows * 16 + 64;  // M = 128 over-read of the A slab
  if (smem < need_a) smem = need_a;
  auto kern = conv_wgrad_tc_kernel<CIN, COUT>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long rows_pad = (rows + kStageRows - 1) / kStageRows * kStageRows;
  const int n_stages = (int)(rows_pad / kStageRows);
  const int groups = (taps + Cfg::kTapsPerCta - 1) / Cfg::kTapsPerCta;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int ctas_x = sms / groups;
  if (ctas_x < 1) ctas_x = 1;
  if (ctas_x > n_stages) ctas_x = n_stages;
  const int spc = (n_stages + ctas_x - 1) / ctas_x;
  ctas_x = (n_stages + spc - 1) / spc;
  char tag[48];
  snprintf(tag, sizeof(tag), "conv_wgrad_tc_%dx%dx%d", CIN, COUT, taps);
  const double valid_rows = (double)rows * P * P / ((double)(P + 1) * (P + 1));
  MivitProfScope prof(tag, 2.0 * valid_rows * taps * CIN * COUT, st);
  kern<<<dim3(ctas_x, groups), 128, smem, st>>>(X, dY, dW, rows_pad, n_stages, spc, taps, sh, halo, xslab_rows);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

// SIMT cross-check: one CTA per (tap, co); threads over ci; loops over all rows.
__global__ void conv_wgrad_simt_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ dY,
                                       float* __restrict__ dW, long long rows, int taps, ConvShifts sh, int CIN, int COUT) {
  const int t = blockIdx.x, co = blockIdx.y;
  for (int ci = threadIdx.x; ci < CIN; ci += blockDim.x) {
    float acc = 0.f;
    for (long long r = 0; r < rows; ++r) {
      const float g = __bfloat162float(dY[r * COUT + co]);
      if (g != 0.f) acc = fmaf(g, __bfloat162float(X[(r + sh.d[t]) * CIN + ci]), acc);
    }
    dW[((size_t)co * CIN + ci) * taps + t] += acc;
  }
}

}  // namespace

// dW (fp32, [cout][cin][k][k]) += ...; the caller zeroes dW first.
int conv_rows_wgrad(const __nv_bfloat16* X, const __nv_bfloat16* dY, float* dW, long long rows, int P, int cin, int cout,
                    int taps, const ConvShifts& sh, int impl, cudaStream_t st) {
  MIVIT_CHECK_ARG(taps == 9 || taps == 1, "taps must be 1 or 9");
  if (impl == 0) {
    conv_wgrad_simt_kernel<<<dim3(taps, cout), 128, 0, st>>>(X, dY, dW, rows, taps, sh, cin, cout);
    mivit_count_launch();
    MIVIT_LAUNCH_CHECK();
    return MIVIT_OK;
  }
  if (impl == 1) {  // pipelined kernel; falls back to the serial one when the stage does not fit
    bool handled = false;
    int rc = conv_rows_wgrad_v3(X, dY, dW, rows, P, cin, cout, taps, sh, st, &handled);   // TMA ring + cluster multicast
    if (rc || handled) return rc;
  }
#define MIVIT_WG_CASE(CI, CO) \
  if (cin == CI && cout == CO) return launch_wgrad<CI, CO>(X, dY, dW, rows, P, taps, sh, st);
  MIVIT_WG_CASE(32, 64)
  MIVIT_WG_CASE(64, 64)
  MIVIT_WG_CASE(64, 128)
  MIVIT_WG_CASE(128, 128)
#undef MIVIT_WG_CASE
  mivit_set_error("conv_rows_wgrad: unsupported channel pair %d -> %d", cin, cout);
  return MIVIT_ERR_INVALID;
}
