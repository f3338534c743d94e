// Pipelined weight-gradient kernel (see conv_wgrad.cu for the formulation):
//     dW[co, ci, tap] = sum_r dY[r, co] * X[r + delta_tap, ci]      (rows = K, both operands MN-major)
// Double-buffered (dY, X) row stages filled with cp.async by 7 producer warps; one thread issues the
// tcgen05.mma chain of a stage and releases the stage buffer with tcgen05.commit -> mbarrier, so the
// loads of stage s+1 overlap the MMAs of stage s.  Accumulators stay in TMEM over the CTA's whole
// row range and are flushed once with fp32 atomics into the [co][ci][kh][kw] gradient.
#include "common.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kStageRows = 128;
constexpr int kProducerWarps = 7;

__device__ __forceinline__ void cp_async16w(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(umma::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_allw() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrivew(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(mbar)) : "memory");
}

template <int CIN, int COUT>
struct WgCfg2 {
  static constexpr int kTapsPerCta = 3;
  static constexpr int kCols = kTapsPerCta * CIN;
  static constexpr int kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128 : kCols <= 256 ? 256 : 512;
  static constexpr int kAChunks = COUT / 8, kBChunks = CIN / 8;
  static constexpr int kABytes = kAChunks * kStageRows * 16;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(320, 1)
conv_wgrad_tc2_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ dY, float* __restrict__ dW,
                      int n_stages, int stages_per_cta, int taps, ConvShifts shifts, int halo, int xslab_rows, int buf_bytes) {
  using C = WgCfg2<CIN, COUT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_uniform(), lane = tid & 31;
  // buffer b: [A slab (dY) | B slab (X)]; the M = 128 over-read of a 64-channel A slab lands in the B slab
  uint8_t* buf0 = smem;
  uint8_t* buf1 = smem + buf_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * buf_bytes);
  uint64_t* full = bars;       // [2]
  uint64_t* empty = bars + 2;  // [2]
  uint64_t* done = bars + 4;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tap0 = blockIdx.y * C::kTapsPerCta;
  const int ntap = min(C::kTapsPerCta, taps - tap0);
  const int s_begin = blockIdx.x * stages_per_cta;
  const int s_end = min(n_stages, s_begin + stages_per_cta);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(full + i, kProducerWarps);
      umma::mbar_init(empty + i, ntap);
    }
    umma::mbar_init(done, ntap);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<C::kTmemCols>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp != 4 && warp < 8) {
    // ===================== producers (warps 0-3, 5-7) =====================
    const int pw = warp < 4 ? warp : warp - 1;
    const int pt = pw * 32 + lane;
    int k = 0;
    for (int s = s_begin; s < s_end; ++s, ++k) {
      const int b = k & 1;
      umma::mbar_wait(empty + b, ((k >> 1) & 1) ^ 1);
      uint8_t* aslab = b ? buf1 : buf0;
      uint8_t* bslab = aslab + C::kABytes;
      const long long r0 = (long long)s * kStageRows;
      const uint4* srca = reinterpret_cast<const uint4*>(dY + r0 * COUT);
      for (int i = pt; i < kStageRows * C::kAChunks; i += kProducerWarps * 32) {
        const int r = i / C::kAChunks, c = i - r * C::kAChunks;
        cp_async16w(aslab + ((size_t)c * kStageRows + r) * 16, srca + i);
      }
      const uint4* srcb = reinterpret_cast<const uint4*>(X + (r0 - halo) * CIN);
      for (int i = pt; i < xslab_rows * C::kBChunks; i += kProducerWarps * 32) {
        const int r = i / C::kBChunks, c = i - r * C::kBChunks;
        cp_async16w(bslab + ((size_t)c * xslab_rows + r) * 16, srcb + i);
      }
      if (s + 2 < s_end) {  // L2 prefetch two stages ahead (see conv_tc2.cu)
        const char* pa = reinterpret_cast<const char*>(dY + (r0 + 2 * kStageRows) * COUT);
        const char* pb = reinterpret_cast<const char*>(X + (r0 + 2 * kStageRows - halo) * CIN);
        const int la = kStageRows * COUT * 2 / 128, lb = (xslab_rows * CIN * 2 + 127) / 128;
        for (int i = pt; i < la + lb; i += kProducerWarps * 32)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(i < la ? pa + (size_t)i * 128 : pb + (size_t)(i - la) * 128));
      }
      cp_async_wait_allw();
      umma::fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrivew(full + b);
    }
  } else {
    // ===================== MMA issuers: warps 4, 8, 9 -> tap 0, 1, 2 of this CTA's tap group.
    // One issuing thread per tap (separate TMEM accumulators) because a single thread's scalar
    // stream cannot keep the tensor pipe busy for N <= 64.
    constexpr uint32_t idesc = umma::make_idesc_bf16(128, CIN, 1, 1);
    const int t = warp == 4 ? 0 : warp - 7;
    if (t < ntap) {
      const uint64_t da_b[2] = {umma::make_desc(umma::smem_u32(buf0), 128u, (uint32_t)kStageRows * 16u),
                                umma::make_desc(umma::smem_u32(buf1), 128u, (uint32_t)kStageRows * 16u)};
      const uint64_t db_b[2] = {umma::make_desc(umma::smem_u32(buf0) + C::kABytes + (uint32_t)halo * 16u, 128u, (uint32_t)xslab_rows * 16u),
                                umma::make_desc(umma::smem_u32(buf1) + C::kABytes + (uint32_t)halo * 16u, 128u, (uint32_t)xslab_rows * 16u)};
      const uint32_t a_hi = (uint32_t)(da_b[0] >> 32), b_hi = (uint32_t)(db_b[0] >> 32);
      const int dl = shifts.d[tap0 + t];
      const uint32_t acc = tmem + (uint32_t)(t * CIN);
      int k = 0;
      for (int s = s_begin; s < s_end; ++s, ++k) {
        const int b = k & 1;
        umma::mbar_wait(full + b, (k >> 1) & 1);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          const uint32_t a_lo0 = (uint32_t)da_b[b], bt = (uint32_t)db_b[b] + (uint32_t)dl;
#pragma unroll
          for (int kk = 0; kk < kStageRows / 16; ++kk) {
            const uint64_t da = ((uint64_t)a_hi << 32) | (a_lo0 + (uint32_t)(kk * 16));
            const uint64_t db = ((uint64_t)b_hi << 32) | (bt + (uint32_t)(kk * 16));
            umma::mma_bf16(acc, da, db, idesc, (k > 0 || kk > 0) ? 1u : 0u);
          }
          umma::commit(empty + b);
          if (s == s_end - 1) umma::commit(done);
        }
        __syncwarp();
      }
    }
  }
  // ===================== flush (warps 0-3): lane = co, columns = (tap, ci) =====================
  if (warp < 4 && s_end > s_begin) {
    umma::mbar_wait(done, 0);
    umma::fence_after_sync();
    if (warp * 32 < COUT) {
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
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<C::kTmemCols>(tmem);
}

template <int CIN, int COUT>
int launch_wgrad2(const __nv_bfloat16* X, const __nv_bfloat16* dY, float* dW, long long rows, int P, int taps,
                  const ConvShifts& sh, cudaStream_t st, bool* fits) {
  using C = WgCfg2<CIN, COUT>;
  const int halo = taps == 1 ? 0 : P + 2;
  int xslab_rows = kStageRows + 2 * halo;
  if ((xslab_rows & 1) == 0) ++xslab_rows;
  int buf_bytes = C::kABytes + C::kBChunks * xslab_rows * 16;
  const int need_a = 16 * kStageRows * 16;  // M = 128 over-read of the A slab must stay inside the buffer
  if (buf_bytes < need_a) buf_bytes = need_a;
  buf_bytes = (buf_bytes + 127) & ~127;
  int smem = 2 * buf_bytes + 128;
  *fits = smem <= 227 * 1024;
  if (!*fits) return MIVIT_OK;
  if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM (TMEM budget)
  auto kern = conv_wgrad_tc2_kernel<CIN, COUT>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long rows_pad = (rows + kStageRows - 1) / kStageRows * kStageRows;
  const int n_stages = (int)(rows_pad / kStageRows);
  const int groups = (taps + C::kTapsPerCta - 1) / C::kTapsPerCta;
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
  kern<<<dim3(ctas_x, groups), 320, smem, st>>>(X, dY, dW, n_stages, spc, taps, sh, halo, xslab_rows, buf_bytes);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

}  // namespace

int conv_rows_wgrad_v2(const __nv_bfloat16* X, const __nv_bfloat16* dY, float* dW, long long rows, int P, int cin, int cout,
                       int taps, const ConvShifts& sh, cudaStream_t st, bool* handled) {
  *handled = true;
  bool fits = true;
  int rc;
#define WG2_CASE(CI, CO)                                                        \
  if (cin == CI && cout == CO) {                                                \
    rc = launch_wgrad2<CI, CO>(X, dY, dW, rows, P, taps, sh, st, &fits);        \
    if (!fits) *handled = false;                                                \
    return rc;                                                                  \
  }
  WG2_CASE(32, 64)
  WG2_CASE(64, 64)
  WG2_CASE(64, 128)
  WG2_CASE(128, 128)
#undef WG2_CASE
  *handled = false;
  return MIVIT_OK;
}
