// Pipelined, warp-specialised version of the shifted-row implicit-GEMM convolution (see
// conv_tc.cu for the layout and the algorithm).  Differences from the serial v1 kernel:
//   * the packed weights of the CTA's output-column slice stay RESIDENT in shared memory for the
//     whole (persistent) kernel -- for 128->128 the output columns are split over two CTAs so that
//     9*128*64*2 B = 144 KB fits -- instead of being re-streamed from L2 for every 128-row tile;
//   * row slabs are double buffered and filled with cp.async by three producer warps;
//   * one thread issues the tcgen05.mma chain of a tile into one of two TMEM accumulators and
//     signals completion with tcgen05.commit -> mbarrier, so the MMAs of tile k+1 overlap the
//     epilogue (TMEM -> registers -> bf16 -> global, BatchNorm statistics) of tile k;
//   * the 1x1 skip convolution of a ResidualBlock (helpers/models.py:216,221) reads the same input
//     as conv1, so it is fused: a second accumulator fed by the centre tap of the same slab.
// Warp roles (288 threads): warps 0-3 epilogue (TMEM lane quarter = warp id), warps 4 and 8 MMA
// issuers (even / odd tiles), warps 5-7 producers.
#include "common.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

constexpr int kTileM = 128;
constexpr int kProducerWarps = 3;

__device__ __forceinline__ bool row_valid2(long long r, long long rows, int P) {
  if (r < 0 || r >= rows) return false;
  const int pitch = P + 1;
  const int q = (int)(r % (long long)(pitch * pitch));
  const int y = q / pitch, x = q - y * pitch;
  return y < P && x < P;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(umma::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(mbar)) : "memory");
}

// Column sums over the 32 lanes of a warp for 32 per-lane values: after the butterfly, lane l
// holds sum_lanes v[l].  31 shuffles instead of 32*5.
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <int CIN, int NS, bool SKIP>
struct Cfg2 {
  static constexpr int kCH = CIN / 8;
  static constexpr int kWSkipBytes = SKIP ? CIN * NS * 2 : 0;
  static constexpr int kAccCols = NS * (SKIP ? 2 : 1);
  static constexpr int kTmemCols = 2 * kAccCols <= 32 ? 32 : 2 * kAccCols <= 64 ? 64 : 2 * kAccCols <= 128 ? 128
                                   : 2 * kAccCols <= 256 ? 256 : 512;
  static constexpr int kGroups = NS / 32;
};

template <int CIN, int NS, bool SKIP, int TAPS>
__global__ void __launch_bounds__(288, 1)
conv_rows_tc2_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ Wp,
                     const __nv_bfloat16* __restrict__ Wsk, __nv_bfloat16* __restrict__ Y, __nv_bfloat16* __restrict__ Ysk,
                     float* __restrict__ stats, float* __restrict__ stats_sk, long long rows, int n_tiles, int P,
                     ConvShifts shifts, int halo, int slab_rows, int cout_total, int nsplit) {
  using C = Cfg2<CIN, NS, SKIP>;
  constexpr int taps = TAPS;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = umma::warp_idx_uniform(), lane = tid & 31;
  const int slab_bytes = (C::kCH * slab_rows * 16 + 127) & ~127;
  uint8_t* wsm = smem;                                   // [taps][CH][NS][8]
  uint8_t* wsk = wsm + taps * CIN * NS * 2;              // [CH][NS][8]           (SKIP)
  uint8_t* slab0 = wsk + C::kWSkipBytes;
  uint8_t* slab1 = slab0 + slab_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(slab1 + slab_bytes);
  uint64_t* full = bars;        // [2] producers -> MMA
  uint64_t* empty = bars + 2;   // [2] MMA (commit) -> producers
  uint64_t* tfull = bars + 4;   // [2] MMA (commit) -> epilogue
  uint64_t* tempty = bars + 6;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int nh = blockIdx.x % nsplit;                    // output-column slice of this CTA
  const int cta_in_slice = blockIdx.x / nsplit, ctas_per_slice = gridDim.x / nsplit;
  const int col0 = nh * NS;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(full + i, kProducerWarps);
      umma::mbar_init(empty + i, 1);
      umma::mbar_init(tfull + i, 1);
      umma::mbar_init(tempty + i, 4);
    }
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc<C::kTmemCols>(tmem_slot);
  // resident weights of this column slice
  {
    const int per_tap = C::kCH * NS;  // 16-byte units per tap in smem
    for (int i = tid; i < taps * per_tap; i += 288) {
      const int t = i / per_tap, r = i - t * per_tap, ch = r / NS, n = r - ch * NS;
      const uint4* src = reinterpret_cast<const uint4*>(Wp) + ((size_t)(t * C::kCH + ch) * cout_total + col0 + n);
      reinterpret_cast<uint4*>(wsm)[i] = __ldg(src);
    }
    if (SKIP) {
      for (int i = tid; i < per_tap; i += 288) {
        const int ch = i / NS, n = i - ch * NS;
        const uint4* src = reinterpret_cast<const uint4*>(Wsk) + ((size_t)ch * cout_total + col0 + n);
        reinterpret_cast<uint4*>(wsk)[i] = __ldg(src);
      }
    }
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 5 && warp <= 7) {
    // ===================== producers: row slab of tile k -> slab[k & 1] =====================
    const int pt = (warp - 5) * 32 + lane;
    int k = 0;
    for (int tile = cta_in_slice; tile < n_tiles; tile += ctas_per_slice, ++k) {
      const int buf = k & 1;
      umma::mbar_wait(empty + buf, ((k >> 1) & 1) ^ 1);
      uint8_t* slab = buf ? slab1 : slab0;
      const long long m0 = (long long)tile * kTileM;
      const uint4* src = reinterpret_cast<const uint4*>(X + (m0 - halo) * CIN);
      const int n16 = slab_rows * C::kCH;
      for (int i = pt; i < n16; i += kProducerWarps * 32) {
        const int r = i / C::kCH, c = i - r * C::kCH;
        cp_async16(slab + ((size_t)c * slab_rows + r) * 16, src + i);
      }
      // L2 prefetch of the slab two work items ahead: only two slabs fit next to the resident weights, so
      // the cp.async of a slab cannot be issued earlier than one item ahead; pulling the lines into L2
      // now turns its HBM latency (~1 us) into an L2 hit.
      {
        const int tile2 = tile + 2 * ctas_per_slice;
        if (tile2 < n_tiles) {
          const char* p2 = reinterpret_cast<const char*>(X + ((long long)tile2 * kTileM - halo) * CIN);
          const int lines = (slab_rows * CIN * 2 + 127) / 128;
          for (int i = pt; i < lines; i += kProducerWarps * 32)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p2 + (size_t)i * 128));
        }
      }
      cp_async_wait_all();
      umma::fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full + buf);
    }
  } else if (warp == 4 || warp == 8) {
    // ===================== MMA issuers: warp 4 -> even work items (buffer 0), warp 8 -> odd (buffer 1).
    // Two issuing threads because one thread's scalar stream (~17 SASS instructions per tcgen05.mma)
    // is slower than the tensor pipe for N = 64 tiles; their MMAs target different accumulators.
    // The issue loop is a single thread's scalar instruction stream, so it is kept to ~2 integer
    // adds per tcgen05.mma: descriptors are (constant high word | start-address low word) and only
    // the low word is advanced (tap shift: +delta rows, K step: +2 chunks).
    constexpr uint32_t idesc = umma::make_idesc_bf16(kTileM, NS, 0, 0);
    const uint64_t da_base[2] = {umma::make_desc(umma::smem_u32(slab0) + (uint32_t)halo * 16u, (uint32_t)slab_rows * 16u, 128u),
                                 umma::make_desc(umma::smem_u32(slab1) + (uint32_t)halo * 16u, (uint32_t)slab_rows * 16u, 128u)};
    const uint64_t db_base = umma::make_desc(umma::smem_u32(wsm), (uint32_t)NS * 16u, 128u);
    const uint64_t dbsk_base = umma::make_desc(umma::smem_u32(wsk), (uint32_t)NS * 16u, 128u);
    const uint32_t a_hi = (uint32_t)(da_base[0] >> 32), b_hi = (uint32_t)(db_base >> 32);
    const uint32_t a_kstep = 2u * (uint32_t)slab_rows;  // two 8-channel chunks per K = 16 step (16-byte units)
    int dl[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) dl[t] = shifts.d[t];
    int k = 0;
    const int my_buf = warp == 4 ? 0 : 1;
    for (int tile = cta_in_slice; tile < n_tiles; tile += ctas_per_slice, ++k) {
      const int buf = k & 1;
      if (buf != my_buf) continue;
      const uint32_t ph = (k >> 1) & 1;
      umma::mbar_wait(full + buf, ph);
      umma::mbar_wait(tempty + buf, ph ^ 1);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t a_lo0 = (uint32_t)da_base[buf];
        const uint32_t b_lo0 = (uint32_t)db_base;
        const uint32_t acc = tmem + (uint32_t)(buf * C::kAccCols);
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
          uint32_t a_lo = a_lo0 + (uint32_t)dl[t];
#pragma unroll
          for (int j = 0; j < CIN / 16; ++j) {
            const uint64_t da = ((uint64_t)a_hi << 32) | a_lo;
            const uint64_t db = ((uint64_t)b_hi << 32) | (b_lo0 + (uint32_t)((t * C::kCH + 2 * j) * NS));
            umma::mma_bf16(acc, da, db, idesc, (t > 0 || j > 0) ? 1u : 0u);
            a_lo += a_kstep;
          }
        }
        if (SKIP) {
          uint32_t a_lo = a_lo0;
          const uint32_t bs_lo0 = (uint32_t)dbsk_base;
#pragma unroll
          for (int j = 0; j < CIN / 16; ++j) {
            const uint64_t da = ((uint64_t)a_hi << 32) | a_lo;
            const uint64_t db = ((uint64_t)b_hi << 32) | (bs_lo0 + (uint32_t)((2 * j) * NS));
            umma::mma_bf16(acc + NS, da, db, idesc, j > 0 ? 1u : 0u);
            a_lo += a_kstep;
          }
        }
        umma::commit(empty + buf);   // slab may be refilled once these MMAs have read it
        umma::commit(tfull + buf);   // accumulator ready for the epilogue
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps 0-3 =====================
    float ssum[SKIP ? 2 : 1][C::kGroups], ssq[SKIP ? 2 : 1][C::kGroups];
#pragma unroll
    for (int o = 0; o < (SKIP ? 2 : 1); ++o)
#pragma unroll
      for (int g = 0; g < C::kGroups; ++g) ssum[o][g] = ssq[o][g] = 0.f;
    int k = 0;
    for (int tile = cta_in_slice; tile < n_tiles; tile += ctas_per_slice, ++k) {
      const int buf = k & 1;
      umma::mbar_wait(tfull + buf, (k >> 1) & 1);
      umma::fence_after_sync();
      const long long r = (long long)tile * kTileM + warp * 32 + lane;
      const bool valid = row_valid2(r, rows, P);
      const uint32_t acc = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * C::kAccCols);
#pragma unroll
      for (int o = 0; o < (SKIP ? 2 : 1); ++o) {
        __nv_bfloat16* dst = (o == 0 ? Y : Ysk) + r * cout_total + col0;
        const bool want_stats = (o == 0 ? stats : stats_sk) != nullptr;
#pragma unroll
        for (int g = 0; g < C::kGroups; ++g) {
          float v[32];
          umma::tmem_ld32(acc + (uint32_t)(o * NS + g * 32), v);
          if (o == (SKIP ? 1 : 0) && g == C::kGroups - 1) {
            // last TMEM read of this accumulator: hand it back to the MMA warp
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + buf);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 pk;
            uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = valid ? v[q * 8 + 2 * e] : 0.f, b = valid ? v[q * 8 + 2 * e + 1] : 0.f;
              __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
              pw[e] = *reinterpret_cast<uint32_t*>(&h);
              const float2 back = __bfloat1622float2(h);
              v[q * 8 + 2 * e] = back.x;       // statistics of the STORED values
              v[q * 8 + 2 * e + 1] = back.y;
            }
            *reinterpret_cast<uint4*>(dst + g * 32 + q * 8) = pk;
          }
          if (want_stats) {
            float sq[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) sq[i] = v[i] * v[i];
            ssum[o][g] += warp_transpose_reduce(v, lane);
            ssq[o][g] += warp_transpose_reduce(sq, lane);
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < (SKIP ? 2 : 1); ++o) {
      float* st = o == 0 ? stats : stats_sk;
      if (st != nullptr && k > 0) {
#pragma unroll
        for (int g = 0; g < C::kGroups; ++g) {
          atomicAdd(st + col0 + g * 32 + lane, ssum[o][g]);
          atomicAdd(st + cout_total + col0 + g * 32 + lane, ssq[o][g]);
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<C::kTmemCols>(tmem);
}

template <int CIN, int NS, bool SKIP, int TAPS>
int launch2(const __nv_bfloat16* X, const __nv_bfloat16* Wp, const __nv_bfloat16* Wsk, __nv_bfloat16* Y, __nv_bfloat16* Ysk,
            float* stats, float* stats_sk, long long rows, int P, int cout_total, const ConvShifts& sh,
            cudaStream_t st, bool* fits) {
  using C = Cfg2<CIN, NS, SKIP>;
  constexpr int taps = TAPS;
  const int halo = P + 2;
  int slab_rows = kTileM + 2 * halo;
  if ((slab_rows & 1) == 0) ++slab_rows;
  const int slab_bytes = (C::kCH * slab_rows * 16 + 127) & ~127;
  int smem = taps * CIN * NS * 2 + C::kWSkipBytes + 2 * slab_bytes + 128;
  *fits = smem <= 227 * 1024;
  if (!*fits) return MIVIT_OK;
  if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM: the TMEM budget assumes it
  auto kern = conv_rows_tc2_kernel<CIN, NS, SKIP, TAPS>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_tiles = (int)((rows + kTileM - 1) / kTileM);
  const int nsplit = cout_total / NS;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int per_slice = sms / nsplit;
  if (per_slice > n_tiles) per_slice = n_tiles;
  if (per_slice < 1) per_slice = 1;
  char tag[48];
  snprintf(tag, sizeof(tag), "conv_rows_tc_%dx%dx%d%s", CIN, cout_total, taps, SKIP ? "+skip" : "");
  const double valid_rows = (double)rows * P * P / ((double)(P + 1) * (P + 1));
  MivitProfScope prof(tag, 2.0 * valid_rows * (taps + (SKIP ? 1 : 0)) * CIN * cout_total, st);
  kern<<<per_slice * nsplit, 288, smem, st>>>(X, Wp, Wsk, Y, Ysk, stats, stats_sk, rows, n_tiles, P, sh, halo, slab_rows,
                                              cout_total, nsplit);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

}  // namespace

// Returns MIVIT_OK and sets *handled = false when the configuration is not covered by the
// pipelined kernel (the caller then falls back to the serial v1 kernel).
int conv_rows_forward_v2(const __nv_bfloat16* X, const __nv_bfloat16* Wp, const __nv_bfloat16* Wsk, __nv_bfloat16* Y,
                         __nv_bfloat16* Ysk, float* stats, float* stats_sk, long long rows, int P, int cin, int cout, int taps,
                         const ConvShifts& sh, cudaStream_t st, bool* handled) {
  const bool skip = Wsk != nullptr;
  *handled = true;
  bool fits = true;
  int rc = MIVIT_OK;
#define V2_CASE(CI, CO, NS_, SK)                                                                                       \
  if (cin == CI && cout == CO && skip == SK) {                                                                         \
    rc = taps == 9 ? launch2<CI, NS_, SK, 9>(X, Wp, Wsk, Y, Ysk, stats, stats_sk, rows, P, cout, sh, st, &fits)        \
                   : launch2<CI, NS_, SK, 1>(X, Wp, Wsk, Y, Ysk, stats, stats_sk, rows, P, cout, sh, st, &fits);       \
    if (!fits) *handled = false;                                                                                       \
    return rc;                                                                                                         \
  }
  V2_CASE(32, 64, 64, true)
  V2_CASE(32, 64, 64, false)
  V2_CASE(64, 64, 64, false)
  V2_CASE(64, 128, 128, true)
  V2_CASE(64, 128, 128, false)
  V2_CASE(128, 128, 64, false)
  V2_CASE(64, 32, 32, false)
  V2_CASE(128, 64, 64, false)
#undef V2_CASE
  *handled = false;
  return MIVIT_OK;
}
