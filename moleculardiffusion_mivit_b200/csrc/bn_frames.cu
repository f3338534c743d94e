// Frame-at-a-time BatchNorm passes of the LAST residual block, fed by TMA.
//
// The two BatchNorm passes around the global average pool (helpers/models.py:226-227,240,254: out = relu(bn2(raw_a) + bn_skip(raw_b)),
// pooled = mean over the P x P pixels) work on whole frames: the backward's upstream gradient is one vector per FRAME
// (dpooled[f, :] / P^2), the forward produces one vector per frame.  The row-streaming kernels of bn.cu pay for that with a
// row -> (frame, y, x) decomposition and a per-row look-up of the frame's vector next to every 16-byte load (bn_bwd_apply<128,0,1>:
// 4.6 TB/s = 70 % of the copy peak, L1 hit rate 63 %, profiles/r02_ncu_bn.md).  Here a persistent CTA takes one frame at a time:
//   * a producer thread requests the frame's VALID pixels of both raw tensors with one 4-D TMA box each ({C, P, P, 1} of the
//     (channel, x, y, frame) view of the pitched-rows layout: pad rows / columns are never read, no address arithmetic) plus the
//     frame's vector as a 1-D bulk copy, two frames ahead of the consumers (double buffer, mbarrier full / empty pairs);
//   * 8 consumer warps walk the frame in shared memory (thread = 8-channel chunk x pixel lane, per-channel constants in
//     registers), write the results with coalesced 16-byte stores and zero the frame's pad row / column.
// Same arithmetic, operation for operation, as bn.cu (bn_pre, the coefficient form k1 g + k2 raw + k3): bit-identical outputs.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"
#include "vit.h"

namespace {

// (channel, x, y, frame) view of a pitched-rows bf16 tensor [frames * (P+1)^2, C]; box = the P x P valid pixels of one frame
int make_frame_map(CUtensorMap* tm, const void* row0, int C, int P, long long n_frames) {
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)(P + 1), (cuuint64_t)(P + 1), (cuuint64_t)n_frames};
  const cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)(P + 1) * C * 2, (cuuint64_t)(P + 1) * (P + 1) * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)P, (cuuint32_t)P, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  mivit_tensor_map_encode_fn encode = mivit_tensor_map_encoder();
  if (encode == nullptr) return MIVIT_ERR_CUDA;
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(row0), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mivit_set_error("cuTensorMapEncodeTiled failed (%d) for the frame view C=%d P=%d frames=%lld", (int)r, C, P, n_frames);
    return MIVIT_ERR_CUDA;
  }
  return MIVIT_OK;
}

__device__ __forceinline__ void load_frame(void* smem_dst, const CUtensorMap* tm, int frame, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          umma::smem_u32(smem_dst)),
      "l"(tm), "r"(0), "r"(0), "r"(0), "r"(frame), "r"(umma::smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void load_bytes(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(umma::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ void load8f(const float* __restrict__ p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

constexpr int kConsumers = 256;   // 8 warps; warp 8 = the TMA producer

// draw_a = k1a gm + k2a raw_a + k3a,  draw_b = k1b gm + k2b raw_b + k3b,  gm = dpooled[f] / P^2 * 1[bn_a(raw_a) + bn_b(raw_b) > 0]
// (bn.cu: bn_bwd_apply_kernel<C, 0, true>); pad rows / columns of both outputs are written as zeros.
template <int C>
__global__ void __launch_bounds__(kConsumers + 32, 1)
bn_bwd_apply_pooled_frames_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                  const float* __restrict__ dpooled, const float* __restrict__ ss_a, const float* __restrict__ coef_a,
                                  __nv_bfloat16* __restrict__ draw_a, const float* __restrict__ ss_b, const float* __restrict__ coef_b,
                                  __nv_bfloat16* __restrict__ draw_b, int n_frames, int P, long long rows, long long rows_pad) {
  constexpr int CH = C / 8, NPL = kConsumers / CH;
  extern __shared__ __align__(128) uint8_t smem[];
  const int PP = P * P, pitch = P + 1, rpf = pitch * pitch;
  const int FB = PP * C * 2;                                   // bytes of a frame's valid pixels
  uint8_t* bufA = smem;                                        // [2][FB]
  uint8_t* bufB = smem + 2 * FB;                               // [2][FB]
  float* dp = reinterpret_cast<float*>(smem + 4 * FB);         // [2][C]
  uint64_t* full = reinterpret_cast<uint64_t*>(dp + 2 * C);    // [2]
  uint64_t* empty = full + 2;                                  // [2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(full + i, 1);
      umma::mbar_init(empty + i, kConsumers / 32);
    }
    umma::mbar_fence_init();
    tma::prefetch_map(&tmA);
    tma::prefetch_map(&tmB);
  }
  __syncthreads();
  if (warp == kConsumers / 32) {
    if (lane == 0) {
      int k = 0;
      for (int f = blockIdx.x; f < n_frames; f += gridDim.x, ++k) {
        const int b = k & 1;
        umma::mbar_wait(empty + b, ((k >> 1) & 1) ^ 1);
        tma::expect_tx(full + b, (uint32_t)(2 * FB + C * 4));
        load_frame(bufA + (size_t)b * FB, &tmA, f, full + b);
        load_frame(bufB + (size_t)b * FB, &tmB, f, full + b);
        load_bytes(dp + b * C, dpooled + (size_t)f * C, C * 4, full + b);
      }
    }
    return;
  }
  const int ch = tid % CH, pl = tid / CH;
  float sa[8], ha[8], sb[8], hb[8], k1a[8], k2a[8], k3a[8], k1b[8], k2b[8], k3b[8];
  load8f(ss_a + ch * 8, sa); load8f(ss_a + C + ch * 8, ha);
  load8f(ss_b + ch * 8, sb); load8f(ss_b + C + ch * 8, hb);
  load8f(coef_a + ch * 8, k1a); load8f(coef_a + C + ch * 8, k2a); load8f(coef_a + 2 * C + ch * 8, k3a);
  load8f(coef_b + ch * 8, k1b); load8f(coef_b + C + ch * 8, k2b); load8f(coef_b + 2 * C + ch * 8, k3b);
  const float inv_pp = 1.0f / (float)PP;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  uint4* outA = reinterpret_cast<uint4*>(draw_a);
  uint4* outB = reinterpret_cast<uint4*>(draw_b);
  const int y0 = pl / P, x0 = pl - y0 * P;
  int k = 0;
  for (int f = blockIdx.x; f < n_frames; f += gridDim.x, ++k) {
    const int b = k & 1;
    const long long frow = (long long)f * rpf;
    // the pad column (x = P, every line) and the pad line (y = P) of both outputs: zeros, independent of the loads
    for (int i = tid; i < (2 * P + 1) * CH; i += kConsumers) {
      const int pix = i / CH, c = i - pix * CH;
      const int y = pix <= P ? pix : P, x = pix <= P ? P : pix - pitch;
      const long long idx = (frow + y * pitch + x) * CH + c;
      outA[idx] = zero;
      outB[idx] = zero;
    }
    umma::mbar_wait(full + b, (k >> 1) & 1);
    float g[8];
    {
      const float4 g0 = *reinterpret_cast<const float4*>(dp + b * C + ch * 8), g1 = *reinterpret_cast<const float4*>(dp + b * C + ch * 8 + 4);
      g[0] = g0.x * inv_pp; g[1] = g0.y * inv_pp; g[2] = g0.z * inv_pp; g[3] = g0.w * inv_pp;
      g[4] = g1.x * inv_pp; g[5] = g1.y * inv_pp; g[6] = g1.z * inv_pp; g[7] = g1.w * inv_pp;
    }
    const uint4* fa = reinterpret_cast<const uint4*>(bufA + (size_t)b * FB);
    const uint4* fb = reinterpret_cast<const uint4*>(bufB + (size_t)b * FB);
    int y = y0, x = x0;
    for (int p = pl; p < PP; p += NPL) {
      const uint4 ua = fa[p * CH + ch], ub = fb[p * CH + ch];
      float a[8], bb[8], oa[8], ob[8];
      unpack8(ua, a);
      unpack8(ub, bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float pre = fmaf(a[i], sa[i], ha[i]);                 // == bn.cu bn_pre<true>
        pre += fmaf(bb[i], sb[i], hb[i]);
        const float gg = pre > 0.f ? g[i] : 0.f;
        oa[i] = fmaf(k1a[i], gg, fmaf(k2a[i], a[i], k3a[i]));
        ob[i] = fmaf(k1b[i], gg, fmaf(k2b[i], bb[i], k3b[i]));
      }
      const long long idx = (frow + y * pitch + x) * CH + ch;
      outA[idx] = pack8(oa);
      outB[idx] = pack8(ob);
      x += NPL;
      while (x >= P) { x -= P; ++y; }
    }
    __syncwarp();
    if (lane == 0) arrive(empty + b);     // this warp is done with the buffer
  }
  // rows between the last frame and the 128-row tile boundary
  if (blockIdx.x == 0)
    for (long long i = rows * CH + tid; i < rows_pad * CH; i += kConsumers) { outA[i] = zero; outB[i] = zero; }
}

// pooled[f, c] = mean over the P x P pixels of relu(bn_a(raw_a) + bn_b(raw_b)) and, for the backward of this BatchNorm pair, the
// per-frame masked sums fsums[f] = [n+ | sum mask * raw_a | sum mask * raw_b] (bn.cu: bn_apply_pool_kernel).
template <int C>
__global__ void __launch_bounds__(kConsumers + 32, 1)
bn_apply_pool_frames_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const float* __restrict__ ss_a, const float* __restrict__ ss_b, float* __restrict__ pooled,
                            float* __restrict__ fsums, int n_frames, int P) {
  constexpr int CH = C / 8, NPL = kConsumers / CH, LPW = 32 / CH;   // pixel lanes per CTA / per warp
  extern __shared__ __align__(128) uint8_t smem[];
  const int PP = P * P;
  const int FB = PP * C * 2;
  uint8_t* bufA = smem;
  uint8_t* bufB = smem + 2 * FB;
  float* red = reinterpret_cast<float*>(smem + 4 * FB);        // [8 warps][4 sums][C]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 8 * 4 * C);
  uint64_t* empty = full + 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(full + i, 1);
      umma::mbar_init(empty + i, kConsumers / 32);
    }
    umma::mbar_fence_init();
    tma::prefetch_map(&tmA);
    tma::prefetch_map(&tmB);
  }
  __syncthreads();
  if (warp == kConsumers / 32) {
    if (lane == 0) {
      int k = 0;
      for (int f = blockIdx.x; f < n_frames; f += gridDim.x, ++k) {
        const int b = k & 1;
        umma::mbar_wait(empty + b, ((k >> 1) & 1) ^ 1);
        tma::expect_tx(full + b, (uint32_t)(2 * FB));
        load_frame(bufA + (size_t)b * FB, &tmA, f, full + b);
        load_frame(bufB + (size_t)b * FB, &tmB, f, full + b);
      }
    }
    return;
  }
  const int ch = tid % CH, pl = tid / CH;
  float sa[8], ha[8], sb[8], hb[8];
  load8f(ss_a + ch * 8, sa); load8f(ss_a + C + ch * 8, ha);
  load8f(ss_b + ch * 8, sb); load8f(ss_b + C + ch * 8, hb);
  const float inv_pp = 1.0f / (float)PP;
  int k = 0;
  for (int f = blockIdx.x; f < n_frames; f += gridDim.x, ++k) {
    const int b = k & 1;
    umma::mbar_wait(full + b, (k >> 1) & 1);
    const uint4* fa = reinterpret_cast<const uint4*>(bufA + (size_t)b * FB);
    const uint4* fb = reinterpret_cast<const uint4*>(bufB + (size_t)b * FB);
    float acc[8] = {}, cnt[8] = {}, ma[8] = {}, mb[8] = {};
    for (int p = pl; p < PP; p += NPL) {
      const uint4 ua = fa[p * CH + ch], ub = fb[p * CH + ch];
      float a[8], bb[8];
      unpack8(ua, a);
      unpack8(ub, bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float pre = fmaf(a[i], sa[i], ha[i]);                 // == bn.cu bn_pre<true>: the backward recomputes this mask
        pre += fmaf(bb[i], sb[i], hb[i]);
        const float on = pre > 0.f ? 1.f : 0.f;
        acc[i] = fmaf(on, pre, acc[i]);
        cnt[i] += on;
        ma[i] = fmaf(on, a[i], ma[i]);
        mb[i] = fmaf(on, bb[i], mb[i]);
      }
    }
    __syncwarp();
    if (lane == 0) arrive(empty + b);     // this warp is done with the buffer
    // pixel lanes of a warp, then the 8 warps through shared memory
#pragma unroll
    for (int o = CH; o < 32; o <<= 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        cnt[i] += __shfl_xor_sync(0xffffffffu, cnt[i], o);
        ma[i] += __shfl_xor_sync(0xffffffffu, ma[i], o);
        mb[i] += __shfl_xor_sync(0xffffffffu, mb[i], o);
      }
    }
    if (lane < CH) {
      float* rw = red + (size_t)warp * 4 * C + ch * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) { rw[i] = acc[i]; rw[C + i] = cnt[i]; rw[2 * C + i] = ma[i]; rw[3 * C + i] = mb[i]; }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int i = tid; i < 4 * C; i += kConsumers) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[(size_t)w * 4 * C + i];
      const int q = i / C, c = i - q * C;
      if (q == 0) pooled[(size_t)f * C + c] = t * inv_pp;
      else if (fsums != nullptr) fsums[(size_t)f * 3 * C + (q - 1) * C + c] = t;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    (void)LPW;
  }
}

// Reduction pass of a BatchNorm backward (bn.cu: bn_bwd_reduce_kernel<C, UP, DUAL>):
//   sums[0][c] += sum m g,  sums[1][c] += sum m g xhat_a,  (DUAL) sums[2][c] += sum m g xhat_b,
//   g = up_a (+ up_b),  m = 1[bn_a(raw_a) (+ bn_b(raw_b)) > 0],  xhat = (raw - mean) * invstd.
// FPS frames per stage ({C, P, P, FPS} boxes; frames past the end read as zeros and contribute nothing).
template <int C, int UP, bool DUAL>
__global__ void __launch_bounds__(kConsumers + 32, 1)
bn_bwd_reduce_frames_kernel(const __grid_constant__ CUtensorMap tmRa, const __grid_constant__ CUtensorMap tmRb,
                            const __grid_constant__ CUtensorMap tmUa, const __grid_constant__ CUtensorMap tmUb,
                            const float* __restrict__ ss_a, const float* __restrict__ mi_a, const float* __restrict__ ss_b,
                            const float* __restrict__ mi_b, float* __restrict__ sums, int n_stages, int P, int fps) {
  constexpr int CH = C / 8, NPL = kConsumers / CH;
  constexpr int NS = (DUAL ? 2 : 1) + UP;
  extern __shared__ __align__(128) uint8_t smem[];
  const int SP = fps * P * P;                                   // pixels per stage
  const int TX = SP * C * 2;                                    // bytes of one tensor's stage
  const int TB = (TX + 127) & ~127;                             // its slot (TMA destinations are 128-byte aligned)
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)2 * NS * TB);
  uint64_t* empty = full + 2;
  float* red = reinterpret_cast<float*>(empty + 2);             // [3][8 warps][C]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(full + i, 1);
      umma::mbar_init(empty + i, kConsumers / 32);
    }
    umma::mbar_fence_init();
    tma::prefetch_map(&tmRa);
    tma::prefetch_map(&tmUa);
    if (DUAL) tma::prefetch_map(&tmRb);
    if (UP == 2) tma::prefetch_map(&tmUb);
  }
  __syncthreads();
  if (warp == kConsumers / 32) {
    if (lane == 0) {
      int k = 0;
      for (int s = blockIdx.x; s < n_stages; s += gridDim.x, ++k) {
        const int b = k & 1;
        umma::mbar_wait(empty + b, ((k >> 1) & 1) ^ 1);
        uint8_t* buf = smem + (size_t)b * NS * TB;
        tma::expect_tx(full + b, (uint32_t)(NS * TX));
        load_frame(buf, &tmRa, s * fps, full + b);
        load_frame(buf + TB, &tmUa, s * fps, full + b);
        if (DUAL) load_frame(buf + 2 * TB, &tmRb, s * fps, full + b);
        if (UP == 2) load_frame(buf + (DUAL ? 3 : 2) * TB, &tmUb, s * fps, full + b);
      }
    }
    return;
  }
  const int ch = tid % CH, pl = tid / CH;
  float sa[8], ha[8], sb[8], hb[8];
  load8f(ss_a + ch * 8, sa); load8f(ss_a + C + ch * 8, ha);
  if (DUAL) { load8f(ss_b + ch * 8, sb); load8f(ss_b + C + ch * 8, hb); }
  float s0[8] = {}, s1[8] = {}, s2[8] = {};   // sum g*x is accumulated and turned into sum g*xhat once at the end
  int k = 0;
  for (int s = blockIdx.x; s < n_stages; s += gridDim.x, ++k) {
    const int b = k & 1;
    umma::mbar_wait(full + b, (k >> 1) & 1);
    const uint8_t* buf = smem + (size_t)b * NS * TB;
    const uint4* ra = reinterpret_cast<const uint4*>(buf);
    const uint4* ua = reinterpret_cast<const uint4*>(buf + TB);
    const uint4* rb = reinterpret_cast<const uint4*>(buf + 2 * TB);
    const uint4* ub = reinterpret_cast<const uint4*>(buf + (DUAL ? 3 : 2) * TB);
    for (int p = pl; p < SP; p += NPL) {
      float a[8], bb[8], g[8];
      unpack8(ra[p * CH + ch], a);
      unpack8(ua[p * CH + ch], g);
      if (DUAL) unpack8(rb[p * CH + ch], bb);
      if (UP == 2) {
        float g2[8];
        unpack8(ub[p * CH + ch], g2);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] += g2[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float pre = fmaf(a[i], sa[i], ha[i]);                 // == bn.cu bn_pre
        if (DUAL) pre += fmaf(bb[i], sb[i], hb[i]);
        const float gi = pre > 0.f ? g[i] : 0.f;
        s0[i] += gi;
        s1[i] = fmaf(gi, a[i], s1[i]);
        if (DUAL) s2[i] = fmaf(gi, bb[i], s2[i]);
      }
    }
    __syncwarp();
    if (lane == 0) arrive(empty + b);
  }
  {
    float m[8], iv[8];
    load8f(mi_a + ch * 8, m);
    load8f(mi_a + C + ch * 8, iv);
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[i] = (s1[i] - m[i] * s0[i]) * iv[i];
    if (DUAL) {
      load8f(mi_b + ch * 8, m);
      load8f(mi_b + C + ch * 8, iv);
#pragma unroll
      for (int i = 0; i < 8; ++i) s2[i] = (s2[i] - m[i] * s0[i]) * iv[i];
    }
  }
  // pixel lanes of a warp (32 / CH of them), then the 8 warps through shared memory, one atomic per channel and CTA
#pragma unroll
  for (int o = CH; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s0[i] += __shfl_xor_sync(0xffffffffu, s0[i], o);
      s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
      if (DUAL) s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
    }
  }
  if (lane < CH) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[(0 * 8 + warp) * C + ch * 8 + i] = s0[i];
      red[(1 * 8 + warp) * C + ch * 8 + i] = s1[i];
      if (DUAL) red[(2 * 8 + warp) * C + ch * 8 + i] = s2[i];
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  for (int i = tid; i < (DUAL ? 3 : 2) * C; i += kConsumers) {
    const int q = i / C, c = i - q * C;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[(q * 8 + w) * C + c];
    atomicAdd(sums + q * C + c, t);
  }
}

// frames per stage and shared memory of the reduction kernel for a given shape (0: does not fit)
int reduce_frames_per_stage(int P, int C, int ns) {
  const int frame_bytes = P * P * C * 2;
  int fps = (86 * 1024) / (ns * frame_bytes);    // ~86 KB per stage, two stages
  if (fps > 256) fps = 256;
  return fps;
}

}  // namespace

bool bn_frames_supported(int P, int C) {
  if (C != 64 && C != 128) return false;
  const size_t smem = (size_t)4 * P * P * C * 2 + 8 * 4 * C * 4 + 64;   // the larger of the two kernels' footprints
  return smem <= 227 * 1024 && P * P >= kConsumers / (C / 8) && P >= 1;
}

int bn_backward_pooled_frames(const float* dpooled, const __nv_bfloat16* raw_a, const float* ss_a, const float* coef_a,
                              __nv_bfloat16* draw_a, const __nv_bfloat16* raw_b, const float* ss_b, const float* coef_b,
                              __nv_bfloat16* draw_b, long long rows, long long rows_pad, int P, int C, cudaStream_t st) {
  MIVIT_CHECK_ARG(bn_frames_supported(P, C), "frame-at-a-time BatchNorm backward: unsupported shape");
  const long long n_frames = rows / ((long long)(P + 1) * (P + 1));
  MIVIT_CHECK_ARG(n_frames < (1ll << 31), "too many frames for one launch (%lld)", n_frames);
  CUtensorMap tmA, tmB;
  int rc = make_frame_map(&tmA, raw_a, C, P, n_frames);
  if (!rc) rc = make_frame_map(&tmB, raw_b, C, P, n_frames);
  if (rc) return rc;
  const int smem = 4 * P * P * C * 2 + 2 * C * 4 + 64;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = (int)(n_frames < sms ? n_frames : sms);
  if (grid <= 0) return MIVIT_OK;
  if (C == 128) {
    auto kern = bn_bwd_apply_pooled_frames_kernel<128>;
    MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, kConsumers + 32, smem, st>>>(tmA, tmB, dpooled, ss_a, coef_a, draw_a, ss_b, coef_b, draw_b, (int)n_frames, P, rows, rows_pad);
  } else {
    auto kern = bn_bwd_apply_pooled_frames_kernel<64>;
    MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, kConsumers + 32, smem, st>>>(tmA, tmB, dpooled, ss_a, coef_a, draw_a, ss_b, coef_b, draw_b, (int)n_frames, P, rows, rows_pad);
  }
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

int bn_apply_pool_frames(const __nv_bfloat16* raw_a, const float* ss_a, const __nv_bfloat16* raw_b, const float* ss_b, float* pooled,
                         float* fsums, long long n_frames, int P, int C, cudaStream_t st) {
  MIVIT_CHECK_ARG(bn_frames_supported(P, C), "frame-at-a-time BatchNorm + pool: unsupported shape");
  MIVIT_CHECK_ARG(n_frames < (1ll << 31), "too many frames for one launch (%lld)", n_frames);
  CUtensorMap tmA, tmB;
  int rc = make_frame_map(&tmA, raw_a, C, P, n_frames);
  if (!rc) rc = make_frame_map(&tmB, raw_b, C, P, n_frames);
  if (rc) return rc;
  const int smem = 4 * P * P * C * 2 + 8 * 4 * C * 4 + 64;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = (int)(n_frames < sms ? n_frames : sms);
  if (grid <= 0) return MIVIT_OK;
  if (C == 128) {
    auto kern = bn_apply_pool_frames_kernel<128>;
    MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, kConsumers + 32, smem, st>>>(tmA, tmB, ss_a, ss_b, pooled, fsums, (int)n_frames, P);
  } else {
    auto kern = bn_apply_pool_frames_kernel<64>;
    MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, kConsumers + 32, smem, st>>>(tmA, tmB, ss_a, ss_b, pooled, fsums, (int)n_frames, P);
  }
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

bool bn_reduce_frames_supported(int P, int C, int ns) {
  return (C == 32 || C == 64 || C == 128) && P >= 1 && reduce_frames_per_stage(P, C, ns) >= 1;
}

// make_frame_map with `fps` frames per box
static int make_stage_map(CUtensorMap* tm, const void* row0, int C, int P, long long n_frames, int fps) {
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)(P + 1), (cuuint64_t)(P + 1), (cuuint64_t)n_frames};
  const cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)(P + 1) * C * 2, (cuuint64_t)(P + 1) * (P + 1) * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)P, (cuuint32_t)P, (cuuint32_t)fps};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  mivit_tensor_map_encode_fn encode = mivit_tensor_map_encoder();
  if (encode == nullptr) return MIVIT_ERR_CUDA;
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(row0), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mivit_set_error("cuTensorMapEncodeTiled failed (%d) for the stage view C=%d P=%d frames=%lld x %d", (int)r, C, P, n_frames, fps);
    return MIVIT_ERR_CUDA;
  }
  return MIVIT_OK;
}

template <int C, int UP, bool DUAL>
static int launch_reduce_frames(const __nv_bfloat16* up_a, const __nv_bfloat16* up_b, const __nv_bfloat16* raw_a, const float* ss_a,
                                const float* mi_a, const __nv_bfloat16* raw_b, const float* ss_b, const float* mi_b, float* sums,
                                long long n_frames, int P, cudaStream_t st) {
  constexpr int NS = (DUAL ? 2 : 1) + UP;
  const int fps = reduce_frames_per_stage(P, C, NS);
  CUtensorMap tmRa, tmRb, tmUa, tmUb;
  int rc = make_stage_map(&tmRa, raw_a, C, P, n_frames, fps);
  if (!rc) rc = make_stage_map(&tmUa, up_a, C, P, n_frames, fps);
  tmRb = tmRa;
  tmUb = tmUa;
  if (!rc && DUAL) rc = make_stage_map(&tmRb, raw_b, C, P, n_frames, fps);
  if (!rc && UP == 2) rc = make_stage_map(&tmUb, up_b, C, P, n_frames, fps);
  if (rc) return rc;
  const int n_stages = (int)((n_frames + fps - 1) / fps);
  const int smem = 2 * NS * ((fps * P * P * C * 2 + 127) & ~127) + 4 * 8 + 3 * 8 * C * 4 + 64;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = n_stages < sms ? n_stages : sms;
  if (grid <= 0) return MIVIT_OK;
  auto kern = bn_bwd_reduce_frames_kernel<C, UP, DUAL>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, kConsumers + 32, smem, st>>>(tmRa, tmRb, tmUa, tmUb, ss_a, mi_a, ss_b, mi_b, sums, n_stages, P, fps);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

int bn_backward_reduce_frames(const __nv_bfloat16* up_a, const __nv_bfloat16* up_b, const __nv_bfloat16* raw_a, const float* ss_a,
                              const float* mi_a, const __nv_bfloat16* raw_b, const float* ss_b, const float* mi_b, float* sums,
                              long long rows, int P, int C, cudaStream_t st) {
  const long long n_frames = rows / ((long long)(P + 1) * (P + 1));
  MIVIT_CHECK_ARG(n_frames < (1ll << 31), "too many frames for one launch (%lld)", n_frames);
  const bool dual = raw_b != nullptr, two = up_b != nullptr;
#define RF(CC)                                                                                                                    \
  if (C == CC) {                                                                                                                  \
    if (dual && two) return launch_reduce_frames<CC, 2, true>(up_a, up_b, raw_a, ss_a, mi_a, raw_b, ss_b, mi_b, sums, n_frames, P, st);   \
    if (dual) return launch_reduce_frames<CC, 1, true>(up_a, up_b, raw_a, ss_a, mi_a, raw_b, ss_b, mi_b, sums, n_frames, P, st);          \
    if (two) return launch_reduce_frames<CC, 2, false>(up_a, up_b, raw_a, ss_a, mi_a, raw_b, ss_b, mi_b, sums, n_frames, P, st);          \
    return launch_reduce_frames<CC, 1, false>(up_a, up_b, raw_a, ss_a, mi_a, raw_b, ss_b, mi_b, sums, n_frames, P, st);                   \
  }
  RF(32) RF(64) RF(128)
#undef RF
  mivit_set_error("BatchNorm width %d not supported", C);
  return MIVIT_ERR_INVALID;
}
