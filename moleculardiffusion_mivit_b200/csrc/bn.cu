// BatchNorm2d in TRAIN mode (batch statistics over all B*F*P*P positions), ReLU, residual add,
// global average pool and the first 1->32 convolution of DeepResNetEmbedding
// (reference helpers/models.py:202-257), on the pitched-rows bf16 layout of conv_tc.cu.
// All kernels here are HBM-bound element-wise / reduction passes: 16-byte vector accesses,
// per-thread register partials, one atomic per channel per CTA.
#include <stdlib.h>

#include "common.cuh"
#include "vit.h"

namespace {

// Row -> (frame-local index, y, x) needs r mod (P+1)^2 and a division by P+1 for EVERY 16-byte
// access; 64-bit '%' costs ~100 instructions and made these passes instruction-bound, so both are
// done with 32-bit multiply-high "magic number" division (Granlund-Montgomery round-up form).
struct FastDiv {
  uint32_t d, m, s;
};
static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.s = s;
  f.m = (uint32_t)((((1ull << s) - d) << 32) / d + 1);
  return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.m, n);
  return f.s == 0 ? n : (t + ((n - t) >> 1)) >> (f.s - 1);
}
struct RowGeom {
  FastDiv rpf, pitch;
  uint32_t rows;
  int P;
};
static inline RowGeom make_geom(long long rows, int P) {
  RowGeom g;
  g.rpf = make_fastdiv((uint32_t)((P + 1) * (P + 1)));
  g.pitch = make_fastdiv((uint32_t)(P + 1));
  g.rows = (uint32_t)rows;
  g.P = P;
  return g;
}
__device__ __forceinline__ bool row_is_valid(uint32_t r, const RowGeom& g) {
  if (r >= g.rows) return false;
  const uint32_t q = r - fast_div(r, g.rpf) * g.rpf.d;
  const uint32_t y = fast_div(q, g.pitch), x = q - y * g.pitch.d;
  return y < (uint32_t)g.P && x < (uint32_t)g.P;
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

// ---------------------------------------------------------------- statistics -> affine ---
// stats = [sum | sumsq] (train) ; writes mean, invstd, scale = gamma*invstd, shift = beta - mean*scale
// and the running-stat update of nn.BatchNorm2d (momentum 0.1, unbiased running variance).
__global__ void bn_finalize_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ num_batches, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, float* __restrict__ scale_out, float* __restrict__ shift_out,
                                   int C, double count, float eps, float momentum, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    const double m = (double)stats[c] / count;
    double v = (double)stats[C + c] / count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    if (running_mean) {
      const double unb = count > 1.0 ? v * count / (count - 1.0) : v;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
      if (c == 0 && num_batches) *num_batches += 1;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float invstd = rsqrtf(var + eps);
  const float sc = gamma[c] * invstd;
  mean_out[c] = mean;
  invstd_out[c] = invstd;
  scale_out[c] = sc;
  shift_out[c] = beta[c] - mean * sc;
}

// ---------------------------------------------------------------- apply (+relu, +residual) ---
// act = relu(raw_a*scale_a + shift_a [+ raw_b*scale_b + shift_b]); pad rows -> 0.
// Thread = (8-channel chunk, row lane): the per-channel constants are loaded ONCE into registers as
// float4 and the thread streams rows.  Every iteration issues the 16-byte loads of kU rows
// unconditionally (pad rows exist in memory) BEFORE any of them is used: the passes are pure HBM
// streams and need ~40 KB in flight per SM; with loads hidden behind the validity branch they ran at
// 40-50 % of HBM bandwidth.
__device__ __forceinline__ void load8f(const float* __restrict__ p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {   // read-once data: do not pollute L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
constexpr int kRowsPerCta = 256;
constexpr int kU = 4;   // rows in flight per thread

// pre-activation of the (possibly residual) BatchNorm output: MUST be the same expression in forward and
// backward -- the backward recomputes the ReLU mask from it instead of reading the activation tensor.
template <bool DUAL>
__device__ __forceinline__ void bn_pre(const float (&a)[8], const float (&sa)[8], const float (&ha)[8], const float (&b)[8],
                                       const float (&sb)[8], const float (&hb)[8], float (&o)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    o[i] = fmaf(a[i], sa[i], ha[i]);
    if (DUAL) o[i] += fmaf(b[i], sb[i], hb[i]);
  }
}

template <int C, bool DUAL>
__global__ void __launch_bounds__(256, 2) bn_apply_kernel(const __nv_bfloat16* __restrict__ raw_a, const float* __restrict__ ss_a,
                                                          const __nv_bfloat16* __restrict__ raw_b, const float* __restrict__ ss_b,
                                                          __nv_bfloat16* __restrict__ act, RowGeom geo, long long rows_pad) {
  constexpr int CH = C / 8, RL = 256 / CH;
  const int ch = threadIdx.x % CH, rl = threadIdx.x / CH;
  float sa[8], ha[8], sb[8], hb[8];
  load8f(ss_a + ch * 8, sa);
  load8f(ss_a + C + ch * 8, ha);
  if (DUAL) {
    load8f(ss_b + ch * 8, sb);
    load8f(ss_b + C + ch * 8, hb);
  }
  const long long r0 = (long long)blockIdx.x * kRowsPerCta;
  const long long r1 = min(rows_pad, r0 + kRowsPerCta);
  for (long long r = r0 + rl; r < r1; r += kU * RL) {
    uint4 va[kU], vb[kU];
    bool ok[kU];   // pad rows (13.8 % of the rows at P = 13) are not read: their loads are predicated off, not branched around
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long rr = r + u * RL;
      ok[u] = rr < r1 && row_is_valid((uint32_t)rr, geo);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long rr = r + u * RL;
      if (ok[u]) {
        va[u] = ldg_stream(reinterpret_cast<const uint4*>(raw_a) + rr * CH + ch);
        if (DUAL) vb[u] = ldg_stream(reinterpret_cast<const uint4*>(raw_b) + rr * CH + ch);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long rr = r + u * RL;
      if (rr >= r1) break;
      float o[8] = {};
      if (ok[u]) {
        float a[8], b[8];
        unpack8(va[u], a);
        if (DUAL) unpack8(vb[u], b);
        bn_pre<DUAL>(a, sa, ha, b, sb, hb, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
      }
      reinterpret_cast<uint4*>(act)[rr * CH + ch] = pack8(o);
    }
  }
}

// Last BatchNorm(+residual)+ReLU of the embedding fused with AdaptiveAvgPool2d(1) (helpers/models.py:226-227,240,254):
//   pooled[f,c] = mean over the P*P pixels of relu(bn2(raw_a) + bn_skip(raw_b)).
// The block output is consumed by nothing else (the backward recomputes the ReLU mask from raw), so the
// [rows,128] activation is never written or re-read: 3 x 1.5 GB of HBM traffic per step less.  One CTA per frame.
// One WARP per frame (lanes = 16-byte channel chunks x row lanes, no shared memory, no block barrier): the warps of an SM
// stream independent frames, so the loads of one overlap the shuffles / stores of another.  (One CTA per frame with a
// shared-memory reduction spent as long in its epilogue as in its loads once the masked sums were added: 0.49 -> 0.98 ms.)
template <int C>
__global__ void __launch_bounds__(256, 2) bn_apply_pool_kernel(const __nv_bfloat16* __restrict__ raw_a, const float* __restrict__ ss_a,
                                                               const __nv_bfloat16* __restrict__ raw_b, const float* __restrict__ ss_b,
                                                               float* __restrict__ pooled, float* __restrict__ fsums, long long n_frames,
                                                               int P) {
  constexpr int CH = C / 8, RL = 32 / CH;
  const int lane = threadIdx.x & 31, ch = lane % CH, rl = lane / CH;
  float sa[8], ha[8], sb[8], hb[8];
  load8f(ss_a + ch * 8, sa);
  load8f(ss_a + C + ch * 8, ha);
  load8f(ss_b + ch * 8, sb);
  load8f(ss_b + C + ch * 8, hb);
  const int pitch = P + 1, PP = P * P;
  const float inv_pp = 1.0f / (float)PP;
  constexpr int U = 4;
  for (long long f = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); f < n_frames; f += (long long)gridDim.x * 8) {
    const long long base = f * pitch * pitch;
    float acc[8] = {}, ma[8] = {}, mb[8] = {}, cnt[8] = {};
    for (int v0 = rl; v0 < PP; v0 += U * RL) {
      uint4 va[U], vb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int v = v0 + u * RL;
        if (v < PP) {
          const int y = v / P, x = v - y * P;
          const long long idx = (base + y * pitch + x) * CH + ch;
          va[u] = ldg_stream(reinterpret_cast<const uint4*>(raw_a) + idx);
          vb[u] = ldg_stream(reinterpret_cast<const uint4*>(raw_b) + idx);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (v0 + u * RL >= PP) break;
        float a[8], b[8], o[8];
        unpack8(va[u], a);
        unpack8(vb[u], b);
        bn_pre<true>(a, sa, ha, b, sb, hb, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // the mask as a float (one FSET), then four FMA / FADD: 5 instructions per element instead of 9 (select + add per sum);
          // fmaf(1, x, s) rounds exactly like s + x, and the counts stay exact (< 2^24 pixels per lane)
          const float on = o[i] > 0.f ? 1.f : 0.f;
          acc[i] = fmaf(on, o[i], acc[i]);
          cnt[i] += on;
          ma[i] = fmaf(on, a[i], ma[i]);      // per-frame masked sums for the backward of this BatchNorm pair (see below)
          mb[i] = fmaf(on, b[i], mb[i]);
        }
      }
    }
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
    if (rl == 0) {
      float4* pd = reinterpret_cast<float4*>(pooled + (size_t)f * C + ch * 8);
      pd[0] = make_float4(acc[0] * inv_pp, acc[1] * inv_pp, acc[2] * inv_pp, acc[3] * inv_pp);
      pd[1] = make_float4(acc[4] * inv_pp, acc[5] * inv_pp, acc[6] * inv_pp, acc[7] * inv_pp);
      if (fsums != nullptr) {
        float4* q = reinterpret_cast<float4*>(fsums + (size_t)f * 3 * C + ch * 8);
        q[0] = make_float4(cnt[0], cnt[1], cnt[2], cnt[3]);
        q[1] = make_float4(cnt[4], cnt[5], cnt[6], cnt[7]);
        q[C / 4] = make_float4(ma[0], ma[1], ma[2], ma[3]);
        q[C / 4 + 1] = make_float4(ma[4], ma[5], ma[6], ma[7]);
        q[C / 2] = make_float4(mb[0], mb[1], mb[2], mb[3]);
        q[C / 2 + 1] = make_float4(mb[4], mb[5], mb[6], mb[7]);
      }
    }
  }
}

// Backward reduction of the LAST BatchNorm pair without a pass over the activations.  Its upstream gradient is the pooled
// gradient broadcast over the frame, g[r,c] = dpooled[f,c]/P^2 * 1[pre > 0], so
//   sum_r g = sum_f dp[f,c]/P^2 * n+[f,c],   sum_r g*raw = sum_f dp[f,c]/P^2 * (sum_pix mask*raw)[f,c]
// and the three per-frame masked sums (n+, sum mask*raw_a, sum mask*raw_b) were written by bn_apply_pool_kernel in the
// forward ([frames][3][C] floats = 1.5 KB per frame instead of re-reading 2 x 50 KB of raw activations per frame).
template <int C>
__global__ void __launch_bounds__(256) bn_bwd_reduce_pooled_kernel(const float* __restrict__ dpooled, const float* __restrict__ fsums,
                                                                   const float* __restrict__ mi_a, const float* __restrict__ mi_b,
                                                                   float* __restrict__ sums, long long n_frames, float inv_pp) {
  constexpr int FL = 256 / C;
  const int c = threadIdx.x % C, fl = threadIdx.x / C;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (long long f = (long long)blockIdx.x * FL + fl; f < n_frames; f += (long long)gridDim.x * FL) {
    const float g = __ldg(dpooled + f * C + c) * inv_pp;
    const float* q = fsums + f * 3 * C + c;
    s0 = fmaf(g, __ldg(q), s0);
    s1 = fmaf(g, __ldg(q + C), s1);
    s2 = fmaf(g, __ldg(q + 2 * C), s2);
  }
  s1 = (s1 - mi_a[c] * s0) * mi_a[C + c];
  s2 = (s2 - mi_b[c] * s0) * mi_b[C + c];
  __shared__ float red[3][FL][C];
  red[0][fl][c] = s0; red[1][fl][c] = s1; red[2][fl][c] = s2;
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += 256) {
    const int k = i / C, cc = i - k * C;
    float t = 0.f;
    for (int j = 0; j < FL; ++j) t += red[k][j][cc];
    atomicAdd(sums + k * C + cc, t);
  }
}

// pooled[f,c] = mean over the P*P valid rows of frame f of act[r,c]   (AdaptiveAvgPool2d(1), :240,254)
template <int C>
__global__ void __launch_bounds__(C) pool_rows_kernel(const __nv_bfloat16* __restrict__ act, float* __restrict__ pooled, int P) {
  const long long f = blockIdx.x;
  const int c = threadIdx.x;
  const int pitch = P + 1;
  const __nv_bfloat16* base = act + (size_t)f * pitch * pitch * C;
  float s = 0.f;
  for (int y = 0; y < P; ++y)
    for (int x = 0; x < P; ++x) s += __bfloat162float(base[(size_t)(y * pitch + x) * C + c]);
  pooled[f * C + c] = s / (float)(P * P);
}

// ---------------------------------------------------------------- backward ----------------
// upstream gradient of the (post-ReLU) activation:
//   g = (up_a [+ up_b]  |  dpooled[f]/P^2) * 1[pre > 0],   pre = bn_pre(raw_a, raw_b) recomputed (the activation
// tensor is not read: one HBM stream less per pass).  UP: 0 = dpooled, 1 = up_a, 2 = up_a + up_b.
// sums[0][c] = sum g, sums[1][c] = sum g*xhat_a, sums[2][c] = sum g*xhat_b
// The pooled upstream (UP == 0) is constant over the (P+1)^2 rows of a frame: it is kept in registers and re-read only
// when the thread's row crosses into another frame (once per ~12 rows), not for every 16-byte chunk.
struct PooledCache {
  uint32_t f = 0xffffffffu;
  float v[8];
};
template <int C, int UP>
__device__ __forceinline__ void upstream8(const uint4& ua, const uint4& ub, const float* __restrict__ dpooled, uint32_t r, int ch,
                                          const RowGeom& geo, PooledCache& pc, float (&u)[8]) {
  if (UP == 0) {
    const uint32_t f = fast_div(r, geo.rpf);
    if (f != pc.f) {
      const float inv = 1.0f / (float)(geo.P * geo.P);
      load8f(dpooled + (size_t)f * C + ch * 8, pc.v);
#pragma unroll
      for (int i = 0; i < 8; ++i) pc.v[i] *= inv;
      pc.f = f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = pc.v[i];
  } else {
    unpack8(ua, u);
    if (UP == 2) {
      float w[8];
      unpack8(ub, w);
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] += w[i];
    }
  }
}

// Thread-private cp.async ring of the two backward passes (see bn_bwd_apply_kernel)
constexpr int kApplyStages = 3;
constexpr int kApplyRowsPerCta = 2048;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

template <int C, int UP, bool DUAL>
__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ up_a, const __nv_bfloat16* __restrict__ up_b, const float* __restrict__ dpooled,
                     const __nv_bfloat16* __restrict__ raw_a, const float* __restrict__ ss_a, const float* __restrict__ mi_a,
                     const __nv_bfloat16* __restrict__ raw_b, const float* __restrict__ ss_b, const float* __restrict__ mi_b,
                     float* __restrict__ sums, RowGeom geo, int rows_per_block) {
  constexpr int CH = C / 8;
  constexpr int RL = 256 / CH;  // row lanes
  constexpr int NS = (DUAL ? 2 : 1) + UP;
  constexpr int kU = NS >= 3 ? 2 : 4;
  constexpr int SLOTS = kU * NS;
  extern __shared__ uint4 ring[];   // [kApplyStages][SLOTS][256], thread-private slots
  const int ch = threadIdx.x % CH, rl = threadIdx.x / CH;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min((long long)geo.rows, r0 + rows_per_block);
  auto issue = [&](int i) {
    const long long r = r0 + rl + (long long)i * (kU * RL);
    uint4* stg = ring + (size_t)(i % kApplyStages) * SLOTS * 256 + threadIdx.x;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long rr = r + u * RL;
      if (rr < r1 && row_is_valid((uint32_t)rr, geo)) {
        const long long idx = rr * CH + ch;
        cp_async16(stg + (u * NS + 0) * 256, reinterpret_cast<const uint4*>(raw_a) + idx);
        if (DUAL) cp_async16(stg + (u * NS + 1) * 256, reinterpret_cast<const uint4*>(raw_b) + idx);
        if (UP >= 1) cp_async16(stg + (u * NS + (DUAL ? 2 : 1)) * 256, reinterpret_cast<const uint4*>(up_a) + idx);
        if (UP == 2) cp_async16(stg + (u * NS + (DUAL ? 3 : 2)) * 256, reinterpret_cast<const uint4*>(up_b) + idx);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int i = 0; i < kApplyStages - 1; ++i) issue(i);
  float sa[8], ha[8], sb[8], hb[8];
  load8f(ss_a + ch * 8, sa);
  load8f(ss_a + C + ch * 8, ha);
  if (DUAL) {
    load8f(ss_b + ch * 8, sb);
    load8f(ss_b + C + ch * 8, hb);
  }
  // sum g*xhat = invstd * (sum g*x - mean * sum g): accumulate sum g*x and fix up once at the end
  float s0[8] = {}, s1[8] = {}, s2[8] = {};
  PooledCache pcache;
  int it = 0;
  for (long long r = r0 + rl; r < r1; r += kU * RL, ++it) {
    issue(it + kApplyStages - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(kApplyStages - 1) : "memory");
    const uint4* stg = ring + (size_t)(it % kApplyStages) * SLOTS * 256 + threadIdx.x;
    uint4 va[kU], vb[kU], ua[kU], ub[kU];
    bool ok[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long rr = r + u * RL;
      ok[u] = rr < r1 && row_is_valid((uint32_t)rr, geo);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (ok[u]) {
        va[u] = stg[(u * NS + 0) * 256];
        if (DUAL) vb[u] = stg[(u * NS + 1) * 256];
        if (UP >= 1) ua[u] = stg[(u * NS + (DUAL ? 2 : 1)) * 256];
        if (UP == 2) ub[u] = stg[(u * NS + (DUAL ? 3 : 2)) * 256];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long rr = r + u * RL;
      if (rr >= r1) break;
      if (!ok[u]) continue;
      float a[8], b[8], pre[8], g[8];
      unpack8(va[u], a);
      if (DUAL) unpack8(vb[u], b);
      bn_pre<DUAL>(a, sa, ha, b, sb, hb, pre);
      upstream8<C, UP>(ua[u], ub[u], dpooled, (uint32_t)rr, ch, geo, pcache, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gi = pre[i] > 0.f ? g[i] : 0.f;
        s0[i] += gi;
        s1[i] = fmaf(gi, a[i], s1[i]);
        if (DUAL) s2[i] = fmaf(gi, b[i], s2[i]);
      }
    }
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
  __shared__ float red[DUAL ? 3 : 2][RL][C + 1];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[0][rl][ch * 8 + i] = s0[i];
    red[1][rl][ch * 8 + i] = s1[i];
    if (DUAL) red[2][rl][ch * 8 + i] = s2[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (DUAL ? 3 : 2) * C; i += 256) {
    const int k = i / C, c = i % C;
    float t = 0.f;
    for (int j = 0; j < RL; ++j) t += red[k][j][c];
    atomicAdd(sums + k * C + c, t);
  }
}

// draw = gamma*invstd * (g - sum_g/cnt - xhat * sum_gx/cnt)  =  k1*g + k2*raw + k3  with per-channel
//   k1 = gamma*invstd,  k2 = -k1*invstd*sum_gx/cnt,  k3 = -k1*sum_g/cnt - k2*mean
// (coef = [k1 | k2 | k3] per BatchNorm, written by bn_bwd_coef_kernel).
__global__ void bn_bwd_coef_kernel(const float* __restrict__ sums, const float* __restrict__ sums_local, const float* __restrict__ mi_a,
                                   const float* __restrict__ gamma_a,
                                   const float* __restrict__ mi_b, const float* __restrict__ gamma_b, float* __restrict__ coef_a,
                                   float* __restrict__ coef_b, float* __restrict__ dgamma_a, float* __restrict__ dbeta_a,
                                   float* __restrict__ dgamma_b, float* __restrict__ dbeta_b, int C, float inv_count, int raw_sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  {
    // raw_sums: slot 1 holds sum g*x instead of sum g*xhat = invstd * (sum g*x - mean * sum g)
    float s1 = sums[C + c], s1l = sums_local[C + c];
    if (raw_sums) {
      s1 = (s1 - mi_a[c] * sums[c]) * mi_a[C + c];
      s1l = (s1l - mi_a[c] * sums_local[c]) * mi_a[C + c];
    }
    const float k1 = gamma_a[c] * mi_a[C + c];
    const float k2 = -k1 * mi_a[C + c] * s1 * inv_count;
    coef_a[c] = k1; coef_a[C + c] = k2; coef_a[2 * C + c] = -k1 * sums[c] * inv_count - k2 * mi_a[c];
    dgamma_a[c] = s1l;
    dbeta_a[c] = sums_local[c];
  }
  if (mi_b != nullptr) {
    const float k1 = gamma_b[c] * mi_b[C + c];
    const float k2 = -k1 * mi_b[C + c] * sums[2 * C + c] * inv_count;
    coef_b[c] = k1; coef_b[C + c] = k2; coef_b[2 * C + c] = -k1 * sums[c] * inv_count - k2 * mi_b[c];
    dgamma_b[c] = sums_local[2 * C + c];
    dbeta_b[c] = sums_local[c];
  }
}

// Half (4 channels = two 32-bit words) of a 16-byte bf16x8 load, and its inverse.
__device__ __forceinline__ void unpack4(const uint4& v, int h, float (&f)[4]) {
  const uint32_t w0 = h ? v.z : v.x, w1 = h ? v.w : v.y;
  f[0] = __uint_as_float(w0 << 16); f[1] = __uint_as_float(w0 & 0xffff0000u);
  f[2] = __uint_as_float(w1 << 16); f[3] = __uint_as_float(w1 & 0xffff0000u);
}
__device__ __forceinline__ void pack4_into(uint4& v, int h, const float (&f)[4]) {
  const __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]), p1 = __floats2bfloat162_rn(f[2], f[3]);
  const uint32_t w0 = *reinterpret_cast<const uint32_t*>(&p0), w1 = *reinterpret_cast<const uint32_t*>(&p1);
  if (h) { v.z = w0; v.w = w1; } else { v.x = w0; v.y = w1; }
}
__device__ __forceinline__ void ldg4f(const float* __restrict__ p, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}

// The 10 (20 with the skip BatchNorm) per-channel constants of a thread's 8 channels do not fit in registers next to the
// rows in flight.  Re-reading them from L1 for EVERY row made this pass L1-bandwidth bound (12-16 LDG.128 of constants per
// 2-3 LDG.128 of data: 4.6 TB/s); they are now read once per 4-channel half and applied to all kU rows in flight, and the
// results overwrite the input registers half by half.
//
// Loads in flight: registers only hold 8 sixteen-byte loads per thread (64 KB per SM at 2 x 256 threads, and only while the
// thread is not computing), which left this pass at 5.0-5.6 TB/s.  The inputs now go through a THREAD-PRIVATE ring in shared
// memory filled with cp.async: slot (stage, row u, stream s) of thread t is ring[((stage*SLOTS + u*NS + s)*256 + t], so a
// thread only ever reads what it copied itself -- cp.async.wait_group is the only synchronisation, there is no barrier --
// consecutive threads touch consecutive 16 bytes (no bank conflicts), pad rows are simply not copied, and two iterations
// (2 x 8 slots x 16 B x 512 threads = 128 KB per SM) are always in flight behind the one being computed.
template <int C, int UP, bool DUAL>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ up_a, const __nv_bfloat16* __restrict__ up_b, const float* __restrict__ dpooled,
                    const __nv_bfloat16* __restrict__ raw_a, const float* __restrict__ ss_a, const float* __restrict__ coef_a,
                    __nv_bfloat16* __restrict__ draw_a, const __nv_bfloat16* __restrict__ raw_b, const float* __restrict__ ss_b,
                    const float* __restrict__ coef_b, __nv_bfloat16* __restrict__ draw_b, RowGeom geo, long long rows_pad) {
  constexpr int CH = C / 8, RL = 256 / CH;
  constexpr int NS = (DUAL ? 2 : 1) + UP;          // input streams
  constexpr int kU = NS >= 3 ? 2 : 4;
  constexpr int SLOTS = kU * NS;
  extern __shared__ uint4 ring[];                  // [kApplyStages][SLOTS][256]
  const int ch = threadIdx.x % CH, rl = threadIdx.x / CH;
  const long long r0 = (long long)blockIdx.x * kApplyRowsPerCta;
  const long long r1 = min(rows_pad, r0 + kApplyRowsPerCta);
  const float inv_pp = 1.0f / (float)(geo.P * geo.P);
  // copies of iteration i (rows r0 + rl + i*kU*RL + u*RL) into stage i % kApplyStages; always commits one group
  auto issue = [&](int i) {
    const long long r = r0 + rl + (long long)i * (kU * RL);
    uint4* stg = ring + (size_t)(i % kApplyStages) * SLOTS * 256 + threadIdx.x;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long rr = r + u * RL;
      if (rr < r1 && row_is_valid((uint32_t)rr, geo)) {
        const long long idx = rr * CH + ch;
        cp_async16(stg + (u * NS + 0) * 256, reinterpret_cast<const uint4*>(raw_a) + idx);
        if (DUAL) cp_async16(stg + (u * NS + 1) * 256, reinterpret_cast<const uint4*>(raw_b) + idx);
        if (UP >= 1) cp_async16(stg + (u * NS + (DUAL ? 2 : 1)) * 256, reinterpret_cast<const uint4*>(up_a) + idx);
        if (UP == 2) cp_async16(stg + (u * NS + (DUAL ? 3 : 2)) * 256, reinterpret_cast<const uint4*>(up_b) + idx);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int i = 0; i < kApplyStages - 1; ++i) issue(i);
  int it = 0;
  for (long long r = r0 + rl; r < r1; r += kU * RL, ++it) {
    issue(it + kApplyStages - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(kApplyStages - 1) : "memory");
    const uint4* stg = ring + (size_t)(it % kApplyStages) * SLOTS * 256 + threadIdx.x;
    uint4 va[kU], vb[kU], ua[kU], ub[kU];
    bool ok[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long rr = r + u * RL;
      ok[u] = rr < r1 && row_is_valid((uint32_t)rr, geo);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (ok[u]) {
        va[u] = stg[(u * NS + 0) * 256];
        if (DUAL) vb[u] = stg[(u * NS + 1) * 256];
        if (UP >= 1) ua[u] = stg[(u * NS + (DUAL ? 2 : 1)) * 256];
        if (UP == 2) ub[u] = stg[(u * NS + (DUAL ? 3 : 2)) * 256];
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c0 = ch * 8 + 4 * h;
      float gg[kU][4];   // masked upstream gradient of the kU rows, this half
      {
        float sa[4], ha[4], sb[4], hb[4];
        ldg4f(ss_a + c0, sa); ldg4f(ss_a + C + c0, ha);
        if (DUAL) { ldg4f(ss_b + c0, sb); ldg4f(ss_b + C + c0, hb); }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
#pragma unroll
          for (int j = 0; j < 4; ++j) gg[u][j] = 0.f;
          if (ok[u]) {
            const uint32_t rr = (uint32_t)(r + u * RL);
            float a[4], b[4], g[4];
            unpack4(va[u], h, a);
            if (DUAL) unpack4(vb[u], h, b);
            if (UP == 0) {   // pooled upstream: constant over the frame, L1/L2 resident ([frames][C] floats)
              // (keeping the frame's 8 values in registers while the kU rows share a frame -- 75 % of the iterations -- was
              //  measured: no change, 3.40-3.51 vs 3.40-3.42 ms for the five launches on the same box; the pass is not L1-bound)
              ldg4f(dpooled + (size_t)fast_div(rr, geo.rpf) * C + c0, g);
#pragma unroll
              for (int j = 0; j < 4; ++j) g[j] *= inv_pp;
            } else {
              unpack4(ua[u], h, g);
              if (UP == 2) {
                float g2[4];
                unpack4(ub[u], h, g2);
#pragma unroll
                for (int j = 0; j < 4; ++j) g[j] += g2[j];
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float pre = fmaf(a[j], sa[j], ha[j]);                 // == bn_pre
              if (DUAL) pre += fmaf(b[j], sb[j], hb[j]);
              gg[u][j] = pre > 0.f ? g[j] : 0.f;
            }
          }
        }
      }
      {
        float k1[4], k2[4], k3[4];
        ldg4f(coef_a + c0, k1); ldg4f(coef_a + C + c0, k2); ldg4f(coef_a + 2 * C + c0, k3);
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          float a[4], o[4] = {0.f, 0.f, 0.f, 0.f};
          if (ok[u]) {
            unpack4(va[u], h, a);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = fmaf(k1[j], gg[u][j], fmaf(k2[j], a[j], k3[j]));
          }
          pack4_into(va[u], h, o);
        }
      }
      if (DUAL) {
        float k1[4], k2[4], k3[4];
        ldg4f(coef_b + c0, k1); ldg4f(coef_b + C + c0, k2); ldg4f(coef_b + 2 * C + c0, k3);
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          float b[4], o[4] = {0.f, 0.f, 0.f, 0.f};
          if (ok[u]) {
            unpack4(vb[u], h, b);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = fmaf(k1[j], gg[u][j], fmaf(k2[j], b[j], k3[j]));
          }
          pack4_into(vb[u], h, o);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long rr = r + u * RL;
      if (rr >= r1) break;
      reinterpret_cast<uint4*>(draw_a)[rr * CH + ch] = va[u];
      if (DUAL) reinterpret_cast<uint4*>(draw_b)[rr * CH + ch] = vb[u];
    }
  }
}

// ---------------------------------------------------------------- first convolution -------
// raw0[r, 0:32] = sum_tap x[f, y+dy, x+dx] * W0[c, tap]   (Conv2d(1,32,3,padding=1,bias=False), :233)
//
// Thread = (image line, 8-channel chunk); it walks the line left to right with a sliding 3x3 window in registers, so a pixel
// costs 3 new input loads (not 9 bounds-checked ones), no row -> (frame, y, x) division, and its 72 weights stay in registers.
// (The first version mapped a thread to one output row: ~200 instructions per 16 stored bytes, 0.29 ms for 0.06 ms of traffic.)
struct Conv0Window {
  float c0[3], c1[3], c2[3];   // columns x-1, x, x+1 of rows y-1, y, y+1
  __device__ __forceinline__ void load_col(float (&c)[3], const float* __restrict__ img, int y, int x, int P) {
    const bool in = x >= 0 && x < P;
    c[0] = (in && y > 0) ? __ldg(img + (y - 1) * P + x) : 0.f;
    c[1] = in ? __ldg(img + y * P + x) : 0.f;
    c[2] = (in && y + 1 < P) ? __ldg(img + (y + 1) * P + x) : 0.f;
  }
  __device__ __forceinline__ void start(const float* __restrict__ img, int y, int P) {
    c1[0] = c1[1] = c1[2] = 0.f;
    load_col(c2, img, y, 0, P);
  }
  __device__ __forceinline__ void advance(const float* __restrict__ img, int y, int x, int P) {   // window centred on column x
#pragma unroll
    for (int k = 0; k < 3; ++k) { c0[k] = c1[k]; c1[k] = c2[k]; }
    load_col(c2, img, y, x + 1, P);
  }
  // tap t = (dy+1)*3 + (dx+1)
  __device__ __forceinline__ float tap(int t) const {
    const int dy = t / 3, dx = t - dy * 3;
    return dx == 0 ? c0[dy] : dx == 1 ? c1[dy] : c2[dy];
  }
};

__global__ void __launch_bounds__(256) conv0_fwd_kernel(const float* __restrict__ frames, const float* __restrict__ W0,
                                                        __nv_bfloat16* __restrict__ raw0, float* __restrict__ stats, long long rows,
                                                        long long rows_pad, int P) {
  __shared__ float red[2][64][33];
  const int pitch = P + 1;
  const int ch = threadIdx.x & 3;  // 4 chunks of 8 channels
  float w[8][9];                   // this thread's 8 output channels x 9 taps, resident in registers
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int t = 0; t < 9; ++t) w[i][t] = __ldg(W0 + (ch * 8 + i) * 9 + t);
  float s0[8] = {}, s1[8] = {};
  const long long n_lines = rows / pitch;   // rows = frames * pitch * pitch
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (long long line = (long long)blockIdx.x * 64 + (threadIdx.x >> 2); line < n_lines; line += (long long)gridDim.x * 64) {
    const long long f = line / pitch;
    const int y = (int)(line - f * pitch);
    uint4* out = reinterpret_cast<uint4*>(raw0) + (size_t)line * pitch * 4 + ch;
    if (y == P) {   // the pad line of the frame
      for (int x = 0; x < pitch; ++x) out[(size_t)x * 4] = zero;
      continue;
    }
    const float* img = frames + (size_t)f * P * P;
    Conv0Window win;
    win.start(img, y, P);
    for (int x = 0; x < P; ++x) {
      win.advance(img, y, x, P);
      float o[8] = {};
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float v = win.tap(t);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(v, w[i][t], o[i]);
      }
      const uint4 pk = pack8(o);
      out[(size_t)x * 4] = pk;
      float rb[8];
      unpack8(pk, rb);   // statistics of the STORED (bf16-rounded) values
#pragma unroll
      for (int i = 0; i < 8; ++i) { s0[i] += rb[i]; s1[i] = fmaf(rb[i], rb[i], s1[i]); }
    }
    out[(size_t)P * 4] = zero;   // the pad column
  }
  // rows between the last frame and the 128-row tile boundary
  if (blockIdx.x == 0)
    for (long long i = rows * 4 + threadIdx.x; i < rows_pad * 4; i += 256) reinterpret_cast<uint4*>(raw0)[i] = zero;
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[0][threadIdx.x >> 2][ch * 8 + i] = s0[i]; red[1][threadIdx.x >> 2][ch * 8 + i] = s1[i]; }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int k = threadIdx.x >> 5, c = threadIdx.x & 31;
    float t = 0.f;
    for (int j = 0; j < 64; ++j) t += red[k][j][c];
    atomicAdd(stats + k * 32 + c, t);
  }
}

// dW0[c, tap] = sum_r draw0[r, c] * x[f, y+dy, x+dx]      (same line-walking mapping; the 16-byte gradient load of the next
// pixel is issued before the 72 FMAs of the current one)
__global__ void __launch_bounds__(256) conv0_wgrad_kernel(const float* __restrict__ frames, const __nv_bfloat16* __restrict__ draw0,
                                                          float* __restrict__ dW0, long long rows, int P) {
  __shared__ float red[288];
  for (int i = threadIdx.x; i < 288; i += 256) red[i] = 0.f;
  __syncthreads();
  const int pitch = P + 1;
  const int ch = threadIdx.x & 3;
  float acc[8][9] = {};
  const long long n_lines = rows / pitch;
  for (long long line = (long long)blockIdx.x * 64 + (threadIdx.x >> 2); line < n_lines; line += (long long)gridDim.x * 64) {
    const long long f = line / pitch;
    const int y = (int)(line - f * pitch);
    if (y == P) continue;
    const uint4* gp = reinterpret_cast<const uint4*>(draw0) + (size_t)line * pitch * 4 + ch;
    const float* img = frames + (size_t)f * P * P;
    Conv0Window win;
    win.start(img, y, P);
    uint4 gnext = ldg_stream(gp);
    for (int x = 0; x < P; ++x) {
      const uint4 gv = gnext;
      if (x + 1 < P) gnext = ldg_stream(gp + (size_t)(x + 1) * 4);
      win.advance(img, y, x, P);
      float g[8];
      unpack8(gv, g);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float v = win.tap(t);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][t] = fmaf(g[i], v, acc[i][t]);
      }
    }
  }
  // lanes l, l+4, ..., l+28 of a warp hold the same channel chunk: butterfly over them first, so that 4 lanes per warp
  // (not 32) touch the shared accumulators -- 64-way contended shared atomics were ~75 % of this kernel's time
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float v = acc[i][t];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((threadIdx.x & 31) < 4) atomicAdd(&red[(ch * 8 + i) * 9 + t], v);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 288; i += 256) atomicAdd(dW0 + i, red[i]);
}

}  // namespace

int bn_finalize(const float* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                long long* num_batches, float* mean, float* invstd, float* scale, float* shift, int C, double count, float eps,
                float momentum, int training, cudaStream_t st) {
  bn_finalize_kernel<<<mivit_ceil_div(C, 128), 128, 0, st>>>(stats, gamma, beta, running_mean, running_var, num_batches, mean,
                                                             invstd, scale, shift, C, count, eps, momentum, training);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

#define BN_DISPATCH_C(C_, CALL)                                    \
  switch (C_) {                                                    \
    case 32: { constexpr int CC = 32; CALL; break; }               \
    case 64: { constexpr int CC = 64; CALL; break; }               \
    case 128: { constexpr int CC = 128; CALL; break; }             \
    default: mivit_set_error("BatchNorm width %d not supported", C_); return MIVIT_ERR_INVALID; \
  }

int bn_apply(const __nv_bfloat16* raw_a, const float* ss_a, const __nv_bfloat16* raw_b, const float* ss_b,
             __nv_bfloat16* act, long long rows, long long rows_pad, int P, int C, cudaStream_t st) {
  MIVIT_CHECK_ARG(rows_pad < (1ll << 31), "too many activation rows for one launch (%lld)", rows_pad);
  const int blocks = mivit_ceil_div(rows_pad, kRowsPerCta);
  const RowGeom geo = make_geom(rows, P);
  // work = ALGORITHMIC bytes: the valid pixels only (the loads of the pad rows, (P+1)^2 / P^2 - 1 = 16 % of the rows at P = 13,
  // are predicated off; their zero stores still reach DRAM -- the ncu figures are in profiles/r02_ncu_bn.md)
  const double valid_rows = (double)rows * P * P / ((double)(P + 1) * (P + 1));
  MivitProfScope prof("bn_apply", valid_rows * C * 2 * (raw_b ? 3 : 2), st);
  if (raw_b != nullptr) {
    BN_DISPATCH_C(C, (bn_apply_kernel<CC, true><<<blocks, 256, 0, st>>>(raw_a, ss_a, raw_b, ss_b, act, geo, rows_pad)));
  } else {
    BN_DISPATCH_C(C, (bn_apply_kernel<CC, false><<<blocks, 256, 0, st>>>(raw_a, ss_a, raw_b, ss_b, act, geo, rows_pad)));
  }
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

int bn_apply_pool(const __nv_bfloat16* raw_a, const float* ss_a, const __nv_bfloat16* raw_b, const float* ss_b, float* pooled,
                  float* fsums, long long n_frames, int P, int C, cudaStream_t st) {
  if (n_frames <= 0) return MIVIT_OK;
  MivitProfScope prof("bn_apply_pool", (double)n_frames * P * P * C * 2 * 2, st);
  if (bn_frames_supported(P, C) && getenv("MIVIT_NO_BN_FRAMES") == nullptr)   // one frame at a time, fed by TMA (bn_frames.cu)
    return bn_apply_pool_frames(raw_a, ss_a, raw_b, ss_b, pooled, fsums, n_frames, P, C, st);
  long long grid = (n_frames + 7) / 8;   // one warp per frame, 8 warps per CTA, grid-stride over the frames
  if (grid > 148 * 2 * 8) grid = 148 * 2 * 8;
  BN_DISPATCH_C(C, (bn_apply_pool_kernel<CC><<<(unsigned)grid, 256, 0, st>>>(raw_a, ss_a, raw_b, ss_b, pooled, fsums, n_frames, P)));
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

int pool_rows(const __nv_bfloat16* act, float* pooled, long long n_frames, int P, int C, cudaStream_t st) {
  if (n_frames <= 0) return MIVIT_OK;
  BN_DISPATCH_C(C, (pool_rows_kernel<CC><<<(unsigned)n_frames, CC, 0, st>>>(act, pooled, P)));
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}


template <int C, int UP, bool DUAL, typename... Args>
static int launch_bwd_reduce(int grid, cudaStream_t st, Args... args) {
  constexpr int NS = (DUAL ? 2 : 1) + UP, kU = NS >= 3 ? 2 : 4;
  constexpr int smem = kApplyStages * kU * NS * 256 * 16;
  auto kern = bn_bwd_reduce_kernel<C, UP, DUAL>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, 256, smem, st>>>(args...);
  return MIVIT_OK;
}
#define BN_BWD_REDUCE_LAUNCH(GRID, ...)                                                                                  \
  do {                                                                                                                   \
    const int up_mode = dpooled != nullptr ? 0 : (up_b != nullptr ? 2 : 1);                                             \
    const bool dual = raw_b != nullptr;                                                                                  \
    int rc_ = MIVIT_OK;                                                                                                  \
    if (up_mode == 0 && dual) { BN_DISPATCH_C(C, (rc_ = launch_bwd_reduce<CC, 0, true>(GRID, st, __VA_ARGS__))); }       \
    else if (up_mode == 0) { BN_DISPATCH_C(C, (rc_ = launch_bwd_reduce<CC, 0, false>(GRID, st, __VA_ARGS__))); }         \
    else if (up_mode == 1 && dual) { BN_DISPATCH_C(C, (rc_ = launch_bwd_reduce<CC, 1, true>(GRID, st, __VA_ARGS__))); }  \
    else if (up_mode == 1) { BN_DISPATCH_C(C, (rc_ = launch_bwd_reduce<CC, 1, false>(GRID, st, __VA_ARGS__))); }         \
    else if (dual) { BN_DISPATCH_C(C, (rc_ = launch_bwd_reduce<CC, 2, true>(GRID, st, __VA_ARGS__))); }                  \
    else { BN_DISPATCH_C(C, (rc_ = launch_bwd_reduce<CC, 2, false>(GRID, st, __VA_ARGS__))); }                           \
    if (rc_) return rc_;                                                                                                 \
  } while (0)

// bn_bwd_apply_kernel with its thread-private cp.async ring as dynamic shared memory (48-96 KB: opt-in attribute)
template <int C, int UP, bool DUAL, typename... Args>
static int launch_bwd_apply(int grid, cudaStream_t st, Args... args) {
  constexpr int NS = (DUAL ? 2 : 1) + UP, kU = NS >= 3 ? 2 : 4;
  constexpr int smem = kApplyStages * kU * NS * 256 * 16;
  auto kern = bn_bwd_apply_kernel<C, UP, DUAL>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, 256, smem, st>>>(args...);
  return MIVIT_OK;
}
#define BN_BWD_APPLY_LAUNCH(GRID, ...)                                                                                   \
  do {                                                                                                                   \
    const int up_mode = dpooled != nullptr ? 0 : (up_b != nullptr ? 2 : 1);                                             \
    const bool dual = raw_b != nullptr;                                                                                  \
    int rc_ = MIVIT_OK;                                                                                                  \
    if (up_mode == 0 && dual) { BN_DISPATCH_C(C, (rc_ = launch_bwd_apply<CC, 0, true>(GRID, st, __VA_ARGS__))); }        \
    else if (up_mode == 0) { BN_DISPATCH_C(C, (rc_ = launch_bwd_apply<CC, 0, false>(GRID, st, __VA_ARGS__))); }          \
    else if (up_mode == 1 && dual) { BN_DISPATCH_C(C, (rc_ = launch_bwd_apply<CC, 1, true>(GRID, st, __VA_ARGS__))); }   \
    else if (up_mode == 1) { BN_DISPATCH_C(C, (rc_ = launch_bwd_apply<CC, 1, false>(GRID, st, __VA_ARGS__))); }          \
    else if (dual) { BN_DISPATCH_C(C, (rc_ = launch_bwd_apply<CC, 2, true>(GRID, st, __VA_ARGS__))); }                   \
    else { BN_DISPATCH_C(C, (rc_ = launch_bwd_apply<CC, 2, false>(GRID, st, __VA_ARGS__))); }                            \
    if (rc_) return rc_;                                                                                                 \
  } while (0)

// ss_* = forward (scale | shift) of the BatchNorm(s): the ReLU mask is recomputed from raw, the activation is not read.
int bn_backward(const __nv_bfloat16* up_a, const __nv_bfloat16* up_b, const float* dpooled, const __nv_bfloat16* raw_a,
                const float* ss_a, const float* mi_a, const float* gamma_a, __nv_bfloat16* draw_a, float* dgamma_a, float* dbeta_a,
                const __nv_bfloat16* raw_b, const float* ss_b, const float* mi_b, const float* gamma_b, __nv_bfloat16* draw_b,
                float* dgamma_b, float* dbeta_b, float* sums /*[12][C] scratch: sums | coef_a | coef_b | local sums*/, long long rows,
                long long rows_pad, int P, int C, double count, const float* fsums, int presummed, cudaStream_t st) {
  MIVIT_CHECK_ARG(rows_pad < (1ll << 31), "too many activation rows for one launch (%lld)", rows_pad);
  MIVIT_CHECK_ARG(!presummed || (raw_b == nullptr && dpooled == nullptr), "pre-summed BatchNorm backward is single-input only");
  if (!presummed) MIVIT_CUDA_CHECK(cudaMemsetAsync(sums, 0, 3 * C * sizeof(float), st));
  const int rpb = kApplyRowsPerCta;
  const int blocks = mivit_ceil_div(rows, rpb);
  const RowGeom geo = make_geom(rows, P);
  if (presummed) {
    // (sum g, sum g*raw) were accumulated by the epilogue of the convolution that produced g (conv_tc4.cu, BSTAT)
  } else if (dpooled != nullptr && fsums != nullptr && raw_b != nullptr) {
    const long long n_frames = rows / ((long long)(P + 1) * (P + 1));
    MivitProfScope prof("bn_bwd_reduce_pooled", (double)n_frames * C * 16, st);
    const int fl = 256 / C;
    int grid = (int)((n_frames + fl - 1) / fl);
    if (grid > 148 * 4) grid = 148 * 4;
    BN_DISPATCH_C(C, (bn_bwd_reduce_pooled_kernel<CC><<<grid, 256, 0, st>>>(dpooled, fsums, mi_a, mi_b, sums, n_frames,
                                                                            1.0f / (float)(P * P))));
    mivit_count_launch();
    MIVIT_LAUNCH_CHECK();
  } else {
    MivitProfScope prof("bn_bwd_reduce", (double)rows * P * P / ((double)(P + 1) * (P + 1)) * C * 2 * ((raw_b ? 2 : 1) + (dpooled ? 0 : up_b ? 2 : 1)), st);
    const int n_streams = (raw_b ? 2 : 1) + (up_b ? 2 : 1);
    if (dpooled == nullptr && up_a != nullptr && bn_reduce_frames_supported(P, C, n_streams) && getenv("MIVIT_NO_BN_FRAMES") == nullptr) {
      // whole frames per stage, fed by TMA (bn_frames.cu)
      const int rc = bn_backward_reduce_frames(up_a, up_b, raw_a, ss_a, mi_a, raw_b, ss_b, mi_b, sums, rows, P, C, st);
      if (rc) return rc;
    } else {
      BN_BWD_REDUCE_LAUNCH(blocks, up_a, up_b, dpooled, raw_a, ss_a, mi_a, raw_b, ss_b, mi_b, sums, geo, rpb);
      mivit_count_launch();
      MIVIT_LAUNCH_CHECK();
    }
  }
  // sums -> coefficient form + parameter gradients (C threads), then the streaming pass
  float* coef_a = sums + 3 * C;
  float* coef_b = sums + 6 * C;
  const float* sums_local = sums;
  if (mivit_bn_sync_world() > 1) {   // synchronised BatchNorm: coefficients from the global sums, dgamma / dbeta from the local ones
    MIVIT_CUDA_CHECK(cudaMemcpyAsync(sums + 9 * C, sums, 3 * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    sums_local = sums + 9 * C;
    const int rc = mivit_bn_sync(sums, 3 * C, st);
    if (rc) return rc;
    count *= mivit_bn_sync_world();
  }
  bn_bwd_coef_kernel<<<mivit_ceil_div(C, 128), 128, 0, st>>>(sums, sums_local, mi_a, gamma_a, raw_b ? mi_b : nullptr, gamma_b, coef_a,
                                                             coef_b, dgamma_a, dbeta_a, dgamma_b, dbeta_b, C, (float)(1.0 / count),
                                                             presummed);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  const int ablocks = mivit_ceil_div(rows_pad, kApplyRowsPerCta);
  if (dpooled != nullptr && raw_b != nullptr && bn_frames_supported(P, C) && getenv("MIVIT_NO_BN_FRAMES") == nullptr) {
    // the pooled-gradient pair: one frame at a time, fed by TMA (bn_frames.cu)
    MivitProfScope prof("bn_bwd_apply", (double)rows * P * P / ((double)(P + 1) * (P + 1)) * C * 2 * 4, st);
    return bn_backward_pooled_frames(dpooled, raw_a, ss_a, coef_a, draw_a, raw_b, ss_b, coef_b, draw_b, rows, rows_pad, P, C, st);
  }
  {
    MivitProfScope prof("bn_bwd_apply", (double)rows * P * P / ((double)(P + 1) * (P + 1)) * C * 2 * (2 * (raw_b ? 2 : 1) + (dpooled ? 0 : up_b ? 2 : 1)), st);
    BN_BWD_APPLY_LAUNCH(ablocks, up_a, up_b, dpooled, raw_a, ss_a, coef_a, draw_a, raw_b, ss_b, coef_b, draw_b, geo, rows_pad);
    mivit_count_launch();
    MIVIT_LAUNCH_CHECK();
  }
  return MIVIT_OK;
}

int conv0_forward(const float* frames, const float* W0, __nv_bfloat16* raw0, float* stats, long long rows, long long rows_pad, int P,
                  cudaStream_t st) {
  MIVIT_CHECK_ARG(rows_pad < (1ll << 31), "too many activation rows for one launch (%lld)", rows_pad);
  MivitProfScope prof("conv0_fwd", (double)rows_pad * 64, st);
  int blocks = mivit_ceil_div(rows / (P + 1), 64);   // 64 image lines per CTA pass
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  conv0_fwd_kernel<<<blocks, 256, 0, st>>>(frames, W0, raw0, stats, rows, rows_pad, P);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

int conv0_wgrad(const float* frames, const __nv_bfloat16* draw0, float* dW0, long long rows, int P, cudaStream_t st) {
  MIVIT_CUDA_CHECK(cudaMemsetAsync(dW0, 0, 288 * sizeof(float), st));
  MivitProfScope prof("conv0_wgrad", (double)rows * 64, st);
  int blocks = mivit_ceil_div(rows / (P + 1), 64);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  conv0_wgrad_kernel<<<blocks, 256, 0, st>>>(frames, draw0, dW0, rows, P);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}
