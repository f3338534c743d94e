// CNN baselines of the reference experiments: every experiment trains a small ResNet beside each ViT
// (Experiments/PSFNoise/trainSettingsPSFNoise.py:114, trainSettingsImagesFeatures.py:170-173).  Reference classes restated here:
//   helpers/models.py:600-635  BasicBlock            conv3x3(stride) - BN - act - conv3x3 - BN - (+ shortcut: identity | conv1x1(stride) - BN) - act
//   helpers/models.py:638-683  LightResNet           conv5x5/2 - BN - act - maxpool3x3/2 - layer1(32) - layer2(64, /2) - layer3(128, /2) -
//                                                    global average pool - fc1(128 -> feature_size) - act - fc2(feature_size -> 1)
//   helpers/models.py:686-701  MultiImageResNet      per-frame LightResNet on [B*F, 1, P, P], predictions averaged over the frames
//   helpers/models.py:704-747  LightImagesFeaturesResNet  the same trunk up to act(fc1): per-frame features
//   helpers/models.py:749-772  MultiImageFeatureResNet    frame-averaged features ++ external features -> Linear - act - Linear
// One block per stage ([1,1,1], the only configuration the reference instantiates), train-mode BatchNorm with batch statistics.
//
// The frames are 9 x 9 ... 15 x 15 and shrink to 5x5 -> 3x3 -> 2x2 -> 1x1 (P = 9): ~1.3 MFLOP per frame against 46.5 MFLOP for
// the DeepResNet embedding, with feature maps far too small for 128-row tensor-core tiles.  Everything is fp32, channels-last
// ([pixels, C] matrices), and the convolutions (any kernel size / stride / padding) run as implicit GEMMs on a 64 x 64 register-
// tiled SIMT kernel whose A operand is gathered on the fly: forward (rows = output pixels), input gradient (rows = input
// pixels) and weight gradient (rows = (tap, c_in), reduction over the output pixels, split over CTAs with fp32 atomics).
#include <math.h>

#include "common.cuh"
#include "vit.h"
#include "../../include/mivit.h"

#define CK(expr)              \
  do {                        \
    int _rc = (expr);         \
    if (_rc) return _rc;      \
  } while (0)

namespace {

struct ConvGeom {
  int N, Hi, Wi, Ci, Ho, Wo, Co, k, s, p;
};
__host__ __device__ inline int conv_out(int h, int k, int s, int p) { return (h + 2 * p - k) / s + 1; }

constexpr int BM = 64, BN = 64, BK = 16;

// MODE 0 forward : C[m = (n,oy,ox)][co]       = sum_{tap,ci} x[n, oy*s-p+ky, ox*s-p+kx, ci] * wr[tap][ci][co]
// MODE 1 dgrad   : C[m = (n,iy,ix)][ci]       = sum_{tap,co} dy[n, (iy+p-ky)/s, (ix+p-kx)/s, co] * wr[tap][ci][co]
// MODE 2 wgrad   : C[m = (tap,ci)][co]       += sum_{pixel}  x[n, oy*s-p+ky, ox*s-p+kx, ci] * dy[pixel][co]     (split over blockIdx.z)
template <int MODE>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const float* __restrict__ src, const float* __restrict__ other,
                                                         float* __restrict__ C, ConvGeom g, float* __restrict__ stats, int k_chunk) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float st_s[2][BN];
  const int tid = threadIdx.x;
  const int taps = g.k * g.k;
  const int M = MODE == 0 ? g.N * g.Ho * g.Wo : MODE == 1 ? g.N * g.Hi * g.Wi : taps * g.Ci;
  const int Ncols = MODE == 0 ? g.Co : MODE == 1 ? g.Ci : g.Co;
  const int K = MODE == 0 ? taps * g.Ci : MODE == 1 ? taps * g.Co : g.N * g.Ho * g.Wo;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * k_chunk, ke = min(K, kb + k_chunk);
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  if (MODE == 0 && stats != nullptr && tid < 2 * BN) st_s[tid / BN][tid % BN] = 0.f;
  float acc[4][4] = {};
  for (int k0 = kb; k0 < ke; k0 += BK) {
    for (int i = tid; i < BM * BK; i += 256) {
      int mm, kk;
      if (MODE == 2) { mm = i % BM; kk = i / BM; } else { kk = i % BK; mm = i / BK; }
      const int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < ke) {
        if (MODE == 0) {
          const int ox = gm % g.Wo, t1 = gm / g.Wo, oy = t1 % g.Ho, n = t1 / g.Ho;
          const int ci = gk % g.Ci, tap = gk / g.Ci, ky = tap / g.k, kx = tap - ky * g.k;
          const int iy = oy * g.s - g.p + ky, ix = ox * g.s - g.p + kx;
          if (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi) v = src[((size_t)(n * g.Hi + iy) * g.Wi + ix) * g.Ci + ci];
        } else if (MODE == 1) {
          const int ix = gm % g.Wi, t1 = gm / g.Wi, iy = t1 % g.Hi, n = t1 / g.Hi;
          const int co = gk % g.Co, tap = gk / g.Co, ky = tap / g.k, kx = tap - ky * g.k;
          const int ty = iy + g.p - ky, tx = ix + g.p - kx;
          if (ty >= 0 && tx >= 0 && ty % g.s == 0 && tx % g.s == 0) {
            const int oy = ty / g.s, ox = tx / g.s;
            if (oy < g.Ho && ox < g.Wo) v = src[((size_t)(n * g.Ho + oy) * g.Wo + ox) * g.Co + co];
          }
        } else {
          const int ci = gm % g.Ci, tap = gm / g.Ci, ky = tap / g.k, kx = tap - ky * g.k;
          const int ox = gk % g.Wo, t1 = gk / g.Wo, oy = t1 % g.Ho, n = t1 / g.Ho;
          const int iy = oy * g.s - g.p + ky, ix = ox * g.s - g.p + kx;
          if (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi) v = src[((size_t)(n * g.Hi + iy) * g.Wi + ix) * g.Ci + ci];
        }
      }
      As[kk][mm] = v;
    }
    for (int i = tid; i < BN * BK; i += 256) {
      int nn, kk;
      if (MODE == 1) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
      const int gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < Ncols && gk < ke) {
        if (MODE == 0) {
          v = other[(size_t)gk * g.Co + gn];                               // wr[tap][ci][co], gk = tap*Ci + ci
        } else if (MODE == 1) {
          const int co = gk % g.Co, tap = gk / g.Co;
          v = other[((size_t)tap * g.Ci + gn) * g.Co + co];                // wr[tap][ci = gn][co]
        } else {
          v = other[(size_t)gk * g.Co + gn];                               // dy[pixel][co]
        }
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][tm]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tn]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + tm + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tn + j;
      if (gn >= Ncols) continue;
      const float v = acc[i][j];
      float* dst = C + (size_t)gm * Ncols + gn;
      if (MODE == 2) {
        atomicAdd(dst, v);
      } else {
        *dst = v;
        cs[j] += v;
        cq[j] = fmaf(v, v, cq[j]);
      }
    }
  }
  if (MODE == 0 && stats != nullptr) {      // BatchNorm batch statistics of the convolution output: per-CTA partial sums, then atomics
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&st_s[0][tn + j], cs[j]);
      atomicAdd(&st_s[1][tn + j], cq[j]);
    }
    __syncthreads();
    if (tid < 2 * BN) {
      const int which = tid / BN, c = n0 + tid % BN;
      if (c < Ncols) atomicAdd(stats + which * g.Co + c, st_s[which][tid % BN]);
    }
  }
}

// nn.Conv2d weight [co][ci][k][k]  <->  wr[tap][ci][co]
__global__ void reorder_w_kernel(const float* __restrict__ w, float* __restrict__ wr, int Co, int Ci, int taps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Co * Ci * taps) return;
  const int co = i % Co, t1 = i / Co, ci = t1 % Ci, tap = t1 / Ci;
  wr[i] = w[((size_t)co * Ci + ci) * taps + tap];
}
__global__ void unreorder_dw_kernel(const float* __restrict__ dwr, float* __restrict__ dw, int Co, int Ci, int taps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Co * Ci * taps) return;
  const int tap = i % taps, t1 = i / taps, ci = t1 % Ci, co = t1 / Ci;
  dw[i] = dwr[((size_t)tap * Ci + ci) * Co + co];
}

// y = act(x * scale + shift [+ res * rscale + rshift | + res])      ([rows, C] fp32; ss = [scale C | shift C])
__global__ void __launch_bounds__(256) rbn_apply_kernel(const float* __restrict__ x, const float* __restrict__ ss,
                                                        const float* __restrict__ res, const float* __restrict__ rss,
                                                        float* __restrict__ y, long long n, int C, int relu) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  float v = fmaf(x[i], ss[c], ss[C + c]);
  if (res != nullptr) v += rss != nullptr ? fmaf(res[i], rss[c], rss[C + c]) : res[i];
  y[i] = relu ? fmaxf(v, 0.f) : v;
}

// sums[c] += sum_rows gm, sums[C + c] += sum_rows gm * xhat, gm = g * [y > 0] (y = NULL: no activation), xhat = (x - mean) * invstd
__global__ void __launch_bounds__(256) rbn_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                             const float* __restrict__ x, const float* __restrict__ mi,
                                                             float* __restrict__ sums, long long rows, int C, int rows_per_cta) {
  // blockDim = 256: lanes walk the channels (C <= 128 -> C threads per row group), 256 / C row groups
  const int c = threadIdx.x % C, grp = threadIdx.x / C, ngrp = 256 / C;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  const float mean = mi[c], invstd = mi[C + c];
  float s0 = 0.f, s1 = 0.f;
  if (grp < ngrp)
    for (long long r = r0 + grp; r < r1; r += ngrp) {
      const size_t i = (size_t)r * C + c;
      float gm = g[i];
      if (y != nullptr && !(y[i] > 0.f)) gm = 0.f;
      s0 += gm;
      s1 = fmaf(gm, (x[i] - mean) * invstd, s1);
    }
  __shared__ float sh[2][256];
  sh[0][threadIdx.x] = s0;
  sh[1][threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.x < C) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < ngrp; ++k) { a += sh[0][k * C + threadIdx.x]; b += sh[1][k * C + threadIdx.x]; }
    atomicAdd(sums + threadIdx.x, a);
    atomicAdd(sums + C + threadIdx.x, b);
  }
}

// dx = gamma * invstd * (gm - mean(gm) - xhat * mean(gm * xhat));  optionally also writes gm (the masked upstream gradient,
// which the shortcut branch of a BasicBlock needs as well)
__global__ void __launch_bounds__(256) rbn_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                            const float* __restrict__ x, const float* __restrict__ mi,
                                                            const float* __restrict__ gamma, const float* __restrict__ sums,
                                                            float inv_count, float* __restrict__ dx, float* __restrict__ gm_out,
                                                            long long n, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  float gm = g[i];
  if (y != nullptr && !(y[i] > 0.f)) gm = 0.f;
  if (gm_out != nullptr) gm_out[i] = gm;
  const float xh = (x[i] - mi[c]) * mi[C + c];
  dx[i] = gamma[c] * mi[C + c] * (gm - sums[c] * inv_count - xh * sums[C + c] * inv_count);
}
__global__ void bn_param_grads_kernel(const float* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { dbeta[c] = sums[c]; dgamma[c] = sums[C + c]; }
}

// nn.MaxPool2d(kernel 3, stride 2, padding 1) on [N, Hi, Wi, C]; idx = winning input pixel (first maximum in scan order, like ATen)
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int* __restrict__ idx,
                                                          int N, int Hi, int Wi, int Ho, int Wo, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * Ho * Wo * C) return;
  const int c = (int)(i % C);
  long long t = i / C;
  const int ox = (int)(t % Wo); t /= Wo;
  const int oy = (int)(t % Ho);
  const int n = (int)(t / Ho);
  float best = -INFINITY;
  int bi = -1;
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    if (iy < 0 || iy >= Hi) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - 1 + kx;
      if (ix < 0 || ix >= Wi) continue;
      const float v = x[((size_t)(n * Hi + iy) * Wi + ix) * C + c];
      if (v > best || bi < 0) { best = v; bi = iy * Wi + ix; }
    }
  }
  y[i] = best;
  idx[i] = bi;
}
// dx (pre-zeroed) [N, Hi, Wi, C] += dy routed to the winning pixel
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ dy, const int* __restrict__ idx,
                                                          float* __restrict__ dx, int N, int Hi, int Wi, int Ho, int Wo, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * Ho * Wo * C) return;
  const int c = (int)(i % C);
  const int n = (int)(i / ((long long)Ho * Wo * C));
  atomicAdd(dx + ((size_t)n * Hi * Wi + idx[i]) * C + c, dy[i]);
}

// global average pool [N, HW, C] -> [N, C] and its backward; mean over the F frames of a sequence [B, F, C] -> [B, C] and backward
__global__ void __launch_bounds__(256) meanpool_kernel(const float* __restrict__ x, float* __restrict__ y, long long N, int HW, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const long long n = i / C;
  const int c = (int)(i % C);
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += x[((size_t)n * HW + p) * C + c];
  y[i] = s / (float)HW;
}
__global__ void __launch_bounds__(256) meanpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long N, int HW,
                                                           int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * HW * C) return;
  const int c = (int)(i % C);
  const long long n = i / ((long long)HW * C);
  dx[i] = dy[n * C + c] / (float)HW;
}
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}
__global__ void add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += b[i];
}
// [B, wa] ++ [B, wb] -> [B, wa + wb]  and the split of its gradient
__global__ void concat2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int B, int wa, int wb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * (wa + wb)) return;
  const int r = i / (wa + wb), c = i - r * (wa + wb);
  out[i] = c < wa ? a[r * wa + c] : b[r * wb + (c - wa)];
}
__global__ void take_cols_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int w_in, int w_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * w_out) return;
  const int r = i / w_out, c = i - r * w_out;
  out[i] = in[r * w_in + c];
}

inline int nblk(long long n, int t = 256) { return mivit_ceil_div(n, t); }
#define LAUNCHED()       \
  do {                   \
    mivit_count_launch(); \
    MIVIT_LAUNCH_CHECK(); \
  } while (0)

int conv_fwd(const float* x, const float* wr, float* y, float* stats, const ConvGeom& g, cudaStream_t st) {
  const int M = g.N * g.Ho * g.Wo, K = g.k * g.k * g.Ci;
  dim3 grid(mivit_ceil_div(g.Co, BN), mivit_ceil_div(M, BM), 1);
  MivitProfScope prof("resnet_conv_fwd", 2.0 * M * K * g.Co, st);
  conv_igemm_kernel<0><<<grid, 256, 0, st>>>(x, wr, y, g, stats, K);
  LAUNCHED();
  return MIVIT_OK;
}
int conv_dgrad(const float* dy, const float* wr, float* dx, const ConvGeom& g, cudaStream_t st) {
  const int M = g.N * g.Hi * g.Wi, K = g.k * g.k * g.Co;
  dim3 grid(mivit_ceil_div(g.Ci, BN), mivit_ceil_div(M, BM), 1);
  MivitProfScope prof("resnet_conv_dgrad", 2.0 * M * K * g.Ci, st);
  conv_igemm_kernel<1><<<grid, 256, 0, st>>>(dy, wr, dx, g, nullptr, K);
  LAUNCHED();
  return MIVIT_OK;
}
// dw [co][ci][k][k] = weight gradient; dwr: scratch [taps][ci][co]
int conv_wgrad(const float* x, const float* dy, float* dwr, float* dw, const ConvGeom& g, cudaStream_t st) {
  const int taps = g.k * g.k, M = taps * g.Ci, K = g.N * g.Ho * g.Wo;
  MIVIT_CUDA_CHECK(cudaMemsetAsync(dwr, 0, (size_t)M * g.Co * sizeof(float), st));
  const int tiles = mivit_ceil_div(g.Co, BN) * mivit_ceil_div(M, BM);
  int split = mivit_ceil_div(148 * 4, tiles);
  if (split > mivit_ceil_div(K, 4 * BK)) split = mivit_ceil_div(K, 4 * BK);
  if (split < 1) split = 1;
  int chunk = (mivit_ceil_div(K, split) + BK - 1) / BK * BK;
  dim3 grid(mivit_ceil_div(g.Co, BN), mivit_ceil_div(M, BM), mivit_ceil_div(K, chunk));
  {
    MivitProfScope prof("resnet_conv_wgrad", 2.0 * M * (double)K * g.Co, st);
    conv_igemm_kernel<2><<<grid, 256, 0, st>>>(x, dy, dwr, g, nullptr, chunk);
    LAUNCHED();
  }
  unreorder_dw_kernel<<<nblk((long long)M * g.Co), 256, 0, st>>>(dwr, dw, g.Co, g.Ci, taps);
  LAUNCHED();
  return MIVIT_OK;
}

// ----------------------------------------------------------------------------- parameter layout / workspace ----------------
struct BlockP { long c1_w, bn1_g, bn1_b, c2_w, bn2_g, bn2_b, sc_w, scbn_g, scbn_b; };
struct RLayout {
  long off = 0;
  int count = 0;
  long sizes[48];
  long add(long n) { const long o = off; sizes[count++] = n; off += n; return o; }
  long conv1_w, bn1_g, bn1_b;
  BlockP blk[3];
  long fc1_w, fc1_b, fc2_w = -1, fc2_b = -1, m0_w = -1, m0_b = -1, m2_w = -1, m2_b = -1;
};
const int kCh[4] = {32, 32, 64, 128};   // trunk widths: stem / layer1 / layer2 / layer3
const int kStride[3] = {1, 2, 2};

int r_layout(const mivit_resnet_config* c, RLayout& L) {
  MIVIT_CHECK_ARG(c != nullptr, "config is NULL");
  MIVIT_CHECK_ARG(c->P >= 3 && c->P <= 64 && c->F >= 1, "bad image size / frame count");
  MIVIT_CHECK_ARG(c->feature_size >= 1 && c->feature_size <= 1024, "bad feature_size");
  MIVIT_CHECK_ARG(c->ext_dim >= 0 && c->ext_dim <= 1024 && (c->ext_dim == 0 || c->hidden >= 1), "bad external feature / hidden size");
  MIVIT_CHECK_ARG(c->activation == 0, "only activation = nn.ReLU runs on the CUDA path");
  L.conv1_w = L.add(32 * 25); L.bn1_g = L.add(32); L.bn1_b = L.add(32);
  for (int b = 0; b < 3; ++b) {
    const int ci = kCh[b], co = kCh[b + 1];
    BlockP& B = L.blk[b];
    B.c1_w = L.add((long)co * ci * 9); B.bn1_g = L.add(co); B.bn1_b = L.add(co);
    B.c2_w = L.add((long)co * co * 9); B.bn2_g = L.add(co); B.bn2_b = L.add(co);
    if (kStride[b] != 1 || ci != co) { B.sc_w = L.add((long)co * ci); B.scbn_g = L.add(co); B.scbn_b = L.add(co); }
    else { B.sc_w = B.scbn_g = B.scbn_b = -1; }
  }
  L.fc1_w = L.add(128L * c->feature_size); L.fc1_b = L.add(c->feature_size);
  if (c->ext_dim == 0) {
    L.fc2_w = L.add(c->feature_size); L.fc2_b = L.add(1);
  } else {
    const long in = c->feature_size + c->ext_dim;
    L.m0_w = L.add(in * c->hidden); L.m0_b = L.add(c->hidden); L.m2_w = L.add(c->hidden); L.m2_b = L.add(1);
  }
  return MIVIT_OK;
}

struct Bump {
  uint8_t* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};
struct BnS { float *stats, *mi, *ss; int C; };   // (sum,sumsq) | (mean,invstd) | (scale,shift)
struct BlockWS {
  ConvGeom g1, g2, gs;
  float *c1raw, *c1act, *c2raw, *scraw, *out;            // forward
  float *d_out_m, *d_c2raw, *d_c1act, *d_c1raw, *d_scraw, *d_in, *d_in_sc;  // backward
  float *wr1, *wr2, *wrs;
  BnS bn1, bn2, bns;
  bool has_sc;
};
struct RWS {
  int NF, H1, H2;
  ConvGeom g0;
  float *a1raw, *a1, *pool, *wr0, *d_pool, *d_a1, *d_a1raw;
  int* pool_idx;
  BnS bn0;
  BlockWS blk[3];
  float *pooled, *f1, *ppred, *fmean, *cat, *hid, *d_hid, *d_cat, *d_fmean, *d_f1, *d_pooled, *d_ppred, *d_last;
  float *bn_sums, *dwr;
  size_t bytes;
};
void take_bn(Bump& b, BnS& s, int C) { s.C = C; s.stats = b.take<float>(2 * C); s.mi = b.take<float>(2 * C); s.ss = b.take<float>(2 * C); }

void r_carve(const mivit_resnet_config* c, int B, void* base, RWS& w) {
  Bump b{reinterpret_cast<uint8_t*>(base)};
  const int NF = B * c->F, P = c->P;
  w.NF = NF;
  w.H1 = conv_out(P, 5, 2, 2);
  w.H2 = conv_out(w.H1, 3, 2, 1);
  w.g0 = ConvGeom{NF, P, P, 1, w.H1, w.H1, 32, 5, 2, 2};
  w.a1raw = b.take<float>((size_t)NF * w.H1 * w.H1 * 32); w.a1 = b.take<float>((size_t)NF * w.H1 * w.H1 * 32);
  w.pool = b.take<float>((size_t)NF * w.H2 * w.H2 * 32); w.pool_idx = b.take<int>((size_t)NF * w.H2 * w.H2 * 32);
  w.wr0 = b.take<float>(32 * 25);
  w.d_pool = b.take<float>((size_t)NF * w.H2 * w.H2 * 32); w.d_a1 = b.take<float>((size_t)NF * w.H1 * w.H1 * 32);
  w.d_a1raw = b.take<float>((size_t)NF * w.H1 * w.H1 * 32);
  take_bn(b, w.bn0, 32);
  int H = w.H2;
  size_t max_w = 32 * 25;
  for (int k = 0; k < 3; ++k) {
    BlockWS& q = w.blk[k];
    const int ci = kCh[k], co = kCh[k + 1], s = kStride[k], Ho = conv_out(H, 3, s, 1);
    q.has_sc = s != 1 || ci != co;
    q.g1 = ConvGeom{NF, H, H, ci, Ho, Ho, co, 3, s, 1};
    q.g2 = ConvGeom{NF, Ho, Ho, co, Ho, Ho, co, 3, 1, 1};
    q.gs = ConvGeom{NF, H, H, ci, Ho, Ho, co, 1, s, 0};
    const size_t no = (size_t)NF * Ho * Ho * co, ni = (size_t)NF * H * H * ci;
    q.c1raw = b.take<float>(no); q.c1act = b.take<float>(no); q.c2raw = b.take<float>(no);
    q.scraw = q.has_sc ? b.take<float>(no) : nullptr; q.out = b.take<float>(no);
    q.d_out_m = b.take<float>(no); q.d_c2raw = b.take<float>(no); q.d_c1act = b.take<float>(no); q.d_c1raw = b.take<float>(no);
    q.d_scraw = q.has_sc ? b.take<float>(no) : nullptr; q.d_in = b.take<float>(ni);
    q.d_in_sc = q.has_sc ? b.take<float>(ni) : nullptr;
    q.wr1 = b.take<float>((size_t)9 * ci * co); q.wr2 = b.take<float>((size_t)9 * co * co);
    q.wrs = q.has_sc ? b.take<float>((size_t)ci * co) : nullptr;
    take_bn(b, q.bn1, co); take_bn(b, q.bn2, co);
    if (q.has_sc) take_bn(b, q.bns, co);
    if ((size_t)9 * co * co > max_w) max_w = (size_t)9 * co * co;
    H = Ho;
  }
  const int fs = c->feature_size;
  w.pooled = b.take<float>((size_t)NF * 128); w.f1 = b.take<float>((size_t)NF * fs); w.ppred = b.take<float>(NF);
  w.fmean = b.take<float>((size_t)B * fs); w.cat = b.take<float>((size_t)B * (fs + c->ext_dim));
  w.hid = b.take<float>((size_t)B * (c->hidden > 0 ? c->hidden : 1)); w.d_hid = b.take<float>((size_t)B * (c->hidden > 0 ? c->hidden : 1));
  w.d_cat = b.take<float>((size_t)B * (fs + c->ext_dim)); w.d_fmean = b.take<float>((size_t)B * fs);
  w.d_f1 = b.take<float>((size_t)NF * fs); w.d_pooled = b.take<float>((size_t)NF * 128); w.d_ppred = b.take<float>(NF);
  w.d_last = b.take<float>((size_t)NF * w.blk[2].g2.Ho * w.blk[2].g2.Wo * 128);
  w.bn_sums = b.take<float>(2 * 128);
  w.dwr = b.take<float>(max_w);
  w.bytes = (b.off + 255) & ~(size_t)255;
}

// running statistics: flat [mean C | var C] for bn1, layer1.{bn1,bn2}, layer2.{bn1,bn2,shortcut.1}, layer3.{bn1,bn2,shortcut.1}
struct RunBn { float* rm; float* rv; long long* nbt; };
struct RunCursor {
  float* base; long long* nbt; long off = 0; int i = 0;
  RunBn next(int C) {
    RunBn r{base ? base + off : nullptr, base ? base + off + C : nullptr, nbt ? nbt + i : nullptr};
    off += 2 * C; ++i;
    return r;
  }
};

// BatchNorm of a raw convolution output whose (sum, sumsq) are already in s.stats
int bn_fin(const mivit_resnet_config* c, BnS& s, const float* gamma, const float* beta, RunBn rb, double count, int training,
           cudaStream_t st) {
  return bn_finalize(s.stats, gamma, beta, rb.rm, rb.rv, rb.nbt, s.mi, s.mi + s.C, s.ss, s.ss + s.C, s.C, count, c->bn_eps,
                     c->bn_momentum, training, st);
}
int bn_apply_r(const float* x, const BnS& s, const float* res, const BnS* rs, float* y, long long n, int relu, cudaStream_t st) {
  rbn_apply_kernel<<<nblk(n), 256, 0, st>>>(x, s.ss, res, rs ? rs->ss : nullptr, y, n, s.C, relu);
  LAUNCHED();
  return MIVIT_OK;
}
// backward of y = act(bn(x) [+ ...]) w.r.t. x and (gamma, beta); g = dL/dy_out, yact = post-activation output (mask) or NULL
int bn_bwd_r(const float* g, const float* yact, const float* x, const BnS& s, const float* gamma, float* dx, float* gm_out,
             float* dgamma, float* dbeta, float* sums, long long rows, cudaStream_t st) {
  const int C = s.C;
  MIVIT_CHECK_ARG(C <= 256 && 256 % C == 0, "BatchNorm width %d not supported", C);
  MIVIT_CUDA_CHECK(cudaMemsetAsync(sums, 0, 2 * C * sizeof(float), st));
  const int rpc = 512;
  rbn_bwd_reduce_kernel<<<mivit_ceil_div(rows, rpc), 256, 0, st>>>(g, yact, x, s.mi, sums, rows, C, rpc);
  LAUNCHED();
  rbn_bwd_apply_kernel<<<nblk(rows * C), 256, 0, st>>>(g, yact, x, s.mi, gamma, sums, 1.0f / (float)rows, dx, gm_out, rows * C, C);
  LAUNCHED();
  bn_param_grads_kernel<<<1, 256, 0, st>>>(sums, dgamma, dbeta, C);
  LAUNCHED();
  return MIVIT_OK;
}
int lin_fwd(const float* X, const float* W, const float* b, float* Y, int M, int N, int K, int relu, cudaStream_t st) {
  return gemm_f32(X, K, 1, W, 1, K, Y, N, M, N, K, b, relu, 0, 1, st);
}
// dW [N][K], db [N] are OVERWRITTEN (pre-zeroed gradient buffer), dX optional
int lin_bwd(const float* X, const float* W, const float* dY, float* dW, float* db, float* dX, int M, int N, int K, cudaStream_t st) {
  int split = M / 256;
  if (split < 1) split = 1;
  if (split > 296) split = 296;
  CK(gemm_f32(dY, 1, N, X, K, 1, dW, K, N, K, M, nullptr, 0, split == 1 ? 1 : 0, split, st));
  CK(colsum_f32(dY, N, M, N, db, st));
  if (dX) CK(gemm_f32(dY, N, 1, W, K, 1, dX, K, M, K, N, nullptr, 0, 0, 1, st));
  return MIVIT_OK;
}

}  // namespace

extern "C" int32_t mivit_resnet_param_count(const mivit_resnet_config* cfg) {
  RLayout L;
  if (r_layout(cfg, L)) return -1;
  return L.count;
}
extern "C" int mivit_resnet_param_sizes(const mivit_resnet_config* cfg, int64_t* sizes, int32_t max_count) {
  RLayout L;
  CK(r_layout(cfg, L));
  MIVIT_CHECK_ARG(sizes != nullptr && max_count >= L.count, "sizes array too small");
  for (int i = 0; i < L.count; ++i) sizes[i] = L.sizes[i];
  return MIVIT_OK;
}
extern "C" int64_t mivit_resnet_workspace_bytes(const mivit_resnet_config* cfg, int32_t B) {
  RLayout L;
  if (r_layout(cfg, L) || B < 1) return -1;
  RWS w;
  r_carve(cfg, B, nullptr, w);
  return (int64_t)w.bytes;
}
extern "C" int32_t mivit_resnet_pred_rows(const mivit_resnet_config* cfg, int32_t B) {
  return (cfg->ext_dim == 0 && !cfg->single_prediction) ? B * cfg->F : B;
}

extern "C" int mivit_resnet_forward(const mivit_resnet_config* c, int32_t B, const float* x, const float* ext, const float* params,
                                    float* bn_running, int64_t* bn_num_batches, void* workspace, float* pred, int32_t training,
                                    void* stream) {
  RLayout L;
  CK(r_layout(c, L));
  MIVIT_CHECK_ARG(B >= 1 && x && params && workspace && pred, "bad arguments");
  MIVIT_CHECK_ARG(c->ext_dim == 0 || ext, "external features required");
  MIVIT_CHECK_ARG(training || bn_running, "eval-mode BatchNorm needs the running statistics");
  cudaStream_t st = (cudaStream_t)stream;
  RWS w;
  r_carve(c, B, workspace, w);
  const float* p = params;
  const int NF = w.NF, fs = c->feature_size;
  RunCursor rc{bn_running, (long long*)bn_num_batches};
  auto zero_stats = [&](BnS& s) { return cudaMemsetAsync(s.stats, 0, 2 * s.C * sizeof(float), st); };
  // stem: conv 5x5 / 2 -> BN -> ReLU -> maxpool 3x3 / 2     (helpers/models.py:643-646, :668-671)
  reorder_w_kernel<<<nblk(32 * 25), 256, 0, st>>>(p + L.conv1_w, w.wr0, 32, 1, 25);
  LAUNCHED();
  MIVIT_CUDA_CHECK(zero_stats(w.bn0));
  CK(conv_fwd(x, w.wr0, w.a1raw, w.bn0.stats, w.g0, st));
  const long long n1 = (long long)NF * w.H1 * w.H1;
  CK(bn_fin(c, w.bn0, p + L.bn1_g, p + L.bn1_b, rc.next(32), (double)n1, training, st));
  CK(bn_apply_r(w.a1raw, w.bn0, nullptr, nullptr, w.a1, n1 * 32, 1, st));
  maxpool_fwd_kernel<<<nblk((long long)NF * w.H2 * w.H2 * 32), 256, 0, st>>>(w.a1, w.pool, w.pool_idx, NF, w.H1, w.H1, w.H2, w.H2, 32);
  LAUNCHED();
  const float* in = w.pool;
  for (int k = 0; k < 3; ++k) {   // BasicBlock (:623-635)
    BlockWS& q = w.blk[k];
    const BlockP& Bp = L.blk[k];
    const int ci = kCh[k], co = kCh[k + 1];
    const long long rows = (long long)NF * q.g1.Ho * q.g1.Wo;
    reorder_w_kernel<<<nblk((long long)9 * ci * co), 256, 0, st>>>(p + Bp.c1_w, q.wr1, co, ci, 9);
    LAUNCHED();
    reorder_w_kernel<<<nblk((long long)9 * co * co), 256, 0, st>>>(p + Bp.c2_w, q.wr2, co, co, 9);
    LAUNCHED();
    MIVIT_CUDA_CHECK(zero_stats(q.bn1));
    MIVIT_CUDA_CHECK(zero_stats(q.bn2));
    CK(conv_fwd(in, q.wr1, q.c1raw, q.bn1.stats, q.g1, st));
    CK(bn_fin(c, q.bn1, p + Bp.bn1_g, p + Bp.bn1_b, rc.next(co), (double)rows, training, st));
    CK(bn_apply_r(q.c1raw, q.bn1, nullptr, nullptr, q.c1act, rows * co, 1, st));
    CK(conv_fwd(q.c1act, q.wr2, q.c2raw, q.bn2.stats, q.g2, st));
    CK(bn_fin(c, q.bn2, p + Bp.bn2_g, p + Bp.bn2_b, rc.next(co), (double)rows, training, st));
    if (q.has_sc) {
      reorder_w_kernel<<<nblk((long long)ci * co), 256, 0, st>>>(p + Bp.sc_w, q.wrs, co, ci, 1);
      LAUNCHED();
      MIVIT_CUDA_CHECK(zero_stats(q.bns));
      CK(conv_fwd(in, q.wrs, q.scraw, q.bns.stats, q.gs, st));
      CK(bn_fin(c, q.bns, p + Bp.scbn_g, p + Bp.scbn_b, rc.next(co), (double)rows, training, st));
      CK(bn_apply_r(q.c2raw, q.bn2, q.scraw, &q.bns, q.out, rows * co, 1, st));
    } else {
      CK(bn_apply_r(q.c2raw, q.bn2, in, nullptr, q.out, rows * co, 1, st));
    }
    in = q.out;
  }
  const int HW = w.blk[2].g2.Ho * w.blk[2].g2.Wo;
  meanpool_kernel<<<nblk((long long)NF * 128), 256, 0, st>>>(in, w.pooled, NF, HW, 128);   // AdaptiveAvgPool2d((1,1)) + flatten
  LAUNCHED();
  CK(lin_fwd(w.pooled, p + L.fc1_w, p + L.fc1_b, w.f1, NF, fs, 128, 1, st));               // fc1 + act
  if (c->ext_dim == 0) {
    float* per_frame = c->single_prediction ? w.ppred : pred;
    CK(lin_fwd(w.f1, p + L.fc2_w, p + L.fc2_b, per_frame, NF, 1, fs, 0, st));               // fc2 -> [B*F, 1]
    if (c->single_prediction) {                                                             // torch.mean over the frames (:697-699)
      meanpool_kernel<<<nblk(B), 256, 0, st>>>(w.ppred, pred, B, c->F, 1);
      LAUNCHED();
    }
  } else {                                                                                  // MultiImageFeatureResNet (:763-772)
    meanpool_kernel<<<nblk((long long)B * fs), 256, 0, st>>>(w.f1, w.fmean, B, c->F, fs);
    LAUNCHED();
    concat2_kernel<<<nblk((long long)B * (fs + c->ext_dim)), 256, 0, st>>>(w.fmean, ext, w.cat, B, fs, c->ext_dim);
    LAUNCHED();
    CK(lin_fwd(w.cat, p + L.m0_w, p + L.m0_b, w.hid, B, c->hidden, fs + c->ext_dim, 1, st));
    CK(lin_fwd(w.hid, p + L.m2_w, p + L.m2_b, pred, B, 1, c->hidden, 0, st));
  }
  return MIVIT_OK;
}

extern "C" int mivit_resnet_backward(const mivit_resnet_config* c, int32_t B, const float* x, const float* ext, const float* dpred,
                                     const float* params, float* grads, void* workspace, void* stream) {
  RLayout L;
  CK(r_layout(c, L));
  MIVIT_CHECK_ARG(B >= 1 && x && dpred && params && grads && workspace, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  RWS w;
  r_carve(c, B, workspace, w);
  const float* p = params;
  float* g = grads;
  const int NF = w.NF, fs = c->feature_size;
  MIVIT_CUDA_CHECK(cudaMemsetAsync(g, 0, (size_t)L.off * sizeof(float), st));
  // head
  if (c->ext_dim == 0) {
    const float* d_frame = dpred;
    if (c->single_prediction) {
      meanpool_bwd_kernel<<<nblk(NF), 256, 0, st>>>(dpred, w.d_ppred, B, c->F, 1);
      LAUNCHED();
      d_frame = w.d_ppred;
    }
    CK(lin_bwd(w.f1, p + L.fc2_w, d_frame, g + L.fc2_w, g + L.fc2_b, w.d_f1, NF, 1, fs, st));
  } else {
    CK(lin_bwd(w.hid, p + L.m2_w, dpred, g + L.m2_w, g + L.m2_b, w.d_hid, B, 1, c->hidden, st));
    relu_bwd_kernel<<<nblk((long long)B * c->hidden), 256, 0, st>>>(w.d_hid, w.hid, w.d_hid, (long long)B * c->hidden);
    LAUNCHED();
    CK(lin_bwd(w.cat, p + L.m0_w, w.d_hid, g + L.m0_w, g + L.m0_b, w.d_cat, B, c->hidden, fs + c->ext_dim, st));
    take_cols_kernel<<<nblk((long long)B * fs), 256, 0, st>>>(w.d_cat, w.d_fmean, B, fs + c->ext_dim, fs);
    LAUNCHED();
    meanpool_bwd_kernel<<<nblk((long long)NF * fs), 256, 0, st>>>(w.d_fmean, w.d_f1, B, c->F, fs);
    LAUNCHED();
  }
  relu_bwd_kernel<<<nblk((long long)NF * fs), 256, 0, st>>>(w.d_f1, w.f1, w.d_f1, (long long)NF * fs);
  LAUNCHED();
  CK(lin_bwd(w.pooled, p + L.fc1_w, w.d_f1, g + L.fc1_w, g + L.fc1_b, w.d_pooled, NF, fs, 128, st));
  const int HW = w.blk[2].g2.Ho * w.blk[2].g2.Wo;
  meanpool_bwd_kernel<<<nblk((long long)NF * HW * 128), 256, 0, st>>>(w.d_pooled, w.d_last, NF, HW, 128);
  LAUNCHED();
  const float* d_out = w.d_last;                 // gradient w.r.t. the block output
  for (int k = 2; k >= 0; --k) {
    BlockWS& q = w.blk[k];
    const BlockP& Bp = L.blk[k];
    const int co = kCh[k + 1];
    const float* in = k == 0 ? w.pool : w.blk[k - 1].out;
    const long long rows = (long long)NF * q.g1.Ho * q.g1.Wo;
    // out = relu(bn2(c2raw) + shortcut): gm = d_out * [out > 0] feeds bn2's backward and the shortcut
    CK(bn_bwd_r(d_out, q.out, q.c2raw, q.bn2, p + Bp.bn2_g, q.d_c2raw, q.d_out_m, g + Bp.bn2_g, g + Bp.bn2_b, w.bn_sums, rows, st));
    CK(conv_wgrad(q.c1act, q.d_c2raw, w.dwr, g + Bp.c2_w, q.g2, st));
    CK(conv_dgrad(q.d_c2raw, q.wr2, q.d_c1act, q.g2, st));
    CK(bn_bwd_r(q.d_c1act, q.c1act, q.c1raw, q.bn1, p + Bp.bn1_g, q.d_c1raw, nullptr, g + Bp.bn1_g, g + Bp.bn1_b, w.bn_sums, rows, st));
    CK(conv_wgrad(in, q.d_c1raw, w.dwr, g + Bp.c1_w, q.g1, st));
    CK(conv_dgrad(q.d_c1raw, q.wr1, q.d_in, q.g1, st));
    const long long n_in = (long long)NF * q.g1.Hi * q.g1.Wi * q.g1.Ci;
    if (q.has_sc) {
      CK(bn_bwd_r(q.d_out_m, nullptr, q.scraw, q.bns, p + Bp.scbn_g, q.d_scraw, nullptr, g + Bp.scbn_g, g + Bp.scbn_b, w.bn_sums, rows, st));
      CK(conv_wgrad(in, q.d_scraw, w.dwr, g + Bp.sc_w, q.gs, st));
      CK(conv_dgrad(q.d_scraw, q.wrs, q.d_in_sc, q.gs, st));      // the 1x1 shortcut's input gradient is added to conv1's
      add_inplace_kernel<<<nblk(n_in), 256, 0, st>>>(q.d_in, q.d_in_sc, n_in);
      LAUNCHED();
    } else {
      add_inplace_kernel<<<nblk(n_in), 256, 0, st>>>(q.d_in, q.d_out_m, n_in);   // identity shortcut
      LAUNCHED();
    }
    d_out = q.d_in;
    (void)co;
  }
  // stem: maxpool -> ReLU -> BN -> conv 5x5
  const long long n1 = (long long)NF * w.H1 * w.H1;
  MIVIT_CUDA_CHECK(cudaMemsetAsync(w.d_a1, 0, (size_t)n1 * 32 * sizeof(float), st));
  maxpool_bwd_kernel<<<nblk((long long)NF * w.H2 * w.H2 * 32), 256, 0, st>>>(d_out, w.pool_idx, w.d_a1, NF, w.H1, w.H1, w.H2, w.H2, 32);
  LAUNCHED();
  CK(bn_bwd_r(w.d_a1, w.a1, w.a1raw, w.bn0, p + L.bn1_g, w.d_a1raw, nullptr, g + L.bn1_g, g + L.bn1_b, w.bn_sums, n1, st));
  CK(conv_wgrad(x, w.d_a1raw, w.dwr, g + L.conv1_w, w.g0, st));
  return MIVIT_OK;
}

extern "C" int mivit_resnet_train_step(const mivit_resnet_config* c, int32_t B, const float* x, const float* ext, const float* target,
                                       float* params, float* grads, float* adam_m, float* adam_v, float* bn_running,
                                       int64_t* bn_num_batches, void* workspace, float* pred, float* loss, float* dpred, float lr,
                                       float beta1, float beta2, float eps, float weight_decay, int64_t step, int32_t apply_update,
                                       void* stream) {
  RLayout L;
  CK(r_layout(c, L));
  MIVIT_CHECK_ARG(target && pred && loss && dpred, "bad arguments");
  CK(mivit_resnet_forward(c, B, x, ext, params, bn_running, bn_num_batches, workspace, pred, 1, stream));
  CK(mse_loss(pred, target, mivit_resnet_pred_rows(c, B), loss, dpred, (cudaStream_t)stream));
  CK(mivit_resnet_backward(c, B, x, ext, dpred, params, grads, workspace, stream));
  if (apply_update) {
    MIVIT_CHECK_ARG(adam_m && adam_v && step >= 1, "optimizer state missing");
    CK(adamw_flat(params, grads, adam_m, adam_v, L.off, lr, beta1, beta2, eps, weight_decay, step, 1.0f, (cudaStream_t)stream));
  }
  return MIVIT_OK;
}
