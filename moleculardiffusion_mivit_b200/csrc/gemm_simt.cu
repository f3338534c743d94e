// Generic fp32 tiled GEMM for the small, memory/latency-bound matrices of the transformer
// stack and heads (K, N in {1..256}); reference ops: nn.Linear forward/backward in
// helpers/models.py:20-23, 64-65, 151, 241, 268-273, 316-320.
//   C[m,n] (+)= sum_k A(m,k) * B(k,n)  (+ bias[n])  (relu)
// with arbitrary element strides, so one kernel covers X*W^T (forward), dY*W (input grad) and
// dY^T*X (weight grad, split-K over blockIdx.z with atomics).
#include "common.cuh"
#include "vit.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool ATOMIC>
__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ A, long long sam, long long sak,
                                                   const float* __restrict__ B, long long sbk, long long sbn,
                                                   float* __restrict__ C, long long scm, int M, int N, int K,
                                                   const float* __restrict__ bias, int relu, int accumulate, int k_chunk) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * k_chunk, ke = min(K, kb + k_chunk);
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  float acc[4][4] = {};
  for (int k0 = kb; k0 < ke; k0 += BK) {
    // A tile: BM x BK.  Pick the thread->element map that walks the contiguous stride.
    for (int i = tid; i < BM * BK; i += 256) {
      int mm, kk;
      if (sak == 1) { kk = i % BK; mm = i / BK; } else { mm = i % BM; kk = i / BM; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < ke) ? A[gm * sam + gk * sak] : 0.f;
    }
    for (int i = tid; i < BN * BK; i += 256) {
      int nn, kk;
      if (sbk == 1) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < ke) ? B[gk * sbk + gn * sbn] : 0.f;
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + tm + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tn + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      float* dst = C + gm * scm + gn;
      if (ATOMIC) {
        atomicAdd(dst, v);
      } else {
        if (bias) v += bias[gn];
        if (accumulate) v += *dst;
        if (relu) v = fmaxf(v, 0.f);
        *dst = v;
      }
    }
  }
}

// out[n] (+)= sum_m X[m, n]   (bias gradients)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, long long ld, int M, int N,
                                                     float* __restrict__ out, int rows_per_block) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int m = r0 + (threadIdx.x >> 5); m < r1; m += 8) s += X[m * ld + n];
  __shared__ float sh[8][33];
  sh[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && n < N) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
    atomicAdd(out + n, t);
  }
}

}  // namespace

int gemm_f32(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
             long long scm, int M, int N, int K, const float* bias, int relu, int accumulate, int split_k,
             cudaStream_t st) {
  if (M <= 0 || N <= 0) return MIVIT_OK;
  dim3 grid(mivit_ceil_div(N, BN), mivit_ceil_div(M, BM), 1);
  if (split_k > 1) {
    // caller guarantees C is zero-initialised (or holds the value to accumulate onto)
    int chunk = mivit_ceil_div(K, split_k);
    chunk = (chunk + BK - 1) / BK * BK;
    grid.z = mivit_ceil_div(K, chunk);
    gemm_kernel<true><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, scm, M, N, K, nullptr, 0, 0, chunk);
  } else {
    gemm_kernel<false><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, scm, M, N, K, bias, relu, accumulate, K);
  }
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

int colsum_f32(const float* X, long long ld, int M, int N, float* out, cudaStream_t st) {
  if (M <= 0 || N <= 0) return MIVIT_OK;
  const int rpb = 128;  // many small CTAs: the reduction is bandwidth-bound and M is 30k+ tokens
  dim3 grid(mivit_ceil_div(N, 32), mivit_ceil_div(M, rpb));
  colsum_kernel<<<grid, 256, 0, st>>>(X, ld, M, N, out, rpb);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}
