// Token-level kernels of the frame-token transformer (reference helpers/models.py):
//   LayerNorm (+ residual)          :88-89,101,106,134,301,334   layernorm_fwd / layernorm_bwd
//   softmax(QK^T/sqrt(d)) V         :37-54                       attention_fwd / attention_bwd
//   activation of the FFN           :72-77                       act_fwd / act_bwd
//   regression token / pos-encoding :138,338-347                 tokens_finish / tokens_finish_bwd
//   x[:,0] or x.mean(1)             :351-354                     pool_tokens / pool_tokens_bwd
// Sequences are tiny (S <= 128 tokens, head_dim 16): one CTA handles one (sequence, head) with
// everything in shared memory; LayerNorm is one warp per token.
#include "common.cuh"
#include "vit.h"

namespace {

// ---------------------------------------------------------------- LayerNorm -------------
// z = x (+ res); y = (z - mean) * rstd * gamma + beta.   Row r of the input maps to output
// row (r / group) * out_group + out_off + r % group  (lets the embedding LN write token slots).
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ z_out, float* __restrict__ y,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int rows, int E, float eps, int group, int out_group, int out_off) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* xr = x + (size_t)r * E;
  const float* rr = res ? res + (size_t)r * E : nullptr;
  float v[8];
  float s = 0.f;
  const int per = (E + 31) / 32;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    v[i] = 0.f;
    if (i < per && c < E) {
      v[i] = xr[c] + (rr ? rr[c] : 0.f);
      s += v[i];
    }
  }
  const float mean = warp_sum(s) / (float)E;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (i < per && c < E) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)E + eps);
  const size_t orow = (size_t)(r / group) * out_group + out_off + (r % group);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (i < per && c < E) {
      if (z_out) z_out[(size_t)r * E + c] = v[i];
      y[orow * E + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
    }
  }
  if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
}

// dz = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dgamma += dy*xhat;  dbeta += dy.
// dy row mapping mirrors the forward output mapping.
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, float* __restrict__ dz,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, int rows,
                                                            int E, int group, int out_group, int out_off) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int per = (E + 31) / 32;
  float ag[8] = {}, ab[8] = {};
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const size_t orow = (size_t)(r / group) * out_group + out_off + (r % group);
    const float m = mean[r], rs = rstd[r];
    float xh[8], g[8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      xh[i] = g[i] = 0.f;
      if (i < per && c < E) {
        const float d = dy[orow * E + c];
        xh[i] = (z[(size_t)r * E + c] - m) * rs;
        g[i] = d * gamma[c];
        s1 += g[i];
        s2 = fmaf(g[i], xh[i], s2);
        ag[i] = fmaf(d, xh[i], ag[i]);
        ab[i] += d;
      }
    }
    s1 = warp_sum(s1) / (float)E;
    s2 = warp_sum(s2) / (float)E;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (i < per && c < E) dz[(size_t)r * E + c] = rs * (g[i] - s1 - xh[i] * s2);
    }
  }
  // block-level reduction of the parameter gradients, then one atomic per column per block
  __shared__ float sg[8][256], sb[8][256];
  const int w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (i < per && c < E) { sg[w][c] = ag[i]; sb[w][c] = ab[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    float tg = 0.f, tb = 0.f;
    for (int k = 0; k < wpb; ++k) { tg += sg[k][c]; tb += sb[k][c]; }
    atomicAdd(dgamma + c, tg);
    atomicAdd(dbeta + c, tb);
  }
}

// ---------------------------------------------------------------- attention -------------
// q,k,v: [B,S,E] fp32 with head h in columns [h*d, (h+1)*d).  One CTA per (b,h).
template <int D>
__global__ void __launch_bounds__(128) attention_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                            const float* __restrict__ v, float* __restrict__ ctx,
                                                            float* __restrict__ probs, int S, int E, int H, float scale) {
  extern __shared__ float sm[];
  float* Ks = sm;                  // [S][D+1]
  float* Vs = Ks + S * (D + 1);    // [S][D+1]
  float* Ps = Vs + S * (D + 1);    // [S][S+1]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const size_t base = (size_t)b * S * E + (size_t)h * D;
  for (int i = threadIdx.x; i < S * D; i += blockDim.x) {
    const int s = i / D, c = i % D;
    Ks[s * (D + 1) + c] = k[base + (size_t)s * E + c];
    Vs[s * (D + 1) + c] = v[base + (size_t)s * E + c];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    float qi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) qi[c] = q[base + (size_t)i * E + c];
    float mx = -INFINITY;
    for (int j = 0; j < S; ++j) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) a = fmaf(qi[c], Ks[j * (D + 1) + c], a);
      a *= scale;
      Ps[i * (S + 1) + j] = a;
      mx = fmaxf(mx, a);
    }
    float sum = 0.f;
    for (int j = 0; j < S; ++j) {
      const float e = expf(Ps[i * (S + 1) + j] - mx);
      Ps[i * (S + 1) + j] = e;
      sum += e;
    }
    const float inv = 1.0f / sum;
    float o[D] = {};
    float* prow = probs + ((size_t)blockIdx.x * S + i) * S;
    for (int j = 0; j < S; ++j) {
      const float p = Ps[i * (S + 1) + j] * inv;
      prow[j] = p;
#pragma unroll
      for (int c = 0; c < D; ++c) o[c] = fmaf(p, Vs[j * (D + 1) + c], o[c]);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) ctx[base + (size_t)i * E + c] = o[c];
  }
}

template <int D>
__global__ void __launch_bounds__(128) attention_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                            const float* __restrict__ v, const float* __restrict__ probs,
                                                            const float* __restrict__ dctx, float* __restrict__ dq,
                                                            float* __restrict__ dk, float* __restrict__ dv, int S, int E,
                                                            int H, float scale) {
  extern __shared__ float sm[];
  float* Qs = sm;                   // [S][D+1]
  float* Ks = Qs + S * (D + 1);
  float* Vs = Ks + S * (D + 1);
  float* Gs = Vs + S * (D + 1);     // dctx
  float* Ps = Gs + S * (D + 1);     // [S][S+1] probabilities
  float* Ds = Ps + S * (S + 1);     // [S][S+1] dS
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const size_t base = (size_t)b * S * E + (size_t)h * D;
  for (int i = threadIdx.x; i < S * D; i += blockDim.x) {
    const int s = i / D, c = i % D;
    const size_t g = base + (size_t)s * E + c;
    Qs[s * (D + 1) + c] = q[g];
    Ks[s * (D + 1) + c] = k[g];
    Vs[s * (D + 1) + c] = v[g];
    Gs[s * (D + 1) + c] = dctx[g];
  }
  for (int i = threadIdx.x; i < S * S; i += blockDim.x)
    Ps[(i / S) * (S + 1) + (i % S)] = probs[(size_t)blockIdx.x * S * S + i];
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    float gi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) gi[c] = Gs[i * (D + 1) + c];
    float rowdot = 0.f;
    for (int j = 0; j < S; ++j) {
      float dp = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) dp = fmaf(gi[c], Vs[j * (D + 1) + c], dp);
      Ds[i * (S + 1) + j] = dp;
      rowdot = fmaf(dp, Ps[i * (S + 1) + j], rowdot);
    }
    float a[D] = {};
    for (int j = 0; j < S; ++j) {
      const float ds = Ps[i * (S + 1) + j] * (Ds[i * (S + 1) + j] - rowdot);
      Ds[i * (S + 1) + j] = ds;
#pragma unroll
      for (int c = 0; c < D; ++c) a[c] = fmaf(ds, Ks[j * (D + 1) + c], a[c]);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) dq[base + (size_t)i * E + c] = a[c] * scale;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    float ak[D] = {}, av[D] = {};
    for (int i = 0; i < S; ++i) {
      const float ds = Ds[i * (S + 1) + j], p = Ps[i * (S + 1) + j];
#pragma unroll
      for (int c = 0; c < D; ++c) {
        ak[c] = fmaf(ds, Qs[i * (D + 1) + c], ak[c]);
        av[c] = fmaf(p, Gs[i * (D + 1) + c], av[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < D; ++c) {
      dk[base + (size_t)j * E + c] = ak[c] * scale;
      dv[base + (size_t)j * E + c] = av[c];
    }
  }
}

// ---- sequence-per-CTA variants (S <= 64): one warp per head, lanes over query (forward, dQ) or key (dK, dV)
// rows.  Flash-attention style bookkeeping: the forward keeps the S scores of a row in registers (one QK^T pass)
// and stores only the row log-sum-exp L_i (in the first B*H*S floats of `probs`); the backward recomputes
// P_ij = exp(s_ij - L_i) and uses D_i = dctx_i . ctx_i, so no S x S matrix ever touches shared memory or HBM.
template <int D>
__device__ __forceinline__ float dot_row(const float (&a)[D], const float* __restrict__ b) {
  float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;   // four independent chains
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(b + c);
    p0 = fmaf(a[c], t.x, p0);
    p1 = fmaf(a[c + 1], t.y, p1);
    p2 = fmaf(a[c + 2], t.z, p2);
    p3 = fmaf(a[c + 3], t.w, p3);
  }
  return (p0 + p1) + (p2 + p3);
}
template <int D>
__device__ __forceinline__ void axpy_row(float w, const float* __restrict__ b, float (&acc)[D]) {
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(b + c);
    acc[c] = fmaf(w, t.x, acc[c]);
    acc[c + 1] = fmaf(w, t.y, acc[c + 1]);
    acc[c + 2] = fmaf(w, t.z, acc[c + 2]);
    acc[c + 3] = fmaf(w, t.w, acc[c + 3]);
  }
}
template <int D>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&r)[D], float mul) {
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + c);
    r[c] = t.x * mul; r[c + 1] = t.y * mul; r[c + 2] = t.z * mul; r[c + 3] = t.w * mul;
  }
}
template <int D>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&r)[D], float mul) {
#pragma unroll
  for (int c = 0; c < D; c += 4) *reinterpret_cast<float4*>(p + c) = make_float4(r[c] * mul, r[c + 1] * mul, r[c + 2] * mul, r[c + 3] * mul);
}

template <int D, int SMAX>
__global__ void __launch_bounds__(256) attention_fwd_seq_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                                const float* __restrict__ v, float* __restrict__ ctx,
                                                                float* __restrict__ lse, int S, int E, int H, float scale) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;            // [S][E]
  float* Vs = sm + S * E;    // [S][E]
  const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = (size_t)b * S * E;
  for (int i = threadIdx.x; i < S * E / 4; i += blockDim.x) {
    reinterpret_cast<float4*>(Ks)[i] = __ldg(reinterpret_cast<const float4*>(k + base) + i);
    reinterpret_cast<float4*>(Vs)[i] = __ldg(reinterpret_cast<const float4*>(v + base) + i);
  }
  __syncthreads();
  if (h >= H) return;
  const int hc = h * D;
  for (int i = lane; i < S; i += 32) {
    float qi[D];
    load_row<D>(q + base + (size_t)i * E + hc, qi, scale);
    float sc[SMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SMAX; ++j)
      if (j < S) {
        sc[j] = dot_row<D>(qi, Ks + j * E + hc);
        mx = fmaxf(mx, sc[j]);
      }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < SMAX; ++j)
      if (j < S) {
        sc[j] = expf(sc[j] - mx);
        sum += sc[j];
      }
    float o[D] = {};
#pragma unroll
    for (int j = 0; j < SMAX; ++j)
      if (j < S) axpy_row<D>(sc[j], Vs + j * E + hc, o);
    store_row<D>(ctx + base + (size_t)i * E + hc, o, 1.0f / sum);
    lse[((size_t)b * H + h) * S + i] = mx + logf(sum);
  }
}

template <int D>
__global__ void __launch_bounds__(256) attention_bwd_seq_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                                const float* __restrict__ v, const float* __restrict__ lse,
                                                                const float* __restrict__ ctx, const float* __restrict__ dctx,
                                                                float* __restrict__ dq, float* __restrict__ dk,
                                                                float* __restrict__ dv, int S, int E, int H, float scale) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                 // [S][E]
  float* Ks = Qs + S * E;
  float* Vs = Ks + S * E;
  float* Gs = Vs + S * E;         // dctx
  float* Ls = Gs + S * E;         // [H][S] row log-sum-exp
  float* Dl = Ls + H * S;         // [H][S] D_i = dctx_i . ctx_i
  const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = (size_t)b * S * E;
  for (int i = threadIdx.x; i < S * E / 4; i += blockDim.x) {
    reinterpret_cast<float4*>(Qs)[i] = __ldg(reinterpret_cast<const float4*>(q + base) + i);
    reinterpret_cast<float4*>(Ks)[i] = __ldg(reinterpret_cast<const float4*>(k + base) + i);
    reinterpret_cast<float4*>(Vs)[i] = __ldg(reinterpret_cast<const float4*>(v + base) + i);
    reinterpret_cast<float4*>(Gs)[i] = __ldg(reinterpret_cast<const float4*>(dctx + base) + i);
  }
  __syncthreads();
  if (h >= H) return;
  const int hc = h * D;
  // pass A (lane = query row i): D_i, then dQ_i = scale * sum_j dS_ij K_j
  for (int i = lane; i < S; i += 32) {
    float gi[D], qi[D], oi[D];
    load_row<D>(Gs + i * E + hc, gi, 1.f);
    load_row<D>(Qs + i * E + hc, qi, scale);
    load_row<D>(ctx + base + (size_t)i * E + hc, oi, 1.f);
    float di = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) di = fmaf(gi[c], oi[c], di);
    const float li = lse[((size_t)b * H + h) * S + i];
    Ls[h * S + i] = li;
    Dl[h * S + i] = di;
    float a[D] = {};
    for (int j = 0; j < S; ++j) {
      const float p = expf(dot_row<D>(qi, Ks + j * E + hc) - li);
      const float ds = p * (dot_row<D>(gi, Vs + j * E + hc) - di);
      axpy_row<D>(ds, Ks + j * E + hc, a);
    }
    store_row<D>(dq + base + (size_t)i * E + hc, a, scale);
  }
  __syncwarp();
  // pass B (lane = key row j): dK_j = scale * sum_i dS_ij Q_i,  dV_j = sum_i P_ij dctx_i
  for (int j = lane; j < S; j += 32) {
    float kj[D], vj[D];
    load_row<D>(Ks + j * E + hc, kj, scale);
    load_row<D>(Vs + j * E + hc, vj, 1.f);
    float ak[D] = {}, av[D] = {};
    for (int i = 0; i < S; ++i) {
      const float p = expf(dot_row<D>(kj, Qs + i * E + hc) - Ls[h * S + i]);
      const float ds = p * (dot_row<D>(vj, Gs + i * E + hc) - Dl[h * S + i]);
      axpy_row<D>(ds, Qs + i * E + hc, ak);
      axpy_row<D>(p, Gs + i * E + hc, av);
    }
    store_row<D>(dk + base + (size_t)j * E + hc, ak, scale);
    store_row<D>(dv + base + (size_t)j * E + hc, av, 1.f);
  }
}

// ---------------------------------------------------------------- activations -----------
__device__ __forceinline__ float act_apply(float x, int mode) {
  if (mode == 0) return fmaxf(x, 0.f);
  if (mode == 1) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));   // F.gelu (exact erf form)
  return x > 0.f ? x : 0.01f * x;                                             // F.leaky_relu default slope
}
__device__ __forceinline__ float act_grad(float x, int mode) {
  if (mode == 0) return x > 0.f ? 1.f : 0.f;
  if (mode == 1) return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.39894228040143268f * expf(-0.5f * x * x);
  return x > 0.f ? 1.f : 0.01f;
}
__global__ void act_fwd_kernel(const float* __restrict__ pre, float* __restrict__ post, long long n, int mode) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) post[i] = act_apply(pre[i], mode);
}
__global__ void act_bwd_kernel(const float* __restrict__ dpost, const float* __restrict__ pre, float* __restrict__ dpre,
                               long long n, int mode) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dpre[i] = dpost[i] * act_grad(pre[i], mode);
}

// torch.nan_to_num(features, nan=0.0) of ModularTransformer.forward (helpers/models.py:535,550): NaN -> 0,
// +inf / -inf -> the largest / smallest finite float32 (torch's defaults for posinf / neginf).
__global__ void nan_to_num_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = in[i];
  if (v != v) v = 0.f;
  else if (isinf(v)) v = v > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
  out[i] = v;
}
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// ---------------------------------------------------------------- token assembly --------
// tokens[b,0,:] = reg (+ proj[b,:]) when use_reg;  tokens[b,s,:] += pos[s,:] when use_pos.
__global__ void tokens_finish_kernel(float* __restrict__ tok, const float* __restrict__ reg, const float* __restrict__ proj,
                                     const float* __restrict__ pos, int B, int S, int E) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * S * E) return;
  const int c = (int)(i % E);
  const int s = (int)((i / E) % S);
  const int b = (int)(i / ((long long)E * S));
  float v = tok[i];
  if (reg != nullptr && s == 0) v = reg[c] + (proj ? proj[(size_t)b * E + c] : 0.f);
  if (pos != nullptr) v += pos[(size_t)s * E + c];
  tok[i] = v;
}
// dreg[c] += sum_b dtok[b,0,c]; dproj[b,c] = dtok[b,0,c]; dpos[s,c] += sum_b dtok[b,s,c]
// (gradient buffers are pre-zeroed; blockIdx.y splits the batch, partial sums go through atomics)
__global__ void tokens_finish_bwd_kernel(const float* __restrict__ dtok, float* __restrict__ dreg, float* __restrict__ dproj,
                                         float* __restrict__ dpos, int B, int S, int E, int b_chunk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over S*E (or E when only the regression token matters)
  const int cols = dpos != nullptr ? S * E : E;
  if (i >= cols) return;
  const int c = i % E, s = i / E;
  const int b0 = blockIdx.y * b_chunk, b1 = min(B, b0 + b_chunk);
  float acc = 0.f;
  for (int b = b0; b < b1; ++b) {
    const float g = dtok[((size_t)b * S + s) * E + c];
    acc += g;
    if (dproj != nullptr && s == 0) dproj[(size_t)b * E + c] = g;
  }
  if (dpos != nullptr) atomicAdd(dpos + i, acc);
  if (dreg != nullptr && s == 0) atomicAdd(dreg + c, acc);
}

// out[b,:] = x[b,0,:] (use_reg) or mean_s x[b,s,:]; written at out[b*ld + c]
__global__ void pool_tokens_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int S, int E, int ld,
                                   int use_reg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * E) return;
  const int b = i / E, c = i % E;
  float v;
  if (use_reg) {
    v = x[((size_t)b * S) * E + c];
  } else {
    v = 0.f;
    for (int s = 0; s < S; ++s) v += x[((size_t)b * S + s) * E + c];
    v /= (float)S;
  }
  out[(size_t)b * ld + c] = v;
}
__global__ void pool_tokens_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int B, int S, int E, int ld,
                                       int use_reg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * S * E) return;
  const int c = (int)(i % E);
  const int s = (int)((i / E) % S);
  const int b = (int)(i / ((long long)E * S));
  const float g = dout[(size_t)b * ld + c];
  dx[i] = use_reg ? (s == 0 ? g : 0.f) : g / (float)S;
}

}  // namespace

#define LAUNCH_1D(kern, n, ...)                                                     \
  do {                                                                              \
    if ((n) > 0) {                                                                  \
      kern<<<mivit_ceil_div((n), 256), 256, 0, st>>>(__VA_ARGS__);                  \
      mivit_count_launch();                                                         \
      MIVIT_LAUNCH_CHECK();                                                         \
    }                                                                               \
  } while (0)

int layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* z_out, float* y,
                  float* mean, float* rstd, int rows, int E, float eps, int group, int out_group, int out_off,
                  cudaStream_t st) {
  MIVIT_CHECK_ARG(E <= 256, "LayerNorm width %d > 256", E);
  if (rows <= 0) return MIVIT_OK;
  if (group <= 0) { group = rows; out_group = rows; out_off = 0; }
  layernorm_fwd_kernel<<<mivit_ceil_div(rows, 8), 256, 0, st>>>(x, res, gamma, beta, z_out, y, mean, rstd, rows, E, eps, group,
                                                                 out_group, out_off);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

int layernorm_bwd(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma, float* dz,
                  float* dgamma, float* dbeta, int rows, int E, int group, int out_group, int out_off, cudaStream_t st) {
  MIVIT_CHECK_ARG(E <= 256, "LayerNorm width %d > 256", E);
  if (rows <= 0) return MIVIT_OK;
  if (group <= 0) { group = rows; out_group = rows; out_off = 0; }
  int blocks = mivit_ceil_div(rows, 8);
  if (blocks > 592) blocks = 592;
  layernorm_bwd_kernel<<<blocks, 256, 0, st>>>(dy, z, mean, rstd, gamma, dz, dgamma, dbeta, rows, E, group, out_group, out_off);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

// the sequence-per-CTA kernels keep log-sum-exp rows in `probs`; forward and backward must take the same branch
static bool attention_seq_ok(int S, int E, int H) {
  return S <= 64 && H <= 8 && E % 4 == 0 && (E / H) % 4 == 0 && (size_t)(4 * S * E + 2 * H * S) * sizeof(float) <= 200 * 1024;
}

// floats the `probs` buffer of one layer needs: row log-sum-exps for the sequence-per-CTA kernels, the full S x S matrices otherwise
size_t attention_probs_floats(int B, int S, int E, int H) {
  return attention_seq_ok(S, E, H) ? (size_t)B * H * S : (size_t)B * H * S * S;
}

template <int D>
static int attention_fwd_d(const float* q, const float* k, const float* v, float* ctx, float* probs, int B, int S, int E,
                           int H, cudaStream_t st) {
  if (attention_seq_ok(S, E, H)) {
    const size_t smem = (size_t)2 * S * E * sizeof(float);
    MivitProfScope prof("attention_fwd", 4.0 * B * H * S * S * D, st);
    if (S <= 32) {
      auto kern = attention_fwd_seq_kernel<D, 32>;
      MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<B, 32 * H, smem, st>>>(q, k, v, ctx, probs, S, E, H, 1.0f / sqrtf((float)D));
    } else {
      auto kern = attention_fwd_seq_kernel<D, 64>;
      MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<B, 32 * H, smem, st>>>(q, k, v, ctx, probs, S, E, H, 1.0f / sqrtf((float)D));
    }
    mivit_count_launch();
    MIVIT_LAUNCH_CHECK();
    return MIVIT_OK;
  }
  const size_t smem = (size_t)(2 * S * (D + 1) + S * (S + 1)) * sizeof(float);
  auto kern = attention_fwd_kernel<D>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<B * H, 128, smem, st>>>(q, k, v, ctx, probs, S, E, H, 1.0f / sqrtf((float)D));
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}
template <int D>
static int attention_bwd_d(const float* q, const float* k, const float* v, const float* probs, const float* ctx, const float* dctx,
                           float* dq, float* dk, float* dv, int B, int S, int E, int H, cudaStream_t st) {
  if (attention_seq_ok(S, E, H)) {
    const size_t smem = (size_t)(4 * S * E + 2 * H * S) * sizeof(float);
    MivitProfScope prof("attention_bwd", 14.0 * B * H * S * S * D, st);
    auto kern = attention_bwd_seq_kernel<D>;
    MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, 32 * H, smem, st>>>(q, k, v, probs, ctx, dctx, dq, dk, dv, S, E, H, 1.0f / sqrtf((float)D));
    mivit_count_launch();
    MIVIT_LAUNCH_CHECK();
    return MIVIT_OK;
  }
  const size_t smem = (size_t)(4 * S * (D + 1) + 2 * S * (S + 1)) * sizeof(float);
  auto kern = attention_bwd_kernel<D>;
  MIVIT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<B * H, 128, smem, st>>>(q, k, v, probs, dctx, dq, dk, dv, S, E, H, 1.0f / sqrtf((float)D));
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

int attention_fwd(const float* q, const float* k, const float* v, float* ctx, float* probs, int B, int S, int E, int H,
                  cudaStream_t st) {
  MIVIT_CHECK_ARG(S >= 1 && S <= 128, "sequence length %d out of range [1,128] (MAX_TOKENS)", S);
  MIVIT_CHECK_ARG(E % H == 0, "embed_dim must be divisible by num_heads");
  if (B <= 0) return MIVIT_OK;
  switch (E / H) {
    case 8: return attention_fwd_d<8>(q, k, v, ctx, probs, B, S, E, H, st);
    case 16: return attention_fwd_d<16>(q, k, v, ctx, probs, B, S, E, H, st);
    case 32: return attention_fwd_d<32>(q, k, v, ctx, probs, B, S, E, H, st);
    default: mivit_set_error("head_dim %d not supported (8, 16, 32)", E / H); return MIVIT_ERR_INVALID;
  }
}
int attention_bwd(const float* q, const float* k, const float* v, const float* probs, const float* ctx, const float* dctx,
                  float* dq, float* dk, float* dv, int B, int S, int E, int H, cudaStream_t st) {
  if (B <= 0) return MIVIT_OK;
  switch (E / H) {
    case 8: return attention_bwd_d<8>(q, k, v, probs, ctx, dctx, dq, dk, dv, B, S, E, H, st);
    case 16: return attention_bwd_d<16>(q, k, v, probs, ctx, dctx, dq, dk, dv, B, S, E, H, st);
    case 32: return attention_bwd_d<32>(q, k, v, probs, ctx, dctx, dq, dk, dv, B, S, E, H, st);
    default: mivit_set_error("head_dim %d not supported (8, 16, 32)", E / H); return MIVIT_ERR_INVALID;
  }
}

int act_fwd(const float* pre, float* post, long long n, int mode, cudaStream_t st) {
  LAUNCH_1D(act_fwd_kernel, n, pre, post, n, mode);
  return MIVIT_OK;
}
int nan_to_num_f32(const float* in, float* out, long long n, cudaStream_t st) {
  LAUNCH_1D(nan_to_num_kernel, n, in, out, n);
  return MIVIT_OK;
}
int add_f32(const float* a, const float* b, float* out, long long n, cudaStream_t st) {
  LAUNCH_1D(add_kernel, n, a, b, out, n);
  return MIVIT_OK;
}
int act_bwd(const float* dpost, const float* pre, float* dpre, long long n, int mode, cudaStream_t st) {
  LAUNCH_1D(act_bwd_kernel, n, dpost, pre, dpre, n, mode);
  return MIVIT_OK;
}
int tokens_finish(float* tok, const float* reg, const float* proj, const float* pos, int B, int S, int E, cudaStream_t st) {
  if (reg == nullptr && pos == nullptr) return MIVIT_OK;
  const long long n = (long long)B * S * E;
  LAUNCH_1D(tokens_finish_kernel, n, tok, reg, proj, pos, B, S, E);
  return MIVIT_OK;
}
int tokens_finish_bwd(const float* dtok, float* dreg, float* dproj, float* dpos, int B, int S, int E, cudaStream_t st) {
  if (dreg == nullptr && dpos == nullptr && dproj == nullptr) return MIVIT_OK;
  const int cols = dpos != nullptr ? S * E : E;
  const int b_chunk = 32;
  dim3 grid(mivit_ceil_div(cols, 128), mivit_ceil_div(B, b_chunk));
  tokens_finish_bwd_kernel<<<grid, 128, 0, st>>>(dtok, dreg, dproj, dpos, B, S, E, b_chunk);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}
int pool_tokens(const float* x, float* out, int B, int S, int E, int ld, int use_reg, cudaStream_t st) {
  const long long n = (long long)B * E;
  LAUNCH_1D(pool_tokens_kernel, n, x, out, B, S, E, ld, use_reg);
  return MIVIT_OK;
}
int pool_tokens_bwd(const float* dout, float* dx, int B, int S, int E, int ld, int use_reg, cudaStream_t st) {
  const long long n = (long long)B * S * E;
  LAUNCH_1D(pool_tokens_bwd_kernel, n, dout, dx, B, S, E, ld, use_reg);
  return MIVIT_OK;
}
