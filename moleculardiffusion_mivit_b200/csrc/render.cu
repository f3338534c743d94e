// Trajectory -> fluorescence frame renderer (sm_100a).
//
// Replaces the triple Python loop of the reference
//   helpers/helpersGeneration.py:128-278 trajectories_to_video
//   helpers/helpersGeneration.py:283-319 trajectory_to_video
//   helpers/helpersGeneration.py:77-97   gaussian_2d
//   helpers/helpersGeneration.py:356-400 normalize_images (fused, optional)
//   Experiments/PSFNoise/trainSettingsPSFNoise.py:196-309 trajs_to_vid_psf_noise
// by the separable closed form: the reference evaluates a GxG exp
// grid per sub-position and renormalises it by its on-grid maximum; that is exactly
//   I * ay (x) ax,  a[j] = exp(-((x_j-c)^2 - min_j (x_j-c)^2) / (2 sigma^2)),
// so the UxU block mean factorises into two length-P vectors of U-term sums.
//
// One warp renders one frame: (1) sub-position centres in float64 (scale, flip, centring),
// reduced to (nearest grid node, float32 offset); (2) the 2*n*P axis table in shared memory;
// (3) every lane accumulates its pixels over the n sub-positions, adds the clipped Gaussian
// background and the Poisson factor from the pixel's own Philox stream, normalises and
// stores -- consecutive lanes write consecutive pixels (coalesced, frames are contiguous).
#include "common.cuh"
#include <stdlib.h>
#include "philox.cuh"
#include "../../include/mivit.h"
#include "vit.h"

namespace {

constexpr int kMaxVariants = 8;

struct RenderDev {
  double scale, ysign, inv2s2_d, step_d, inv_step_d, inv_n;
  int P, U, n, center, draw, G, limit, T, F;
  float imean, istd;  // per-sub-position intensity mean/std (V1) or per-frame (PSFNoise)
  float bg_mean, bg_std, bg_hi, poisson;
  int normalize, mean_noise;
  float norm_sub, norm_div, inv_norm_div;
  float inv2s2, step;
  uint32_t k0, k1;
  unsigned long long seq_offset;
  const unsigned long long* seq_off_dev;   // optional DEVICE counter added to seq_offset (lets a captured CUDA graph advance the ids)
  long long out_seq_stride;
};
__device__ __forceinline__ uint32_t global_seq(const RenderDev& d, long long s) {
  return (uint32_t)(d.seq_offset + (d.seq_off_dev != nullptr ? *d.seq_off_dev : 0ull) + (unsigned long long)s);
}

struct PsfNoiseDev {
  int n_psf, n_noise;
  float inv2s2[kMaxVariants];
  double inv2s2_d[kMaxVariants];
  float bg_std[kMaxVariants], bg_hi[kMaxVariants];
};

// per-warp shared memory carve-up
struct WarpSmem {
  float* tab;   // [n][2][P]   axis 0 = x (columns), axis 1 = y (rows)
  float* inten; // [n]
  float* c0;    // [2n]  (x: p, y: n+p)
  int* jc;      // [2n]
  float* msq;   // [n]   min squared distance (x+y) in HR pixels^2, for the NaN rule
};

__device__ __forceinline__ WarpSmem carve(float* base, int n, int P) {
  WarpSmem w;
  w.tab = base;
  w.inten = base + 2 * n * P;
  w.c0 = w.inten + n;
  w.jc = reinterpret_cast<int*>(w.c0 + 2 * n);
  w.msq = reinterpret_cast<float*>(w.jc + 2 * n);
  return w;
}
__host__ __device__ inline int warp_smem_floats(int n, int P) { return 2 * n * P + 6 * n; }
// per-warp floats of the fused render->embedding kernel: tables + the rendered frame, both rounded to 16 bytes (the frame is
// read back with 16-byte loads)
__host__ __device__ inline int embed_warp_floats(int n, int P) { return ((warp_smem_floats(n, P) + 3) & ~3) + ((P * P + 3) & ~3); }

// (1) centres of the n sub-positions of frame f  (helpersGeneration.py:289-293)
// frame mean of the scaled positions (the centring offset); every caller sums in the reference's order
__device__ __forceinline__ void frame_mean(const RenderDev& d, const double* __restrict__ seg, double& mx, double& my) {
  mx = 0.0; my = 0.0;
  if (d.center) {
    for (int p = 0; p < d.n; ++p) {  // sequential like np.mean(axis=0)
      mx += seg[2 * p] * d.scale;
      my += seg[2 * p + 1] * d.scale * d.ysign;
    }
    mx *= d.inv_n;   // (a reciprocal multiply instead of np.mean's division: 1 ulp of float64, no double-division subroutine)
    my *= d.inv_n;
  }
}
// sub-position p: (nearest grid node, float32 offset) per axis and the squared offset for the NaN rule
__device__ __forceinline__ void centre_one(const RenderDev& d, const double* __restrict__ seg, double mx, double my, int p,
                                           WarpSmem& w) {
  const double cx = (seg[2 * p] * d.scale - mx) * (double)d.U;
  const double cy = (seg[2 * p + 1] * d.scale * d.ysign - my) * (double)d.U;
  // nearest grid node: any node gives a correct (node, offset) pair, the nearest one keeps the float32 offset small
  int jx = (int)fmin(fmax(rint((cx + d.limit) * d.inv_step_d), 0.0), (double)(d.G - 1));
  int jy = (int)fmin(fmax(rint((cy + d.limit) * d.inv_step_d), 0.0), (double)(d.G - 1));
  const double ox = cx - (-(double)d.limit + jx * d.step_d);
  const double oy = cy - (-(double)d.limit + jy * d.step_d);
  w.jc[p] = jx;
  w.jc[d.n + p] = jy;
  w.c0[p] = (float)ox;
  w.c0[d.n + p] = (float)oy;
  const float fx = (float)ox, fy = (float)oy;
  w.msq[p] = (float)((double)fx * (double)fx + (double)fy * (double)fy);
}
__device__ __forceinline__ void frame_centres(const RenderDev& d, const double* __restrict__ traj_seq, int f,
                                              int lane, WarpSmem& w) {
  const double* seg = traj_seq + (size_t)f * d.n * 2;
  double mx, my;
  frame_mean(d, seg, mx, my);
  for (int p = lane; p < d.n; p += 32) centre_one(d, seg, mx, my, p, w);
}

// (2) axis table: block means of exp(-(k (k - 2 c0)) / 2 sigma^2)
__device__ __forceinline__ void axis_table(const RenderDev& d, float inv2s2, int lane, WarpSmem& w) {
  const int P = d.P, U = d.U, n = d.n;
  const int entries = 2 * n * P;
  const float invU = 1.0f;  // division below (matches np.mean)
  (void)invU;
  for (int e = lane; e < entries; e += 32) {
    const int p = e / (2 * P);
    const int r = e - p * 2 * P;
    const int axis = r / P;
    const int b = r - axis * P;
    const int idx = axis * n + p;
    const int jc = w.jc[idx];
    const float c0 = w.c0[idx];
    float acc = 0.0f;
    for (int u = 0; u < U; ++u) {
      const float k = __fmul_rn((float)(b * U + u - jc), d.step);
      const float t = __fmul_rn(k, __fsub_rn(k, 2.0f * c0));
      acc += expf(__fmul_rn(-t, inv2s2));
    }
    w.tab[(p * 2 + axis) * P + b] = acc / (float)U;
  }
}

__device__ __forceinline__ float pixel_signal(const RenderDev& d, const WarpSmem& w, int a, int b) {
  float acc = 0.0f;
  const int P = d.P;
  for (int p = 0; p < d.n; ++p) {
    const float t = __fmul_rn(w.inten[p], w.tab[(p * 2 + 1) * P + a]);
    acc = __fadd_rn(acc, __fmul_rn(t, w.tab[(p * 2 + 0) * P + b]));
  }
  return acc;
}

// ---- V1 renderer (trajectories_to_video): shared per-frame device code of render_v1_kernel, the fused render->embedding
// kernel and its weight-gradient twin.  Noise layout: "pair" layout of philox.cuh.
struct V1Noise {
  const uint32_t* alias;   // shared-memory copy of the alias table (valid when use_alias)
  int k0;
  bool use_alias;
  PoissonConst pc;         // PTRS fall-back for pn > kAliasMaxLambda
  const float* ptab;
};

// Compile-time frame geometry of the common configurations (0 = read the value from RenderDev at run time): with P, U and n
// known the table / pixel loops unroll, the index divisions vanish and shared-memory offsets become immediates.  ncu on the
// first round-2 version (runtime geometry): 3270 warp instructions per 13 x 13 frame, a third of them address arithmetic.
#define TU_OR(G, dflt) (G::kU ? G::kU : (dflt))
template <int TP, int TU, int TN>
struct Geo {
  static constexpr int kU = TU;
  __device__ __forceinline__ static int P(const RenderDev& d) { return TP ? TP : d.P; }
  __device__ __forceinline__ static int U(const RenderDev& d) { return TU ? TU : d.U; }
  __device__ __forceinline__ static int n(const RenderDev& d) { return TN ? TN : d.n; }
};

// spot intensities of the n sub-positions of frame f  (helpersGeneration.py:300)
__device__ __forceinline__ void v1_intensity_one(const RenderDev& d, WarpSmem& w, int f, int n, int p, uint32_t seq) {
  float z = 0.0f, z1;
  if (!d.mean_noise) {
    const uint4 r = philox4x32_10((uint32_t)(f * n + p), 0u, seq, stream_word(MIVIT_STREAM_INTENSITY, 0), d.k0, d.k1);
    box_muller_fast(r.x, r.y, z, z1);
  }
  float I = __fadd_rn(d.imean, __fmul_rn(d.istd, z));
  if ((double)w.msq[p] * d.inv2s2_d > 745.0) I = __int_as_float(0x7fc00000);  // spot underflows: NaN frame (:305-308)
  w.inten[p] = I;
}
template <class G>
__device__ __forceinline__ void v1_intensities(const RenderDev& d, WarpSmem& w, int f, int lane, uint32_t seq) {
  const int n = G::n(d);
  for (int p = lane; p < n; p += 32) v1_intensity_one(d, w, f, n, p, seq);
}

// axis table of the V1 kernels: tab[p][0][b] = block mean of the x profile, tab[p][1][a] = I_p * block mean of the y profile
// (the intensity is folded into the row factor, so a pixel costs one FMA per sub-position); exps on the MUFU (ex2.approx)
template <class G>
__device__ __forceinline__ void axis_table_v1(const RenderDev& d, int lane, WarpSmem& w) {
  const int P = G::P(d), U = G::U(d), n = G::n(d);
  const float c2 = -d.inv2s2 * 1.4426950408889634f;   // exp(-t/2s^2) = 2^(t * c2)
  const float cs = c2 * d.step;
  const float q = fast_ex2(2.0f * cs * d.step);
  const float invU = 1.0f / (float)U;
  // lane = one (axis, block) entry r of the 2P per sub-position, loop over the sub-positions: the (p, axis, b) decomposition
  // of a flat entry index cost 19 of 51 instructions per entry (ncu source page); here it is done once per lane and the
  // shared-memory offsets of the unrolled loop are immediates.
  for (int r = lane; r < 2 * P; r += 32) {
    const int axis = r >= P ? 1 : 0;
    const int b = r - axis * P;
    const int bU = b * U;
    const float* c0p = w.c0 + axis * n;
    const int* jcp = w.jc + axis * n;
    float* tp = w.tab + r;
#pragma unroll 10
    for (int p = 0; p < n; ++p) {
      const float twoc0 = 2.0f * c0p[p];
      const float kb = (float)(bU - jcp[p]) * d.step;
      // The exponent is quadratic in u, f(u) = c2 (kb + u s)(kb + u s - 2 c0), so consecutive samples differ by a geometric
      // factor whose own ratio q = 2^(2 c2 s^2) is a launch constant: two exps per entry instead of U
      //   e_0 = 2^f(0),  r_0 = 2^(c2 s (2 kb - 2 c0 + s)),  e_(u+1) = e_u r_u,  r_(u+1) = r_u q
      // (relative error <= ~U roundings, 1e-6 at U = 5; explicit roundings: a contracted acc = fma(e, r, acc) would differ
      // between the kernels' instantiations).  Far from the spot e_0 underflows to 0 and stays 0: the exact samples are
      // < 2^-126 there; a ratio beyond 2^126 only occurs where e_0 has underflowed and is clamped so that 0 * r stays 0.
      float e_u = fast_ex2(__fmul_rn(__fmul_rn(kb, __fsub_rn(kb, twoc0)), c2));
      float r_u = fast_ex2(fminf(__fmul_rn(__fsub_rn(fmaf(2.0f, kb, d.step), twoc0), cs), 126.0f));
      float acc = e_u;
#pragma unroll
      for (int u = 1; u < (TU_OR(G, 8)); ++u) {
        if (u < U) {
          e_u = __fmul_rn(e_u, r_u);
          r_u = __fmul_rn(r_u, q);
          acc = __fadd_rn(acc, e_u);
        }
      }
      for (int u = TU_OR(G, 8); u < U; ++u) {   // generic geometry with an upsampling factor beyond 8: rolled remainder
        e_u = __fmul_rn(e_u, r_u);
        r_u = __fmul_rn(r_u, q);
        acc = __fadd_rn(acc, e_u);
      }
      acc *= invU;
      if (axis) acc *= w.inten[p];
      tp[p * 2 * P] = acc;            // tab[(p * 2 + axis) * P + b]
    }
  }
}

// One frame: every lane walks the frame's pixel PAIRS (row a, columns b0 = 2*pr, b0 + 1), computes signal + clipped Gaussian
// background (:312-313), multiplicative Poisson (:316-317), fused normalisation (:395) and hands (pixel index, value) to `sink`.
// No divergent branch on the one-pixel pair that ends a row of odd length: its second column index is clamped and only the
// sink call is predicated.
// kPairLoad: the two x-profile entries of a pixel pair are read as one 8-byte load (the warp's table must be 8-byte aligned:
// true for render_v1_kernel, whose per-warp carve-up has an even number of floats).
template <class G, bool kPairLoad = false, typename Sink>
__device__ __forceinline__ void v1_frame_pixels(const RenderDev& d, const WarpSmem& w, int f, int lane, uint32_t seq,
                                                const V1Noise& nz, Sink&& sink) {
  const int P = G::P(d), n = G::n(d);
  const int ppr = (P + 1) >> 1;              // pairs per row
  const int pairs = ppr * P;
  const float inv_pn = d.poisson != -1.0f ? 1.0f / d.poisson : 1.0f;
  for (int q = lane; q < pairs; q += 32) {
    const int a = q / ppr, b0 = (q - a * ppr) * 2;
    const bool two = b0 + 1 < P;
    float v0 = 0.0f, v1 = 0.0f;
    if (d.draw) {
      const float* ty = w.tab + P + a;                    // tab[p][1][a]
      const float* tx0 = w.tab + b0;                      // tab[p][0][b0]
      const float* tx1 = w.tab + (two ? b0 + 1 : b0);
      if (kPairLoad) {
        // (the pair that ends a row of odd length reads one float past the x profile -- the first y entry -- for a pixel
        //  that is never stored)
#pragma unroll 10
        for (int p = 0; p < n; ++p) {
          const float t = ty[p * 2 * P];
          const float2 x2 = *reinterpret_cast<const float2*>(tx0 + p * 2 * P);
          v0 = fmaf(t, x2.x, v0);
          v1 = fmaf(t, x2.y, v1);
        }
      } else {
#pragma unroll 10
        for (int p = 0; p < n; ++p) {
          const float t = ty[p * 2 * P];
          v0 = fmaf(t, tx0[p * 2 * P], v0);
          v1 = fmaf(t, tx1[p * 2 * P], v1);
        }
      }
    }
    float z0 = 0.0f, z1 = 0.0f, k0f = d.poisson, k1f = d.poisson;
    if (!d.mean_noise) {
      const uint4 r = philox4x32_10((uint32_t)(f * pairs + q), 0u, seq, stream_word(MIVIT_STREAM_PIXEL, 0), d.k0, d.k1);
      box_muller_fast(r.x, r.y, z0, z1);
      if (d.poisson != -1.0f) {
        if (nz.use_alias) {
          k0f = alias_draw(nz.alias, nz.k0, r.z);
          k1f = alias_draw(nz.alias, nz.k0, r.w);
        } else {  // PTRS on the pixel's own uniform stream: blocks 1, 2, ... of item = pixel index
          PixelStream st;
          st.seq = seq; st.sw = stream_word(MIVIT_STREAM_PIXEL, 0); st.k0 = d.k0; st.k1 = d.k1;
          st.item = (uint32_t)(f * P * P + a * P + b0); st.q = 2;
          k0f = poisson_draw_const(nz.pc, nz.ptab, st);
          if (two) { st.item += 1u; st.q = 2; k1f = poisson_draw_const(nz.pc, nz.ptab, st); }
        }
      }
    }
    v0 += fminf(fmaxf(fmaf(d.bg_std, z0, d.bg_mean), 0.0f), d.bg_hi);
    v1 += fminf(fmaxf(fmaf(d.bg_std, z1, d.bg_mean), 0.0f), d.bg_hi);
    if (d.poisson != -1.0f) { v0 = v0 * k0f * inv_pn; v1 = v1 * k1f * inv_pn; }
    if (d.normalize) { v0 = (v0 - d.norm_sub) * d.inv_norm_div; v1 = (v1 - d.norm_sub) * d.inv_norm_div; }
    const int pix = a * P + b0;
    sink(pix, v0);
    if (two) sink(pix + 1, v1);
  }
}

// per-CTA noise set-up: alias table (kernel parameter -> shared memory) or the PTRS log-pmf table
__device__ __forceinline__ V1Noise v1_noise_setup(const RenderDev& d, const AliasTable& at, uint32_t* alias_s, float* ptab) {
  V1Noise nz;
  nz.use_alias = at.valid != 0;
  nz.alias = alias_s;
  nz.k0 = at.k0;
  nz.ptab = ptab;
  nz.pc = PoissonConst{};
  // (the PTRS constants cost ~90 instructions per warp: only when the alias table does not cover pn)
  if (!nz.use_alias && d.poisson != -1.0f && !d.mean_noise) nz.pc = poisson_setup(d.poisson);
  if (d.poisson != -1.0f && !d.mean_noise) {
    if (nz.use_alias) {
      for (int i = threadIdx.x; i < kAliasEntries; i += blockDim.x) alias_s[i] = at.e[i];
    } else if (nz.pc.ptrs) {
      poisson_fill_table(nz.pc, ptab, threadIdx.x, blockDim.x);
    }
  }
  return nz;
}

template <class G>
__device__ __forceinline__ void v1_frame_tables(const RenderDev& d, const double* __restrict__ traj, long long s, int f, int lane,
                                                uint32_t seq, WarpSmem& w) {
  if (!d.draw) return;
  frame_centres(d, traj + (size_t)s * d.T * 2, f, lane, w);
  __syncwarp();
  v1_intensities<G>(d, w, f, lane, seq);
  __syncwarp();
  axis_table_v1<G>(d, lane, w);
  __syncwarp();
}

// (256, 5): 48 registers instead of 56 (12 bytes of spills) buys a fifth resident CTA per SM, which hides the barrier after the
// cooperative set-up phase: 0.0511 -> 0.0494 ms per 30 720 frames; 4- and 2-warp CTAs were slower (0.0533 / 0.0648 ms).
template <int TP, int TU, int TN>
__global__ void __launch_bounds__(256, 5) render_v1_kernel(const double* __restrict__ traj, long long n_frames_total,
                                                        RenderDev d, const __grid_constant__ AliasTable at,
                                                        float* __restrict__ out) {
  using G = Geo<TP, TU, TN>;
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  WarpSmem w = carve(smem + (size_t)warp * warp_smem_floats(G::n(d), G::P(d)), G::n(d), G::P(d));
  __shared__ float ptab[kPoissonTable];
  __shared__ uint32_t alias_s[kAliasEntries];
  const V1Noise nz = v1_noise_setup(d, at, alias_s, ptab);
  // Centres (float64) and spot intensities (one Philox block each) of the CTA's frames, one THREAD per (frame, sub-position):
  // with one warp per frame these two steps ran on n = 10 of 32 lanes in every warp -- 360 of 1820 warp instructions per
  // frame (ncu source page, profiles/r02_ncu_render.md); spread over the CTA they fill 3 warps instead of 8.
  if (d.draw) {
    const int n = G::n(d);
    for (int t = threadIdx.x; t < warps * n; t += blockDim.x) {
      const int wi = t / n, p = t - wi * n;
      const long long gfi = (long long)blockIdx.x * warps + wi;
      if (gfi >= n_frames_total) continue;
      const long long si = gfi / d.F;
      const int fi = (int)(gfi - si * d.F);
      WarpSmem wo = carve(smem + (size_t)wi * warp_smem_floats(n, G::P(d)), n, G::P(d));
      const double* seg = traj + (size_t)si * d.T * 2 + (size_t)fi * n * 2;
      double mx, my;
      frame_mean(d, seg, mx, my);
      centre_one(d, seg, mx, my, p, wo);
      v1_intensity_one(d, wo, fi, n, p, global_seq(d, si));
    }
  }
  __syncthreads();
  const long long gf = (long long)blockIdx.x * warps + warp;  // global frame index
  if (gf >= n_frames_total) return;
  const long long s = gf / d.F;
  const int f = (int)(gf - s * d.F);
  const uint32_t seq = global_seq(d, s);
  if (d.draw) axis_table_v1<G>(d, lane, w);
  __syncwarp();
  float* dst = out + s * d.out_seq_stride + (long long)f * G::P(d) * G::P(d);
  v1_frame_pixels<G, true>(d, w, f, lane, seq, nz, [&](int pix, float v) { dst[pix] = v; });
}

// Renderer fused with the frame embedding of LinearProjectionEmbedding / CNNEmbedding (helpers/models.py:146-199:
// both are emb[f,:] = W[E,P*P] . frame[f] + b): the frame lives only in the warp's shared memory, the kernel writes
// [N,F,E] embeddings (and, optionally, the frames).  Persistent CTAs keep W^T ([P*P][E], so that lanes read consecutive
// output features) resident in shared memory; w_transposed = 0 takes the nn.Linear layout [E][P*P] (the flat parameter
// buffer of the ViT) and transposes while staging.
template <int TP, int TU, int TN>
__global__ void __launch_bounds__(256) render_embed_linear_kernel(const double* __restrict__ traj, long long n_frames_total,
                                                                  RenderDev d, const __grid_constant__ AliasTable at,
                                                                  const float* __restrict__ W, int w_transposed,
                                                                  const float* __restrict__ bias, int E,
                                                                  float* __restrict__ emb, float* __restrict__ frames_out) {
  using G = Geo<TP, TU, TN>;
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int P = G::P(d), n = G::n(d), PP = P * P;
  float* wts = smem;                                       // [PP][E]
  const int per_warp = embed_warp_floats(n, P);
  float* mine = smem + (((size_t)PP * E + 3) & ~(size_t)3) + (size_t)warp * per_warp;
  WarpSmem w = carve(mine, n, P);
  float* px = mine + ((warp_smem_floats(n, P) + 3) & ~3);  // [PP] rendered frame, 16-byte aligned
  if (w_transposed) {
    for (int i = threadIdx.x; i < PP * E; i += blockDim.x) wts[i] = __ldg(W + i);
  } else {
    for (int i = threadIdx.x; i < PP * E; i += blockDim.x) { const int e = i / PP, pix = i - e * PP; wts[pix * E + e] = __ldg(W + i); }
  }
  __shared__ float ptab[kPoissonTable];
  __shared__ uint32_t alias_s[kAliasEntries];
  const V1Noise nz = v1_noise_setup(d, at, alias_s, ptab);
  __syncthreads();
  for (long long gf = (long long)blockIdx.x * warps + warp; gf < n_frames_total; gf += (long long)gridDim.x * warps) {
    const long long s = gf / d.F;
    const int f = (int)(gf - s * d.F);
    const uint32_t seq = global_seq(d, s);
    v1_frame_tables<G>(d, traj, s, f, lane, seq, w);
    float* fo = frames_out != nullptr ? frames_out + s * d.out_seq_stride + (long long)f * PP : nullptr;
    v1_frame_pixels<G>(d, w, f, lane, seq, nz, [&](int pix, float v) {
      px[pix] = v;
      if (fo != nullptr) fo[pix] = v;
    });
    __syncwarp();
    const bool emb_al8 = (reinterpret_cast<unsigned long long>(emb) & 7ull) == 0ull;
    if ((E & 1) == 0) {
      // lane = two adjacent output features: per 4 pixels one 16-byte load of the frame (broadcast), four 8-byte loads of the
      // weights and 8 FMAs (the scalar loop below spends a pixel load, a weight load and an address per FMA: 1350 of the
      // kernel's ~2500 instructions per frame).  Same summation order per output as a K-loop GEMM.
      for (int e = 2 * lane; e < E; e += 64) {
        float a0 = 0.f, a1 = 0.f;
        const float* wp = wts + e;
        int pix = 0;
        for (; pix + 4 <= PP; pix += 4) {
          const float4 f4 = *reinterpret_cast<const float4*>(px + pix);
          const float2 w0 = *reinterpret_cast<const float2*>(wp + (size_t)pix * E);
          const float2 w1 = *reinterpret_cast<const float2*>(wp + (size_t)(pix + 1) * E);
          const float2 w2 = *reinterpret_cast<const float2*>(wp + (size_t)(pix + 2) * E);
          const float2 w3 = *reinterpret_cast<const float2*>(wp + (size_t)(pix + 3) * E);
          a0 = fmaf(f4.x, w0.x, a0); a1 = fmaf(f4.x, w0.y, a1);
          a0 = fmaf(f4.y, w1.x, a0); a1 = fmaf(f4.y, w1.y, a1);
          a0 = fmaf(f4.z, w2.x, a0); a1 = fmaf(f4.z, w2.y, a1);
          a0 = fmaf(f4.w, w3.x, a0); a1 = fmaf(f4.w, w3.y, a1);
        }
        for (; pix < PP; ++pix) {
          const float2 w0 = *reinterpret_cast<const float2*>(wp + (size_t)pix * E);
          a0 = fmaf(px[pix], w0.x, a0); a1 = fmaf(px[pix], w0.y, a1);
        }
        const float o0 = a0 + __ldg(bias + e), o1 = a1 + __ldg(bias + e + 1);
        if (emb_al8) {
          *reinterpret_cast<float2*>(emb + gf * E + e) = make_float2(o0, o1);
        } else {
          emb[gf * E + e] = o0;
          emb[gf * E + e + 1] = o1;
        }
      }
    } else {
      for (int e = lane; e < E; e += 32) {
        float acc = 0.f;
        for (int pix = 0; pix < PP; ++pix) acc = fmaf(px[pix], wts[pix * E + e], acc);   // same summation order as a K-loop GEMM
        emb[gf * E + e] = acc + __ldg(bias + e);
      }
    }
    __syncwarp();
  }
}

// Weight gradient of the fused render->embedding layer: dW[e][pix] += sum_frames demb[frame][e] * frame[pix],
// db[e] += sum_frames demb[frame][e], with the frames RE-RENDERED from the trajectories (same Philox streams, hence
// bit-identical to the forward's frames) instead of being stored: the frames of a Linear / CNN-embedding ViT never exist
// in HBM, forward or backward.  A CTA renders 8 frames (one per warp) into shared memory, then all 256 threads update
// their register tile of dW: warp w owns EPW consecutive output features, lane l the pixels l, l + 32, ...
// KG: the frame geometry is the reference's (U = 5, n = 10, P = 7 / 9 / 13 / 15 as implied by PJ) and compiled in, like in the
// forward kernels (same code, same bits: the re-rendered frames equal the forward's).
template <int EPW, int PJ, bool KG>
__global__ void __launch_bounds__(256) render_embed_wgrad_kernel(const double* __restrict__ traj, long long n_frames_total,
                                                                 RenderDev d, const __grid_constant__ AliasTable at,
                                                                 const float* __restrict__ demb, int E,
                                                                 float* __restrict__ dW, float* __restrict__ db) {
  using G = Geo<KG ? (PJ == 2 ? 7 : PJ == 3 ? 9 : PJ == 6 ? 13 : 15) : 0, KG ? 5 : 0, KG ? 10 : 0>;
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = G::P(d), n = G::n(d), PP = P * P;
  const int per_warp = warp_smem_floats(n, P) + PP;
  float* mine = smem + (size_t)warp * per_warp;
  WarpSmem w = carve(mine, n, P);
  float* px = mine + warp_smem_floats(n, P);               // [PP] rendered frame of this warp
  float* dsh = smem + (size_t)8 * per_warp;                // [8][E] demb rows of the CTA's 8 frames
  __shared__ float ptab[kPoissonTable];
  __shared__ uint32_t alias_s[kAliasEntries];
  const V1Noise nz = v1_noise_setup(d, at, alias_s, ptab);
  float acc[EPW][PJ];
  float bacc = 0.0f;                                       // lane 0..EPW-1 of every warp: bias gradient of feature warp*EPW+lane
#pragma unroll
  for (int i = 0; i < EPW; ++i)
#pragma unroll
    for (int j = 0; j < PJ; ++j) acc[i][j] = 0.0f;
  __syncthreads();
  const int e0 = warp * EPW;
  for (long long g0 = (long long)blockIdx.x * 8; g0 < n_frames_total; g0 += (long long)gridDim.x * 8) {
    const long long gf = g0 + warp;
    const bool live = gf < n_frames_total;
    // centres and intensities of the CTA's 8 frames, one thread per (frame, sub-position) (see render_v1_kernel)
    if (d.draw) {
      for (int t = threadIdx.x; t < 8 * n; t += 256) {
        const int wi = t / n, p = t - wi * n;
        const long long gfi = g0 + wi;
        if (gfi >= n_frames_total) continue;
        const long long si = gfi / d.F;
        const int fi = (int)(gfi - si * d.F);
        WarpSmem wo = carve(smem + (size_t)wi * per_warp, n, P);
        const double* seg = traj + (size_t)si * d.T * 2 + (size_t)fi * n * 2;
        double mx, my;
        frame_mean(d, seg, mx, my);
        centre_one(d, seg, mx, my, p, wo);
        v1_intensity_one(d, wo, fi, n, p, global_seq(d, si));
      }
    }
    __syncthreads();
    if (live) {
      const long long s = gf / d.F;
      const int f = (int)(gf - s * d.F);
      const uint32_t seq = global_seq(d, s);
      if (d.draw) axis_table_v1<G>(d, lane, w);
      __syncwarp();
      v1_frame_pixels<G>(d, w, f, lane, seq, nz, [&](int pix, float v) { px[pix] = v; });
      for (int e = lane; e < E; e += 32) dsh[warp * E + e] = __ldg(demb + gf * E + e);
    } else {
      for (int pix = lane; pix < PP; pix += 32) px[pix] = 0.0f;
      for (int e = lane; e < E; e += 32) dsh[warp * E + e] = 0.0f;
    }
    __syncthreads();
    if (e0 < E) {
#pragma unroll 1
      for (int fr = 0; fr < 8; ++fr) {
        const float* pf = smem + (size_t)fr * per_warp + warp_smem_floats(n, P);
        float pv[PJ];
#pragma unroll
        for (int j = 0; j < PJ; ++j) { const int pix = lane + 32 * j; pv[j] = pix < PP ? pf[pix] : 0.0f; }
#pragma unroll
        for (int i = 0; i < EPW; ++i) {
          const float g = e0 + i < E ? dsh[fr * E + e0 + i] : 0.0f;
#pragma unroll
          for (int j = 0; j < PJ; ++j) acc[i][j] = fmaf(g, pv[j], acc[i][j]);
        }
        if (lane < EPW && e0 + lane < E) bacc += dsh[fr * E + e0 + lane];
      }
    }
    __syncthreads();
  }
  if (e0 < E) {
#pragma unroll
    for (int i = 0; i < EPW; ++i)
#pragma unroll
      for (int j = 0; j < PJ; ++j) {
        const int pix = lane + 32 * j;
        if (pix < PP && e0 + i < E) atomicAdd(dW + (size_t)(e0 + i) * PP + pix, acc[i][j]);
      }
    if (lane < EPW && e0 + lane < E && db != nullptr) atomicAdd(db + e0 + lane, bacc);
  }
}

__global__ void __launch_bounds__(256) render_psfnoise_kernel(const double* __restrict__ traj, long long n_frames_total,
                                                              RenderDev d, PsfNoiseDev v, float* __restrict__ out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  WarpSmem w = carve(smem + (size_t)warp * warp_smem_floats(d.n, d.P), d.n, d.P);
  const long long gf = (long long)blockIdx.x * warps + warp;
  if (gf >= n_frames_total) return;
  const long long s = gf / d.F;
  const int f = (int)(gf - s * d.F);
  const uint32_t seq = global_seq(d, s);
  const int P = d.P, n = d.n, PP = d.P * d.P;

  if (d.draw) frame_centres(d, traj + (size_t)s * d.T * 2, f, lane, w);
  float spot = 0.0f;
  if (d.draw) {  // one intensity per frame, shared by sub-positions and PSFs (trainSettingsPSFNoise.py:279,286)
    float z = 0.0f, z1;
    if (!d.mean_noise) {
      const uint4 r = philox4x32_10((uint32_t)f, 0u, seq, stream_word(MIVIT_STREAM_INTENSITY, 0), d.k0, d.k1);
      box_muller(r.x, r.y, z, z1);
    }
    spot = __fdiv_rn(__fadd_rn(d.imean, __fmul_rn(d.istd, z)), (float)n);
  }
  __syncwarp();
  for (int i = 0; i < v.n_psf; ++i) {
    if (d.draw) {
      for (int p = lane; p < n; p += 32)
        w.inten[p] = ((double)w.msq[p] * v.inv2s2_d[i] > 745.0) ? __int_as_float(0x7fc00000) : spot;
      __syncwarp();
      axis_table(d, v.inv2s2[i], lane, w);
      __syncwarp();
    }
    for (int pix = lane; pix < PP; pix += 32) {
      const int a = pix / P, b = pix - a * P;
      float prev = d.draw ? pixel_signal(d, w, a, b) : 0.0f;
      for (int j = 0; j < v.n_noise; ++j) {
        PixelStream st;
        float zb = 0.0f, z1;
        if (!d.mean_noise) {
          st.item = (uint32_t)(f * PP + pix); st.seq = seq;
          st.sw = stream_word(MIVIT_STREAM_PIXEL, (uint32_t)(1 + i * v.n_noise + j));
          st.k0 = d.k0; st.k1 = d.k1; st.q = 0;
          st.cur = philox4x32_10(st.item, 0u, seq, st.sw, d.k0, d.k1);
          box_muller(st.cur.x, st.cur.y, zb, z1);
        }
        const float bg = fminf(fmaxf(__fadd_rn(d.bg_mean, __fmul_rn(v.bg_std[j], zb)), 0.0f), v.bg_hi[j]);  // :303
        const float lam = __fmul_rn(__fadd_rn(prev, bg), d.poisson);
        const float k = d.mean_noise ? lam : poisson_draw(lam, st);
        const float val = __fdiv_rn(k, d.poisson);                                                          // :305
        if (j == 0) prev = val;  // out[psf,0,f] is overwritten by its noisy version (:303-305)
        out[((((size_t)s * v.n_psf + i) * v.n_noise + j) * d.F + f) * PP + pix] = val;
      }
    }
    __syncwarp();
  }
}

// trajectories_to_video_multiple_settings / trajectory_to_mult_settings (helpers/helpersGeneration.py:422-540): one intensity
// per FRAME shared by its sub-positions (:505,:512) and four outputs per frame -- noise free (:523), + clipped Gaussian
// background (:524-525), + proper Poisson(x*pn)/pn (:527), + skimage.filters.gaussian(sigma = 0.5) of the Poisson frame (:530),
// i.e. scipy's separable 5-tap filter with 'nearest' borders: axis 0 then axis 1, float64 accumulation, a float32 image after
// each axis.  The Poisson frame and the axis-0 result live in the warp's shared memory.
struct GaussTaps { double w[5]; };
__global__ void __launch_bounds__(256) render_multi_kernel(const double* __restrict__ traj, long long n_frames_total, RenderDev d,
                                                           GaussTaps taps, float* __restrict__ out_none, float* __restrict__ out_gauss,
                                                           float* __restrict__ out_pois, float* __restrict__ out_filt) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int P = d.P, n = d.n, PP = d.P * d.P;
  const int per_warp = warp_smem_floats(n, P) + 2 * PP;
  float* mine = smem + (size_t)warp * per_warp;
  WarpSmem w = carve(mine, n, P);
  float* fr = mine + warp_smem_floats(n, P);   // [PP] Poisson frame
  float* t0 = fr + PP;                          // [PP] after the axis-0 pass
  const long long gf = (long long)blockIdx.x * warps + warp;
  if (gf >= n_frames_total) return;
  const long long s = gf / d.F;
  const int f = (int)(gf - s * d.F);
  const uint32_t seq = global_seq(d, s);

  if (d.draw) frame_centres(d, traj + (size_t)s * d.T * 2, f, lane, w);
  float spot = 0.0f;
  if (d.draw) {
    float z = 0.0f, z1;
    if (!d.mean_noise) {
      const uint4 r = philox4x32_10((uint32_t)f, 0u, seq, stream_word(MIVIT_STREAM_INTENSITY, 0), d.k0, d.k1);
      box_muller(r.x, r.y, z, z1);
    }
    spot = __fdiv_rn(__fadd_rn(d.imean, __fmul_rn(d.istd, z)), (float)n);   // :505, :512
  }
  __syncwarp();
  if (d.draw) {
    for (int p = lane; p < n; p += 32) w.inten[p] = ((double)w.msq[p] * d.inv2s2_d > 745.0) ? __int_as_float(0x7fc00000) : spot;
    __syncwarp();
    axis_table(d, d.inv2s2, lane, w);
    __syncwarp();
  }
  const size_t base = (size_t)gf * PP;
  for (int pix = lane; pix < PP; pix += 32) {
    const int a = pix / P, b = pix - a * P;
    const float sig = d.draw ? pixel_signal(d, w, a, b) : 0.0f;
    PixelStream st;
    float zb = 0.0f, z1;
    if (!d.mean_noise) {
      st.item = (uint32_t)(f * PP + pix); st.seq = seq; st.sw = stream_word(MIVIT_STREAM_PIXEL, 0);
      st.k0 = d.k0; st.k1 = d.k1; st.q = 0;
      st.cur = philox4x32_10(st.item, 0u, seq, st.sw, d.k0, d.k1);
      box_muller(st.cur.x, st.cur.y, zb, z1);
    }
    const float bg = fminf(fmaxf(__fadd_rn(d.bg_mean, __fmul_rn(d.bg_std, zb)), 0.0f), d.bg_hi);   // :524-525
    const float g = __fadd_rn(sig, bg);
    const float lam = __fmul_rn(g, d.poisson);
    const float k = d.mean_noise ? lam : poisson_draw(lam, st);
    const float val = __fdiv_rn(k, d.poisson);                                                       // :527
    out_none[base + pix] = sig;
    out_gauss[base + pix] = g;
    out_pois[base + pix] = val;
    fr[pix] = val;
  }
  __syncwarp();
  for (int pix = lane; pix < PP; pix += 32) {   // axis 0 (rows), 'nearest' border
    const int a = pix / P, b = pix - a * P;
    double acc = 0.0;
#pragma unroll
    for (int i = -2; i <= 2; ++i) acc += taps.w[i + 2] * (double)fr[min(max(a + i, 0), P - 1) * P + b];
    t0[pix] = (float)acc;
  }
  __syncwarp();
  for (int pix = lane; pix < PP; pix += 32) {   // axis 1 (columns)
    const int a = pix / P, b = pix - a * P;
    double acc = 0.0;
#pragma unroll
    for (int j = -2; j <= 2; ++j) acc += taps.w[j + 2] * (double)t0[a * P + min(max(b + j, 0), P - 1)];
    out_filt[base + pix] = (float)acc;
  }
}

// Richardson-Lucy deconvolution with a total-variation step (helpers/helpersGeneration.py:542-587 tv_gradient,
// richardson_lucy_tv, richardson_lucy_tv_iter_list), one warp per P x P frame, everything in the warp's shared memory.
// Per iteration (reference order and dtypes): relative_blur = image / (conv(estimate, psf) + 1e-6) in float64,
// correction = conv(relative_blur, psf mirrored), estimate *= correction (float64 product stored as float32), the TV gradient
// and its step in float32, clip to [0, 1]; estimates after the listed (0-based) iterations are written out.  conv is scipy's
// fftconvolve(mode='same') as the direct float64 sum it equals (the reference's FFT of the float32 estimate runs in single
// precision, so the two agree to ~3e-4 after 11 iterations -- see oracle/render_oracle.py).
struct RlIters { int it[8]; int n; int last; };
__global__ void __launch_bounds__(256) rl_tv_kernel(const float* __restrict__ images, long long n_frames, int P, int K,
                                                    const double* __restrict__ psf, RlIters its, float tv_weight,
                                                    float* __restrict__ out) {
  extern __shared__ double sm_d[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int PP = P * P, KK = K * K, c = (K - 1) / 2;
  double* psf_s = sm_d;                                   // [KK] shared by the CTA
  double* rb = sm_d + KK + (size_t)warp * PP;             // [PP] relative blur (float64)
  float* fbase = reinterpret_cast<float*>(sm_d + KK + (size_t)warps * PP) + (size_t)warp * 4 * PP;
  float *img = fbase, *est = fbase + PP, *dxn = fbase + 2 * PP, *dyn = fbase + 3 * PP;
  for (int i = threadIdx.x; i < KK; i += blockDim.x) psf_s[i] = psf[i];
  __syncthreads();
  const long long fr = (long long)blockIdx.x * warps + warp;
  if (fr >= n_frames) return;
  for (int p = lane; p < PP; p += 32) {
    img[p] = fmaxf(images[fr * PP + p], 1e-6f);           // np.clip(image, 1e-6, None)
    est[p] = 0.5f;
  }
  __syncwarp();
  for (int iter = 0; iter <= its.last; ++iter) {
    for (int p = lane; p < PP; p += 32) {                 // image / (estimate (*) psf + 1e-6)
      const int i = p / P, j = p - i * P;
      double acc = 0.0;
      for (int u = 0; u < K; ++u) {
        const int ii = i + c - u;
        if (ii < 0 || ii >= P) continue;
        for (int v = 0; v < K; ++v) {
          const int jj = j + c - v;
          if (jj >= 0 && jj < P) acc += (double)est[ii * P + jj] * psf_s[u * K + v];
        }
      }
      rb[p] = (double)img[p] / (acc + 1e-6);
    }
    __syncwarp();
    float e_new[8];                                       // PP <= 256 pixels: up to 8 per lane
    int cnt = 0;
    for (int p = lane; p < PP; p += 32, ++cnt) {          // estimate *= relative_blur (*) psf[::-1, ::-1]
      const int i = p / P, j = p - i * P;
      double acc = 0.0;
      for (int u = 0; u < K; ++u) {
        const int ii = i + c - u;
        if (ii < 0 || ii >= P) continue;
        for (int v = 0; v < K; ++v) {
          const int jj = j + c - v;
          if (jj >= 0 && jj < P) acc += rb[ii * P + jj] * psf_s[(K - 1 - u) * K + (K - 1 - v)];
        }
      }
      e_new[cnt] = (float)((double)est[p] * acc);
    }
    cnt = 0;
    for (int p = lane; p < PP; p += 32, ++cnt) est[p] = e_new[cnt];
    __syncwarp();
    for (int p = lane; p < PP; p += 32) {                 // tv_gradient: normalised forward differences (float32)
      const int i = p / P, j = p - i * P;
      const float e = est[p];
      const float dx = (j + 1 < P) ? __fsub_rn(est[p + 1], e) : 0.0f;      // np.diff(..., append=last column) -> 0
      const float dy = (i + 1 < P) ? __fsub_rn(est[p + P], e) : 0.0f;
      const float mag = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), 1e-8f));
      dxn[p] = __fdiv_rn(dx, mag);
      dyn[p] = __fdiv_rn(dy, mag);
    }
    __syncwarp();
    cnt = 0;
    for (int p = lane; p < PP; p += 32, ++cnt) {
      const int i = p / P, j = p - i * P;
      float g = 0.0f;
      if (j + 1 < P) g = __fsub_rn(g, dxn[p]);            // grad[:, :-1] -= dx_norm[:, :-1]
      if (j >= 1) g = __fadd_rn(g, dxn[p - 1]);           // grad[:, 1:]  += dx_norm[:, :-1]
      if (i + 1 < P) g = __fsub_rn(g, dyn[p]);            // grad[:-1, :] -= dy_norm[:-1, :]
      if (i >= 1) g = __fadd_rn(g, dyn[p - P]);           // grad[1:, :]  += dy_norm[:-1, :]
      const float e = __fsub_rn(est[p], __fmul_rn(tv_weight, g));
      e_new[cnt] = fminf(fmaxf(e, 0.0f), 1.0f);
    }
    cnt = 0;
    for (int p = lane; p < PP; p += 32, ++cnt) est[p] = e_new[cnt];
    __syncwarp();
    for (int k = 0; k < its.n; ++k)
      if (its.it[k] == iter)
        for (int p = lane; p < PP; p += 32) out[((size_t)fr * its.n + k) * PP + p] = est[p];
  }
}

// One warp per sequence: D draw, steps, float64 prefix sum.
struct DGroups { float mean[16], var[16]; };

__global__ void __launch_bounds__(128) brownian_kernel(long long N, int T, DGroups g, int n_groups, double div,
                                                       uint32_t k0, uint32_t k1, unsigned long long seq_offset,
                                                       double* __restrict__ traj, float* __restrict__ D_out) {
  const int lane = threadIdx.x & 31;
  const long long s = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= N) return;
  const unsigned long long gid = seq_offset + (unsigned long long)s;
  const uint32_t seq = (uint32_t)gid;
  const int grp = (int)(gid % (unsigned long long)n_groups);
  float D = 0.0f;
  if (lane == 0) {
    const float mu = g.mean[grp], sd = sqrtf(g.var[grp]);
    for (uint32_t blk = 0; blk < 64 && !(D > 0.0f); ++blk) {
      const uint4 r = philox4x32_10(0u, blk, seq, stream_word(MIVIT_STREAM_D, 0), k0, k1);
      float z[4];
      box_muller(r.x, r.y, z[0], z[1]);
      box_muller(r.z, r.w, z[2], z[3]);
      for (int i = 0; i < 4 && !(D > 0.0f); ++i) D = __fadd_rn(mu, __fmul_rn(sd, z[i]));
    }
    if (!(D > 0.0f)) D = mu > 0.0f ? mu : 1.0f;
    D_out[s] = D;
  }
  D = __shfl_sync(0xffffffffu, D, 0);
  const float sig = sqrtf(2.0f * D);
  const int C = (T + 31) / 32;
  const int t0 = lane * C, t1 = min(T, t0 + C);
  double sx = 0.0, sy = 0.0;
  for (int t = max(t0, 1); t < t1; ++t) {  // position 0 is the origin; step t moves to position t
    const uint4 r = philox4x32_10((uint32_t)t, 0u, seq, stream_word(MIVIT_STREAM_TRAJ, 0), k0, k1);
    float zx, zy;
    box_muller(r.x, r.y, zx, zy);
    sx += (double)__fmul_rn(sig, zx);
    sy += (double)__fmul_rn(sig, zy);
  }
  double px = sx, py = sy;  // inclusive scan of lane totals
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double ux = __shfl_up_sync(0xffffffffu, px, o), uy = __shfl_up_sync(0xffffffffu, py, o);
    if (lane >= o) { px += ux; py += uy; }
  }
  double ax = px - sx, ay = py - sy;  // exclusive prefix
  double* dst = traj + (size_t)s * T * 2;
  for (int t = t0; t < t1; ++t) {
    if (t >= 1) {
      const uint4 r = philox4x32_10((uint32_t)t, 0u, seq, stream_word(MIVIT_STREAM_TRAJ, 0), k0, k1);
      float zx, zy;
      box_muller(r.x, r.y, zx, zy);
      ax += (double)__fmul_rn(sig, zx);
      ay += (double)__fmul_rn(sig, zy);
    }
    dst[2 * t] = ax / div;
    dst[2 * t + 1] = ay / div;
  }
}

// Alias table (Walker / Vose) of Poisson(lam) restricted to k in [k0, k0 + 256), float64 on the host; oracle/noise.py restates
// it operation by operation.  pmf by the recurrence p(k+1) = p(k) lam / (k+1) from the mode (no lgamma: identical in C and numpy).
void build_alias_table(double lam, AliasTable& at) {
  at.valid = 0;
  at.k0 = 0;
  for (int i = 0; i < kAliasEntries; ++i) at.e[i] = 0;
  if (!(lam > 0.0) || lam > (double)kAliasMaxLambda) return;
  const int K = kAliasEntries;
  const int mode = (int)floor(lam);
  const int k0 = mode > K / 2 ? mode - K / 2 : 0;
  double p[kAliasEntries];
  p[mode - k0] = 1.0;
  for (int k = mode; k + 1 < k0 + K; ++k) p[k + 1 - k0] = p[k - k0] * lam / (double)(k + 1);
  for (int k = mode; k - 1 >= k0; --k) p[k - 1 - k0] = p[k - k0] * (double)k / lam;
  double sum = 0.0;
  for (int i = 0; i < K; ++i) sum += p[i];
  double sc[kAliasEntries], prob[kAliasEntries];
  int alias[kAliasEntries], small[kAliasEntries], large[kAliasEntries], ns = 0, nl = 0;
  for (int i = 0; i < K; ++i) {
    sc[i] = p[i] / sum * (double)K;
    prob[i] = 1.0;
    alias[i] = i;
    if (sc[i] < 1.0) small[ns++] = i; else large[nl++] = i;
  }
  while (ns > 0 && nl > 0) {
    const int sm = small[--ns], lg = large[--nl];
    prob[sm] = sc[sm];
    alias[sm] = lg;
    sc[lg] = (sc[lg] + sc[sm]) - 1.0;
    if (sc[lg] < 1.0) small[ns++] = lg; else large[nl++] = lg;
  }
  for (int i = 0; i < K; ++i) {
    double t = floor(prob[i] * 16777216.0 + 0.5);
    if (t > 16777215.0) t = 16777215.0;
    if (t < 0.0) t = 0.0;
    at.e[i] = ((uint32_t)alias[i] << 24) | (uint32_t)t;
  }
  at.k0 = k0;
  at.valid = 1;
}

int fill_dev(const mivit_render_params* prm, int T, uint64_t seed, uint64_t seq_offset, RenderDev& d) {
  MIVIT_CHECK_ARG(prm != nullptr, "render params are NULL");
  MIVIT_CHECK_ARG(prm->P >= 1 && prm->P <= 128, "output_size %d out of range [1,128]", prm->P);
  MIVIT_CHECK_ARG(prm->U >= 1 && prm->U <= 64, "upsampling_factor %d out of range [1,64]", prm->U);
  MIVIT_CHECK_ARG(prm->n >= 1 && prm->n <= 1024, "nPosPerFrame %d out of range [1,1024]", prm->n);
  MIVIT_CHECK_ARG(T % prm->n == 0, "T is not divisble by posPerFrame");
  MIVIT_CHECK_ARG(prm->sigma_hr > 0.0, "PSF sigma must be positive");
  d.scale = prm->scale;
  d.ysign = prm->flip_y ? -1.0 : 1.0;
  d.P = prm->P; d.U = prm->U; d.n = prm->n; d.center = prm->center; d.draw = prm->draw;
  d.G = prm->P * prm->U;
  d.limit = (d.G - 1) / 2;
  d.step_d = d.G > 1 ? 2.0 * d.limit / (double)(d.G - 1) : 1.0;
  if (d.step_d == 0.0) d.step_d = 1.0;
  d.step = (float)d.step_d;
  d.inv_step_d = 1.0 / d.step_d;
  d.inv_n = 1.0 / (double)prm->n;
  d.inv2s2_d = 1.0 / (2.0 * prm->sigma_hr * prm->sigma_hr);
  d.inv2s2 = (float)d.inv2s2_d;
  d.T = T; d.F = T / prm->n;
  d.imean = prm->part_mean; d.istd = prm->part_std;
  d.bg_mean = prm->bg_mean; d.bg_std = prm->bg_std;
  d.bg_hi = (float)((double)prm->bg_mean + 3.0 * (double)prm->bg_std);
  d.poisson = prm->poisson;
  d.normalize = prm->normalize; d.norm_sub = prm->norm_sub; d.norm_div = prm->norm_div;
  d.inv_norm_div = prm->norm_div != 0.0f ? 1.0f / prm->norm_div : 0.0f;
  d.mean_noise = prm->mean_noise;
  d.k0 = (uint32_t)(seed & 0xFFFFFFFFull); d.k1 = (uint32_t)(seed >> 32);
  d.seq_offset = seq_offset;
  d.seq_off_dev = nullptr;
  return MIVIT_OK;
}

int pick_warps(int n, int P, int* warps, size_t* smem) {
  const size_t per_warp = (size_t)warp_smem_floats(n, P) * sizeof(float);
  int w = 8;
  while (w > 1 && per_warp * w > 48 * 1024) w >>= 1;
  MIVIT_CHECK_ARG(per_warp * w <= 200 * 1024, "nPosPerFrame*output_size too large for shared memory (%zu bytes per frame)", per_warp);
  *warps = w;
  *smem = per_warp * w;
  return MIVIT_OK;
}

}  // namespace

extern "C" int mivit_render_v1(const double* traj, int64_t N, int32_t T, const mivit_render_params* prm, uint64_t seed,
                               uint64_t seq_offset, float* out, int64_t out_seq_stride, void* stream) {
  return render_v1_launch(traj, N, T, prm, seed, seq_offset, nullptr, out, out_seq_stride, (cudaStream_t)stream);
}

int render_v1_launch(const double* traj, long long N, int T, const mivit_render_params* prm, uint64_t seed, uint64_t seq_offset,
                     const uint64_t* seq_off_dev, float* out, long long out_seq_stride, cudaStream_t stream) {
  RenderDev d;
  int rc = fill_dev(prm, T, seed, seq_offset, d);
  if (rc) return rc;
  d.seq_off_dev = reinterpret_cast<const unsigned long long*>(seq_off_dev);
  MIVIT_CHECK_ARG(N >= 0, "negative N");
  if (N == 0) return MIVIT_OK;
  MIVIT_CHECK_ARG(traj && out, "NULL device pointer");
  // V1 draws one intensity per sub-position: N(mean/n, std/n)  (helpersGeneration.py:300)
  d.imean = (float)((double)prm->part_mean / prm->n);
  d.istd = (float)((double)prm->part_std / prm->n);
  d.out_seq_stride = out_seq_stride;
  int warps; size_t smem;
  rc = pick_warps(d.n, d.P, &warps, &smem);
  if (rc) return rc;
  const long long frames = (long long)N * d.F;
  MivitProfScope prof("render_v1", (double)N * ((double)T * 16.0 + (double)d.F * d.P * d.P * 4.0), (cudaStream_t)stream);
  AliasTable at;
  build_alias_table(d.mean_noise || d.poisson == -1.0f ? 0.0 : (double)d.poisson, at);
  // compile-time geometry for the reference's experiment configurations (U = 5; P = 9 / 13 / 7 / 15; n = 10 or any)
#define MIVIT_V1(TP, TU, TN)                                                                                                    \
  do {                                                                                                                          \
    if (smem > 48 * 1024)                                                                                                       \
      MIVIT_CUDA_CHECK(cudaFuncSetAttribute(render_v1_kernel<TP, TU, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    render_v1_kernel<TP, TU, TN><<<mivit_ceil_div(frames, warps), warps * 32, smem, (cudaStream_t)stream>>>(traj, frames, d, at, out); \
  } while (0)
  if (d.U == 5 && d.P == 13 && d.n == 10) MIVIT_V1(13, 5, 10);
  else if (d.U == 5 && d.P == 9 && d.n == 10) MIVIT_V1(9, 5, 10);
  else if (d.U == 5 && d.P == 13) MIVIT_V1(13, 5, 0);
  else if (d.U == 5 && d.P == 9) MIVIT_V1(9, 5, 0);
  else if (d.U == 5 && d.P == 7) MIVIT_V1(7, 5, 0);
  else if (d.U == 5 && d.P == 15) MIVIT_V1(15, 5, 0);
  else MIVIT_V1(0, 0, 0);
#undef MIVIT_V1
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

// shared launcher of the fused render -> embedding kernel (C ABI below; vit_model.cu for the from-trajectories training step)
int render_embed_linear_launch(const double* traj, long long N, int T, const mivit_render_params* prm, uint64_t seed,
                               uint64_t seq_offset, const uint64_t* seq_off_dev, const float* W, int w_transposed, const float* bias,
                               int E, float* emb, float* frames_out, long long frames_seq_stride, cudaStream_t st) {
  RenderDev d;
  int rc = fill_dev(prm, T, seed, seq_offset, d);
  if (rc) return rc;
  d.seq_off_dev = reinterpret_cast<const unsigned long long*>(seq_off_dev);
  MIVIT_CHECK_ARG(N >= 0 && E >= 1 && E <= 1024, "bad N / embed_dim");
  if (N == 0) return MIVIT_OK;
  MIVIT_CHECK_ARG(traj && W && bias && emb, "NULL device pointer");
  d.imean = (float)((double)prm->part_mean / prm->n);
  d.istd = (float)((double)prm->part_std / prm->n);
  d.out_seq_stride = frames_seq_stride;
  const int PP = d.P * d.P;
  const size_t per_warp = (size_t)embed_warp_floats(d.n, d.P) * sizeof(float);
  const size_t wbytes = (((size_t)PP * E + 3) & ~(size_t)3) * sizeof(float);
  int warps = 8;
  while (warps > 1 && wbytes + per_warp * warps > 200 * 1024) warps >>= 1;
  const size_t smem = wbytes + per_warp * warps;
  MIVIT_CHECK_ARG(smem <= 220 * 1024, "embedding weight (%d x %d) does not fit shared memory next to the frame tables", E, PP);
  const long long frames = (long long)N * d.F;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int per_sm = (int)((220 * 1024) / smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  long long blocks = (long long)sms * per_sm;
  if (blocks > mivit_ceil_div(frames, warps)) blocks = mivit_ceil_div(frames, warps);
  AliasTable at;
  build_alias_table(d.mean_noise || d.poisson == -1.0f ? 0.0 : (double)d.poisson, at);
  MivitProfScope prof("render_embed_linear", (double)N * ((double)T * 16.0 + (double)d.F * E * 4.0), st);
#define MIVIT_EMB(TP, TU, TN)                                                                                                   \
  do {                                                                                                                          \
    MIVIT_CUDA_CHECK(cudaFuncSetAttribute(render_embed_linear_kernel<TP, TU, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    render_embed_linear_kernel<TP, TU, TN><<<(unsigned)blocks, warps * 32, smem, st>>>(traj, frames, d, at, W, w_transposed, bias, E, emb, \
                                                                                       frames_out);                           \
  } while (0)
  if (d.U == 5 && d.P == 13 && d.n == 10) MIVIT_EMB(13, 5, 10);
  else if (d.U == 5 && d.P == 9 && d.n == 10) MIVIT_EMB(9, 5, 10);
  else if (d.U == 5 && d.P == 7 && d.n == 10) MIVIT_EMB(7, 5, 10);
  else if (d.U == 5 && d.P == 15 && d.n == 10) MIVIT_EMB(15, 5, 10);
  else MIVIT_EMB(0, 0, 0);
#undef MIVIT_EMB
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

template <int EPW, int PJ, bool KG>
static int wgrad_launch_g(const double* traj, long long frames, const RenderDev& d, const AliasTable& at, const float* demb, int E,
                          float* dW, float* db, cudaStream_t st) {
  const size_t per_warp = (size_t)(warp_smem_floats(d.n, d.P) + d.P * d.P) * sizeof(float);
  const size_t smem = per_warp * 8 + (size_t)8 * E * sizeof(float);
  MIVIT_CHECK_ARG(smem <= 200 * 1024, "nPosPerFrame*output_size too large for shared memory (%zu bytes per frame)", per_warp);
  if (smem > 48 * 1024)
    MIVIT_CUDA_CHECK(cudaFuncSetAttribute(render_embed_wgrad_kernel<EPW, PJ, KG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long blocks = (long long)sms * 2;
  if (blocks > mivit_ceil_div(frames, 8)) blocks = mivit_ceil_div(frames, 8);
  render_embed_wgrad_kernel<EPW, PJ, KG><<<(unsigned)blocks, 256, smem, st>>>(traj, frames, d, at, demb, E, dW, db);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}
template <int EPW, int PJ>
static int wgrad_launch_t(const double* traj, long long frames, const RenderDev& d, const AliasTable& at, const float* demb, int E,
                          float* dW, float* db, cudaStream_t st) {
  constexpr int kP = PJ == 2 ? 7 : PJ == 3 ? 9 : PJ == 6 ? 13 : 15;
  if (d.U == 5 && d.n == 10 && d.P == kP) return wgrad_launch_g<EPW, PJ, true>(traj, frames, d, at, demb, E, dW, db, st);
  return wgrad_launch_g<EPW, PJ, false>(traj, frames, d, at, demb, E, dW, db, st);
}

// dW [E][P*P] and db [E] are ACCUMULATED into (the caller zeroes them)
int render_embed_wgrad_launch(const double* traj, long long N, int T, const mivit_render_params* prm, uint64_t seed,
                              uint64_t seq_offset, const uint64_t* seq_off_dev, const float* demb, int E, float* dW, float* db,
                              cudaStream_t st) {
  RenderDev d;
  int rc = fill_dev(prm, T, seed, seq_offset, d);
  if (rc) return rc;
  d.seq_off_dev = reinterpret_cast<const unsigned long long*>(seq_off_dev);
  MIVIT_CHECK_ARG(N >= 0 && E >= 1 && E <= 256, "bad N / embed_dim (the fused weight gradient covers embed_dim <= 256)");
  MIVIT_CHECK_ARG(d.P <= 16, "the fused weight gradient covers output_size <= 16");
  if (N == 0) return MIVIT_OK;
  MIVIT_CHECK_ARG(traj && demb && dW, "NULL device pointer");
  d.imean = (float)((double)prm->part_mean / prm->n);
  d.istd = (float)((double)prm->part_std / prm->n);
  d.out_seq_stride = 0;
  const long long frames = (long long)N * d.F;
  AliasTable at;
  build_alias_table(d.mean_noise || d.poisson == -1.0f ? 0.0 : (double)d.poisson, at);
  const int epw = (E + 7) / 8, pj = (d.P * d.P + 31) / 32;
  MivitProfScope prof("render_embed_wgrad", (double)N * ((double)T * 16.0 + (double)d.F * E * 4.0), st);
#define MIVIT_WG(EPW_, PJ_) return wgrad_launch_t<EPW_, PJ_>(traj, frames, d, at, demb, E, dW, db, st)
#define MIVIT_WG_E(PJ_)                                  \
  do {                                                   \
    if (epw <= 4) MIVIT_WG(4, PJ_);                      \
    if (epw <= 8) MIVIT_WG(8, PJ_);                      \
    if (epw <= 16) MIVIT_WG(16, PJ_);                    \
    MIVIT_WG(32, PJ_);                                   \
  } while (0)
  if (pj <= 2) MIVIT_WG_E(2);
  if (pj <= 3) MIVIT_WG_E(3);
  if (pj <= 6) MIVIT_WG_E(6);
  if (epw <= 16) { if (epw <= 4) MIVIT_WG(4, 8); if (epw <= 8) MIVIT_WG(8, 8); MIVIT_WG(16, 8); }
  // E > 128 with P > 13: two passes over feature halves would be needed; not a configuration of the reference
  mivit_set_error("fused embedding weight gradient: embed_dim %d with output_size %d is not supported", E, d.P);
  return MIVIT_ERR_INVALID;
#undef MIVIT_WG_E
#undef MIVIT_WG
}

extern "C" int mivit_render_embed_linear(const double* traj, int64_t N, int32_t T, const mivit_render_params* prm, uint64_t seed,
                                         uint64_t seq_offset, const float* Wt, const float* bias, int32_t E, float* emb,
                                         float* frames_out, int64_t frames_seq_stride, void* stream) {
  return render_embed_linear_launch(traj, N, T, prm, seed, seq_offset, nullptr, Wt, 1, bias, E, emb, frames_out, frames_seq_stride,
                                    (cudaStream_t)stream);
}

extern "C" int mivit_render_embed_linear_wgrad(const double* traj, int64_t N, int32_t T, const mivit_render_params* prm,
                                               uint64_t seed, uint64_t seq_offset, const float* demb, int32_t E, float* dW,
                                               float* db, void* stream) {
  return render_embed_wgrad_launch(traj, N, T, prm, seed, seq_offset, nullptr, demb, E, dW, db, (cudaStream_t)stream);
}

extern "C" int mivit_poisson_alias_table(double lam, uint32_t* entries_host, int32_t* k0_host) {
  MIVIT_CHECK_ARG(entries_host && k0_host, "NULL pointer");
  AliasTable at;
  build_alias_table(lam, at);
  for (int i = 0; i < kAliasEntries; ++i) entries_host[i] = at.e[i];
  *k0_host = at.k0;
  return at.valid ? MIVIT_OK : MIVIT_ERR_INVALID;
}

extern "C" int mivit_render_multi(const double* traj, int64_t N, int32_t T, const mivit_render_params* prm, uint64_t seed,
                                  uint64_t seq_offset, float* out_no_noise, float* out_gauss, float* out_poisson,
                                  float* out_filter, void* stream) {
  RenderDev d;
  int rc = fill_dev(prm, T, seed, seq_offset, d);
  if (rc) return rc;
  MIVIT_CHECK_ARG(N >= 0, "negative N");
  MIVIT_CHECK_ARG(prm->poisson > 0.0f, "the multiple-settings renderer needs poisson_noise > 0");
  if (N == 0) return MIVIT_OK;
  MIVIT_CHECK_ARG(traj && out_no_noise && out_gauss && out_poisson && out_filter, "NULL device pointer");
  d.imean = prm->part_mean;   // one draw per FRAME, divided by n in the kernel (:505, :512)
  d.istd = prm->part_std;
  d.normalize = 0;
  GaussTaps taps;             // scipy.ndimage.gaussian_filter(sigma = 0.5, truncate = 4.0): radius 2, normalised float64 weights
  double sum = 0.0;
  for (int k = -2; k <= 2; ++k) { taps.w[k + 2] = exp(-0.5 / (0.5 * 0.5) * (double)(k * k)); sum += taps.w[k + 2]; }
  for (int k = 0; k < 5; ++k) taps.w[k] /= sum;
  const size_t per_warp = (size_t)(warp_smem_floats(d.n, d.P) + 2 * d.P * d.P) * sizeof(float);
  int warps = 8;
  while (warps > 1 && per_warp * warps > 96 * 1024) warps >>= 1;
  MIVIT_CHECK_ARG(per_warp * warps <= 200 * 1024, "nPosPerFrame*output_size too large for shared memory (%zu bytes per frame)", per_warp);
  const size_t smem = per_warp * warps;
  if (smem > 48 * 1024)
    MIVIT_CUDA_CHECK(cudaFuncSetAttribute(render_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long frames = (long long)N * d.F;
  render_multi_kernel<<<mivit_ceil_div(frames, warps), warps * 32, smem, (cudaStream_t)stream>>>(traj, frames, d, taps, out_no_noise,
                                                                                                  out_gauss, out_poisson, out_filter);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

extern "C" int mivit_rl_tv(const float* images, int64_t n_images, int32_t P, const double* psf_dev, int32_t K,
                           const int32_t* iterations_host, int32_t n_iterations, float tv_weight, float* out, void* stream) {
  MIVIT_CHECK_ARG(n_images >= 0 && P >= 1 && P <= 16, "Richardson-Lucy/TV frames must be at most 16 x 16");
  MIVIT_CHECK_ARG(K >= 1 && K % 2 == 1 && K <= 31, "PSF size must be odd and at most 31");
  MIVIT_CHECK_ARG(iterations_host && n_iterations >= 1 && n_iterations <= 8, "1 to 8 iteration counts");
  if (n_images == 0) return MIVIT_OK;
  MIVIT_CHECK_ARG(images && psf_dev && out, "NULL device pointer");
  RlIters its;
  its.n = n_iterations;
  its.last = 0;
  for (int k = 0; k < n_iterations; ++k) {
    MIVIT_CHECK_ARG(iterations_host[k] >= 0 && iterations_host[k] <= 10000, "bad iteration index");
    its.it[k] = iterations_host[k];
  }
  its.last = iterations_host[n_iterations - 1];    // the reference runs iterations_list[-1] + 1 iterations (:575)
  const int warps = 8, PP = P * P;
  const size_t smem = (size_t)(K * K + warps * PP) * sizeof(double) + (size_t)warps * 4 * PP * sizeof(float);
  if (smem > 48 * 1024) MIVIT_CUDA_CHECK(cudaFuncSetAttribute(rl_tv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rl_tv_kernel<<<mivit_ceil_div(n_images, warps), warps * 32, smem, (cudaStream_t)stream>>>(images, n_images, P, K, psf_dev, its,
                                                                                              tv_weight, out);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

extern "C" int mivit_render_psfnoise(const double* traj, int64_t N, int32_t T, const mivit_render_params* prm,
                                     const float* psf_div_host, int32_t n_psf, const float* noise_frac_host,
                                     int32_t n_noise, float part_mean_global, uint64_t seed, uint64_t seq_offset,
                                     float* out, void* stream) {
  RenderDev d;
  int rc = fill_dev(prm, T, seed, seq_offset, d);
  if (rc) return rc;
  MIVIT_CHECK_ARG(n_psf >= 1 && n_noise >= 1 && psf_div_host && noise_frac_host, "No settings given");
  MIVIT_CHECK_ARG(n_psf <= kMaxVariants && n_noise <= kMaxVariants, "at most %d PSF and %d noise settings", kMaxVariants, kMaxVariants);
  MIVIT_CHECK_ARG(prm->poisson > 0.0f, "PSFNoise renderer needs poisson_noise > 0");
  if (N == 0) return MIVIT_OK;
  MIVIT_CHECK_ARG(traj && out, "NULL device pointer");
  d.ysign = 1.0;  // no y flip in this variant
  d.normalize = 0;
  PsfNoiseDev v;
  v.n_psf = n_psf; v.n_noise = n_noise;
  for (int i = 0; i < n_psf; ++i) {
    MIVIT_CHECK_ARG(psf_div_host[i] > 0.0f, "PSF setting must be positive");
    const double sg = prm->sigma_hr / (double)psf_div_host[i];  // :290
    v.inv2s2_d[i] = 1.0 / (2.0 * sg * sg);
    v.inv2s2[i] = (float)v.inv2s2_d[i];
  }
  for (int j = 0; j < n_noise; ++j) {
    const double bs = (double)part_mean_global * (double)noise_frac_host[j];  // :302
    v.bg_std[j] = (float)bs;
    v.bg_hi[j] = (float)((double)prm->bg_mean + 3.0 * (double)v.bg_std[j]);
  }
  int warps; size_t smem;
  rc = pick_warps(d.n, d.P, &warps, &smem);
  if (rc) return rc;
  if (smem > 48 * 1024)
    MIVIT_CUDA_CHECK(cudaFuncSetAttribute(render_psfnoise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long frames = (long long)N * d.F;
  MivitProfScope prof("render_psfnoise", (double)N * ((double)T * 16.0 + (double)n_psf * n_noise * d.F * d.P * d.P * 4.0), (cudaStream_t)stream);
  render_psfnoise_kernel<<<mivit_ceil_div(frames, warps), warps * 32, smem, (cudaStream_t)stream>>>(traj, frames, d, v, out);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

extern "C" int mivit_brownian(int64_t N, int32_t T, const float* group_mean_host, const float* group_var_host,
                              int32_t n_groups, double div, uint64_t seed, uint64_t seq_offset, double* traj,
                              float* D_out, void* stream) {
  MIVIT_CHECK_ARG(N >= 0 && T >= 1, "bad N/T");
  MIVIT_CHECK_ARG(n_groups >= 1 && n_groups <= 16 && group_mean_host && group_var_host, "need 1..16 D groups");
  MIVIT_CHECK_ARG(div != 0.0, "div must be non-zero");
  if (N == 0) return MIVIT_OK;
  MIVIT_CHECK_ARG(traj && D_out, "NULL device pointer");
  DGroups g;
  for (int i = 0; i < 16; ++i) { g.mean[i] = i < n_groups ? group_mean_host[i] : 0.f; g.var[i] = i < n_groups ? group_var_host[i] : 0.f; }
  brownian_kernel<<<mivit_ceil_div(N, 4), 128, 0, (cudaStream_t)stream>>>(
      N, T, g, n_groups, div, (uint32_t)(seed & 0xFFFFFFFFull), (uint32_t)(seed >> 32),
      seq_offset, traj, D_out);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}
