// Trajectory-feature producer of the ImagesFeatures experiment: the 25 values the ViT receives as `features`
// (reference helpers/helpersFeatures.py:448-520 compute_diffusion_features, order of :7-34; called per trajectory
// from helpers/helpersGeneration.py:674-719 create_video_and_feature_pairs on the frame-averaged positions
// :48-74).  The reference spends 12.6 ms per 30-point trajectory in Python loops and scipy; here one warp owns one
// trajectory, everything is float64 like the reference's numpy arithmetic.
//
// The only non-closed-form piece is the bounded power-law fit MSD(t) = 4 D t^alpha + c (:135-191, scipy
// curve_fit/trf from p0 = [msd[0]/(4 dt), 1, 0.001], bounds D >= 1e-5, 1e-5 <= alpha <= 10, c >= 0).  It is solved
// by variable projection: for a given alpha the model is linear in (D, c) -- a 2-variable bound-constrained least
// squares with a closed-form solution -- so the fit is a 1-D minimisation over alpha, bracketed downhill from the
// reference's starting point alpha = 1 and finished by golden-section search.  This is the constrained minimiser
// trf converges to; agreement is limited by trf's own stopping tolerance (ftol = xtol = 1e-8), see the tests.
#include <math.h>

#include "common.cuh"
#include "../../include/mivit.h"

namespace {

constexpr int kMaxLagsPerLane = 8;   // M <= 256 lags (trajectories of up to 512 points)

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double wmin(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double sgn(double v) { return (double)((v > 0.0) - (v < 0.0)); }

struct Fit {
  double sse, D, c;
};

// SSE of the best (D, c) for this alpha; every lane returns the same values.  tl[k] = log(t) of this lane's lags.
__device__ Fit profile_fit(double alpha, int M, int lane, const double* __restrict__ msd, const double (&tl)[kMaxLagsPerLane],
                           double Sy) {
  double phi[kMaxLagsPerLane];
  double spp = 0.0, sp = 0.0, spy = 0.0;
#pragma unroll
  for (int k = 0; k < kMaxLagsPerLane; ++k) {
    const int i = lane + 32 * k;
    phi[k] = 0.0;
    if (i < M) {
      phi[k] = 4.0 * exp(alpha * tl[k]);
      spp += phi[k] * phi[k];
      sp += phi[k];
      spy += phi[k] * msd[i];
    }
  }
  spp = wsum(spp); sp = wsum(sp); spy = wsum(spy);
  const double n = (double)M, Dmin = 1e-5;
  double D, c;
  const double det = n * spp - sp * sp;
  bool interior = false;
  if (det > 0.0) {
    D = (n * spy - sp * Sy) / det;
    c = (Sy - D * sp) / n;
    interior = D >= Dmin && c >= 0.0;
  }
  double Dc[2][2];
  int ncand = 1;
  if (interior) {
    Dc[0][0] = D; Dc[0][1] = c;
  } else {   // optimum on an edge of the box: c = 0 or D = Dmin
    Dc[0][0] = fmax(Dmin, spp > 0.0 ? spy / spp : Dmin); Dc[0][1] = 0.0;
    Dc[1][0] = Dmin; Dc[1][1] = fmax(0.0, (Sy - Dmin * sp) / n);
    ncand = 2;
  }
  Fit best;
  best.sse = INFINITY; best.D = Dc[0][0]; best.c = Dc[0][1];
  for (int q = 0; q < ncand; ++q) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kMaxLagsPerLane; ++k) {
      const int i = lane + 32 * k;
      if (i < M) {
        const double r = Dc[q][0] * phi[k] + Dc[q][1] - msd[i];
        s += r * r;
      }
    }
    s = wsum(s);
    if (s < best.sse) { best.sse = s; best.D = Dc[q][0]; best.c = Dc[q][1]; }
  }
  return best;
}

// one warp per trajectory; shared memory per warp: x[L] y[L] sl[L] msd[L] r4[L] (double) + order[L] hull[2L+2] (int)
__global__ void __launch_bounds__(128) features_kernel(const double* __restrict__ traj, long long N, int L, double dt,
                                                       double* __restrict__ out) {
  extern __shared__ double fsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long s = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (s >= N) return;
  const size_t per_warp = (size_t)5 * L + (size_t)(3 * L + 4 + 1) / 2;   // doubles
  double* x = fsm + warp * per_warp;
  double* y = x + L;
  double* sl = y + L;
  double* msd = sl + L;
  double* r4 = msd + L;
  int* order = reinterpret_cast<int*>(r4 + L);
  int* hull = order + L;
  double* o = out + s * 25;
  const double NaN = __longlong_as_double(0x7ff8000000000000ll);
  if (L < 3) {
    if (lane < 25) o[lane] = NaN;
    return;
  }
  for (int i = lane; i < L; i += 32) {
    x[i] = traj[(s * L + i) * 2];
    y[i] = traj[(s * L + i) * 2 + 1];
  }
  __syncwarp();
  // ---- step lengths (:441-446) -----------------------------------------------------------------
  const int ns = L - 1;
  double tot = 0.0, bottom = 0.0, smin = INFINITY, smax = -INFINITY, nsmall = 0.0, nlarge = 0.0;
  for (int i = lane; i < ns; i += 32) {
    const double dx = x[i + 1] - x[i], dy = y[i + 1] - y[i];
    const double q = dx * dx + dy * dy, v = sqrt(q);
    sl[i] = v;
    tot += v; bottom += q;
    smin = fmin(smin, v); smax = fmax(smax, v);
    nsmall += v < 0.1; nlarge += v > 0.4;
  }
  tot = wsum(tot); bottom = wsum(bottom); smin = wmin(smin); smax = wmax(smax); nsmall = wsum(nsmall); nlarge = wsum(nlarge);
  const double mean_sl = tot / ns;
  double ss = 0.0;
  for (int i = lane; i < ns; i += 32) { const double d = sl[i] - mean_sl; ss += d * d; }
  ss = wsum(ss);
  // ---- dot products of consecutive steps (:404-438) ----------------------------------------------
  const int nd = L - 2;
  double dsum = 0.0, dpos = 0.0, dsame = 0.0;
  for (int i = lane; i < nd; i += 32) {
    const double ax = x[i + 1] - x[i], ay = y[i + 1] - y[i], bx = x[i + 2] - x[i + 1], by = y[i + 2] - y[i + 1];
    const double d0 = ax * bx + ay * by;
    dsum += d0; dpos += d0 > 0.0;
    if (i + 1 < nd) {
      const double cx = x[i + 3] - x[i + 2], cy = y[i + 3] - y[i + 2];
      dsame += sgn(bx * cx + by * cy) == sgn(d0);
    }
  }
  dsum = wsum(dsum); dpos = wsum(dpos); dsame = wsum(dsame);
  // ---- MSD and fourth moments per lag (:102-132, :250-284) -------------------------------------------
  const int Nl = L > 20 ? (int)(L * 0.5) : L;
  const int M = Nl - 1;
  for (int lag = 1 + lane; lag <= M; lag += 32) {
    double s2 = 0.0, s4 = 0.0;
    for (int j = 0; j + lag < L; ++j) {
      const double dx = x[j + lag] - x[j], dy = y[j + lag] - y[j];
      const double a = dx * dx, b = dy * dy;
      s2 += a + b;
      s4 += a * a + b * b;
    }
    msd[lag - 1] = s2 / (double)(L - lag);
    r4[lag - 1] = s4 / (double)(L - lag);
  }
  __syncwarp();
  // ---- largest squared pair distance (:62-97) -----------------------------------------------------
  double maxd = 0.0;
  for (int i = lane; i < L; i += 32)
    for (int j = i + 1; j < L; ++j) {
      const double dx = x[j] - x[i], dy = y[j] - y[i];
      maxd = fmax(maxd, dx * dx + dy * dy);
    }
  maxd = wmax(maxd);
  // ---- reductions over the MSD curve ------------------------------------------------------------
  double Sy = 0.0, gsum = 0.0, gcnt = 0.0, rsum = 0.0;
  for (int i = lane; i < M; i += 32) {
    Sy += msd[i];
    if (i + 1 < L && msd[i] > 0.0) { gsum += r4[i] / (2.0 * msd[i] * msd[i]); gcnt += 1.0; }
    if (i + 1 < M) rsum += msd[i] / msd[i + 1] - (double)(i + 1) / (double)(i + 2);
  }
  Sy = wsum(Sy); gsum = wsum(gsum); gcnt = wsum(gcnt); rsum = wsum(rsum);
  const double ybar = Sy / M;
  double sst = 0.0;
  for (int i = lane; i < M; i += 32) { const double d = msd[i] - ybar; sst += d * d; }
  sst = wsum(sst);
  // ---- power-law fit (:135-191) ---------------------------------------------------------------------
  double alpha = 0.0, Dfit = 0.0, r2 = 0.0;
  if (M <= 32 * kMaxLagsPerLane) {
    double tl[kMaxLagsPerLane];
#pragma unroll
    for (int k = 0; k < kMaxLagsPerLane; ++k) tl[k] = log((double)(lane + 32 * k + 1) * dt);
    const double lo = 1e-5, hi = 10.0;
    double b = 1.0, h = 0.0625;
    Fit fb = profile_fit(b, M, lane, msd, tl, Sy);
    double a = fmax(lo, b - h), c = fmin(hi, b + h);
    Fit fa = profile_fit(a, M, lane, msd, tl, Sy), fc = profile_fit(c, M, lane, msd, tl, Sy);
    for (int it = 0; it < 64 && !(fb.sse <= fa.sse && fb.sse <= fc.sse); ++it) {   // walk downhill, doubling the step
      h *= 2.0;
      if (fa.sse < fc.sse) {
        if (a <= lo) { b = a; fb = fa; break; }
        c = b; fc = fb; b = a; fb = fa; a = fmax(lo, b - h); fa = profile_fit(a, M, lane, msd, tl, Sy);
      } else {
        if (c >= hi) { b = c; fb = fc; break; }
        a = b; fa = fb; b = c; fb = fc; c = fmin(hi, b + h); fc = profile_fit(c, M, lane, msd, tl, Sy);
      }
    }
    const double gr = 0.3819660112501051;
    for (int it = 0; it < 100 && (c - a) > 1e-11 * fmax(1.0, fabs(b)); ++it) {      // golden-section inside [a, c]
      const bool right = (c - b) > (b - a);
      const double u = right ? b + gr * (c - b) : b - gr * (b - a);
      const Fit fu = profile_fit(u, M, lane, msd, tl, Sy);
      if (fu.sse < fb.sse) {
        if (right) { a = b; fa = fb; } else { c = b; fc = fb; }
        b = u; fb = fu;
      } else {
        if (right) { c = u; fc = fu; } else { a = u; fa = fu; }
      }
    }
    alpha = b; Dfit = fb.D;
    r2 = 1.0 - fb.sse / sst;
  }
  // ---- kurtosis of the projection on the dominant direction (:287-324) ---------------------------------
  double mx = 0.0, my = 0.0;
  for (int i = lane; i < L; i += 32) { mx += x[i]; my += y[i]; }
  mx = wsum(mx) / L; my = wsum(my) / L;
  double cxx = 0.0, cyy = 0.0, cxy = 0.0;
  for (int i = lane; i < L; i += 32) { const double a = x[i] - mx, b = y[i] - my; cxx += a * a; cyy += b * b; cxy += a * b; }
  cxx = wsum(cxx) / (L - 1); cyy = wsum(cyy) / (L - 1); cxy = wsum(cxy) / (L - 1);
  const double hd = 0.5 * (cxx - cyy), lam = 0.5 * (cxx + cyy) + sqrt(hd * hd + cxy * cxy);
  double vx = cxy, vy = lam - cxx;
  if (vx == 0.0 && vy == 0.0) { vx = cxx >= cyy ? 1.0 : 0.0; vy = 1.0 - vx; }
  const double vn = sqrt(vx * vx + vy * vy);
  vx /= vn; vy /= vn;
  double pm = 0.0;
  for (int i = lane; i < L; i += 32) pm += vx * x[i] + vy * y[i];
  pm = wsum(pm) / L;
  double m2 = 0.0, m4 = 0.0;
  for (int i = lane; i < L; i += 32) { const double d = vx * x[i] + vy * y[i] - pm, q = d * d; m2 += q; m4 += q * q; }
  m2 = wsum(m2) / L; m4 = wsum(m4) / L;
  // ---- convex hull area (:381-402): rank sort + Andrew's monotone chain + shoelace -----------------------
  for (int i = lane; i < L; i += 32) {
    int r = 0;
    for (int j = 0; j < L; ++j) r += (x[j] < x[i]) || (x[j] == x[i] && (y[j] < y[i] || (y[j] == y[i] && j < i)));
    order[r] = i;
  }
  __syncwarp();
  double area = 0.0;
  if (lane == 0) {
    int k = 0;
    for (int pass = 0; pass < 2; ++pass) {
      const int base = k;
      for (int q = 0; q < L; ++q) {
        const int p = order[pass == 0 ? q : L - 1 - q];
        while (k - base >= 2) {
          const int a = hull[k - 2], b = hull[k - 1];
          if ((x[b] - x[a]) * (y[p] - y[a]) - (y[b] - y[a]) * (x[p] - x[a]) > 0.0) break;
          --k;
        }
        hull[k++] = p;
      }
      --k;   // the last point of a chain is the first of the next
    }
    if (k >= 3) {
      double acc = 0.0;
      for (int q = 0; q < k; ++q) {
        const int a = hull[q], b = hull[(q + 1) % k];
        acc += x[a] * y[b] - x[b] * y[a];
      }
      area = 0.5 * fabs(acc);
    }
  }
  // ---- assemble (:480-518) --------------------------------------------------------------------------
  if (lane == 0) {
    const double top = (x[L - 1] - x[0]) * (x[L - 1] - x[0]) + (y[L - 1] - y[0]) * (y[L - 1] - y[0]);
    double eff = 0.0, eff_log = -INFINITY;
    if (bottom != 0.0) { eff = top / ((double)(L - 1) * bottom); eff_log = log(eff); }
    const double lL = log((double)L);
    const double r0 = sqrt(maxd) / 2.0;
    o[0] = alpha;
    o[1] = Dfit;
    o[2] = r2;
    o[3] = eff_log;
    o[4] = eff;
    o[5] = tot == 0.0 ? 1.0 : lL / (lL + log(sqrt(maxd) / tot));
    o[6] = gcnt > 0.0 ? gsum / gcnt : NaN;
    o[7] = m4 / (m2 * m2);
    o[8] = M >= 2 ? rsum / (double)(M - 1) : NaN;
    o[9] = (r0 == 0.0 || Dfit == 0.0) ? 0.0 : 1.0 - exp(0.2045 - 0.25117 * (Dfit * L) / (r0 * r0));
    o[10] = (double)L;
    o[11] = mean_sl;
    o[12] = ybar;
    o[13] = dsum / nd;
    o[14] = nd > 1 ? dsame / (double)(nd - 1) : NaN;
    o[15] = dpos / nd;
    o[16] = tot;
    o[17] = smin;
    o[18] = smax;
    o[19] = smax - smin;
    o[20] = tot / L;
    o[21] = (mean_sl > 0.0 && ns > 1) ? sqrt(ss / (double)(ns - 1)) / mean_sl : NaN;
    o[22] = nsmall / ns;
    o[23] = nlarge / ns;
    o[24] = area;
  }
}

// out[s, f, :] = mean over the n sub-positions of frame f (helpersGeneration.py:48-74)
__global__ void average_frames_kernel(const double* __restrict__ traj, long long N, int T, int n, double* __restrict__ out) {
  const int F = T / n;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * F * 2) return;
  const int d = (int)(idx & 1);
  const long long sf = idx >> 1;
  const long long s = sf / F;
  const int f = (int)(sf - s * F);
  const double* p = traj + (s * T + (long long)f * n) * 2 + d;
  double acc = 0.0;
  for (int k = 0; k < n; ++k) acc += p[2 * k];
  out[idx] = acc / (double)n;
}

}  // namespace

extern "C" int mivit_average_frames(const double* traj, int64_t N, int32_t T, int32_t n, double* out, void* stream) {
  MIVIT_CHECK_ARG(N >= 0 && T >= 1 && n >= 1, "bad N / T / nPosPerFrame");
  if (N == 0 || T / n == 0) return MIVIT_OK;
  MIVIT_CHECK_ARG(traj && out, "NULL device pointer");
  const long long total = (long long)N * (T / n) * 2;
  average_frames_kernel<<<mivit_ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(traj, N, T, n, out);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}

extern "C" int mivit_diffusion_features(const double* traj, int64_t N, int32_t L, double dt, double* out, void* stream) {
  MIVIT_CHECK_ARG(N >= 0 && L >= 0, "bad N / trajectory length");
  MIVIT_CHECK_ARG(L <= 512, "trajectories of more than 512 points are not supported (%d)", L);
  MIVIT_CHECK_ARG(dt > 0.0, "dt must be positive");
  if (N == 0) return MIVIT_OK;
  MIVIT_CHECK_ARG(traj && out, "NULL device pointer");
  const int Ls = L < 3 ? 3 : L;
  const size_t per_warp = ((size_t)5 * Ls + (size_t)(3 * Ls + 4 + 1) / 2) * sizeof(double);
  int warps = 4;
  while (warps > 1 && per_warp * warps > 96 * 1024) warps >>= 1;
  const size_t smem = per_warp * warps;
  if (smem > 48 * 1024)
    MIVIT_CUDA_CHECK(cudaFuncSetAttribute(features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MivitProfScope prof("diffusion_features", (double)N * L * 16.0, (cudaStream_t)stream);
  features_kernel<<<mivit_ceil_div(N, warps), warps * 32, smem, (cudaStream_t)stream>>>(traj, N, L, dt, out);
  mivit_count_launch();
  MIVIT_LAUNCH_CHECK();
  return MIVIT_OK;
}
