// Counter-based RNG for the renderer and the trajectory source.
// Philox4x32-10 (Salmon et al., SC'11).  Stream layout (mirrored bit-for-bit by
// oracle/philox.py and oracle/noise.py so noisy frames can be checked draw-for-draw):
//     key     = (seed_lo, seed_hi)
//     counter = (item, block, seq_id, stream | variant << 8)
// The reference draws from the global np.random state
// (helpers/helpersGeneration.py:300,312,317), which has no parallel equivalent; a
// counter stream keyed by the GLOBAL sequence id makes the generated data set
// independent of how sequences are sharded over GPUs.
#pragma once
#include <stdint.h>

#define MIVIT_STREAM_TRAJ 0u
#define MIVIT_STREAM_D 1u
#define MIVIT_STREAM_INTENSITY 2u
#define MIVIT_STREAM_PIXEL 3u
#define MIVIT_STREAM_LOCERR 4u

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint32_t stream_word(uint32_t stream, uint32_t variant) {
  return (stream & 0xFFu) | (variant << 8);
}

// uint32 -> (0,1] float32 (curand's _curand_uniform)
__device__ __forceinline__ float u01(uint32_t x) {
  return __fmaf_rn((float)x, 2.3283064365386963e-10f, 2.3283064365386963e-10f / 2.0f);
}

// two words -> two standard normals
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& z0, float& z1) {
  const float u = u01(xa);
  const float v = u01(xb) * 6.2831853071795860f;
  const float r = sqrtf(-2.0f * logf(u));
  float s, c;
  sincosf(v, &s, &c);
  z0 = r * s;
  z1 = r * c;
}

// ---- V1 renderer noise layout ("pair" layout, round 2) ----------------------------------------------------------------
// The V1 renderer (helpers/helpersGeneration.py:312-317) needs ONE standard normal (background) and ONE Poisson(pn) draw with
// a launch-constant pn per pixel.  One Philox4x32-10 block therefore serves a PAIR of horizontally adjacent pixels:
//     counter = (frame * pairs_per_frame + row * ceil(P/2) + pair_in_row, 0, seq_id, PIXEL stream)
//     words x,y -> Box-Muller (angle shifted to (-pi, pi]) -> z for the left / right pixel
//     word  z   -> Poisson draw of the left pixel, word w -> of the right pixel, through an alias table (Walker / Vose):
//                  j = word >> 24, accept k0 + j if (word & 0xFFFFFF) < thresh[j], else k0 + alias[j].
// The table (256 entries, built on the host in float64 by mivit_poisson_alias_table, restated in oracle/noise.py) covers
// k in [k0, k0 + 256) and is exact up to the 2^-24 threshold quantisation and the mass outside +-6.5 sigma (< 1e-10);
// it applies for pn <= 380.  Larger pn fall back to the PTRS sampler on the pixel's own uniform stream (blocks >= 1).
constexpr int kAliasEntries = 256;
constexpr float kAliasMaxLambda = 380.0f;
struct AliasTable {
  uint32_t e[kAliasEntries];   // alias index << 24 | threshold (24 bits)
  int k0;
  int valid;
};

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// two words -> two standard normals, MUFU only: r = sqrt(-2 ln u), angle in (-pi, pi]
__device__ __forceinline__ void box_muller_fast(uint32_t xa, uint32_t xb, float& z0, float& z1) {
  const float u = u01(xa);
  const float th = (u01(xb) - 0.5f) * 6.2831853071795860f;
  const float r = sqrtf(-1.3862943611198906f * fast_lg2(u));   // -2 ln 2 * log2(u)
  z0 = r * __sinf(th);
  z1 = r * __cosf(th);
}

__device__ __forceinline__ float alias_draw(const uint32_t* __restrict__ tab, int k0, uint32_t word) {
  const uint32_t ent = tab[word >> 24];
  const uint32_t j = (word & 0xFFFFFFu) < (ent & 0xFFFFFFu) ? (word >> 24) : (ent >> 24);
  return (float)(k0 + (int)j);
}

__constant__ float kLogFact[16] = {
    0.0f, 0.0f, 0.693147180559945f, 1.791759469228055f, 3.178053830347946f, 4.787491742782046f,
    6.579251212010101f, 8.525161361065415f, 10.604602902745251f, 12.801827480081469f,
    15.104412573075516f, 17.502307845873887f, 19.987214495661885f, 22.552163853123425f,
    25.191221182738680f, 27.899271383840890f};

// log Poisson pmf, float32, cancellation free for large lam (see oracle/noise.py)
__device__ __forceinline__ float poisson_logpmf(float k, float lam, float li, float lf, float loglam) {
  if (k < 12.0f) return -lam + k * loglam - kLogFact[(int)k];
  const float dk = (li - k) + lf;  // lam - k
  const float t = k * log1pf(dk / k);
  const float inv = 1.0f / k;
  const float corr = inv * (1.0f / 12.0f - inv * inv * (1.0f / 360.0f));
  return t - dk - 0.5f * logf(6.2831853071795860f * k) - corr;
}

// Uniform-word stream of one pixel: words 2,3 of block 0, then blocks 1,2,...
struct PixelStream {
  uint32_t item, seq, sw, k0, k1;
  uint4 cur;
  int q;
  __device__ __forceinline__ uint32_t next() {
    uint32_t w;
    if (q < 2) {
      w = q == 0 ? cur.z : cur.w;
    } else {
      const int r = (q - 2) & 3;
      if (r == 0) cur = philox4x32_10(item, 1u + (uint32_t)((q - 2) >> 2), seq, sw, k0, k1);
      w = r == 0 ? cur.x : r == 1 ? cur.y : r == 2 ? cur.z : cur.w;
    }
    ++q;
    return w;
  }
};

// PTRS set-up and log-pmf table for ONE lambda.  The V1 renderer multiplies every pixel by Poisson(pn)/pn with the same
// pn (helpersGeneration.py:316-317), so the set-up (sqrt, 2 logs, 3 divisions) is done once per thread instead of once per
// pixel and the right-hand side of the acceptance test, log pmf(k), comes from a shared-memory table filled with
// poisson_logpmf itself -- the values, hence every draw, are bit-identical to poisson_draw(lam, st).  ncu (source view,
// r01_render): the per-pixel set-up and the three logf + log1pf of the slow acceptance path were 28 % of the renderer's
// instructions, executed by ~10 of 32 lanes (a third of the pixels miss the squeeze test).
constexpr int kPoissonTable = 256;
struct PoissonConst {
  float lam, loglam, b, a, log_invalpha, vr, li, lf;
  int k0;       // the table holds log pmf(k) for k in [k0, k0 + kPoissonTable)
  bool ptrs;    // lam >= 10
};
__device__ __forceinline__ PoissonConst poisson_setup(float lam) {
  PoissonConst c;
  c.lam = lam;
  c.ptrs = lam >= 10.0f;
  const float slam = sqrtf(lam);
  c.loglam = logf(lam);
  c.b = __fadd_rn(0.931f, __fmul_rn(2.53f, slam));
  c.a = __fadd_rn(-0.059f, __fmul_rn(0.02483f, c.b));
  c.log_invalpha = logf(1.1239f + 1.1328f / (c.b - 3.4f));
  c.vr = 0.9277f - 3.6224f / (c.b - 2.0f);
  c.li = floorf(lam);
  c.lf = lam - c.li;
  c.k0 = lam > 128.0f ? (int)lam - 128 : 0;
  return c;
}
__device__ __forceinline__ void poisson_fill_table(const PoissonConst& c, float* __restrict__ tab, int tid, int nthreads) {
  for (int i = tid; i < kPoissonTable; i += nthreads) tab[i] = poisson_logpmf((float)(c.k0 + i), c.lam, c.li, c.lf, c.loglam);
}

__device__ __forceinline__ float poisson_draw(float lam, PixelStream& st);

// poisson_draw(c.lam, st) with the launch-constant pieces taken from c / tab (same arithmetic, same draws)
__device__ __forceinline__ float poisson_draw_const(const PoissonConst& c, const float* __restrict__ tab, PixelStream& st) {
  if (!c.ptrs) return poisson_draw(c.lam, st);
  for (int it = 0; it < 512; ++it) {
    const float U = u01(st.next()) - 0.5f;
    const float V = u01(st.next());
    const float us = 0.5f - fabsf(U);
    const float d = __fadd_rn(__fmul_rn(__fadd_rn(__fdiv_rn(__fmul_rn(2.0f, c.a), us), c.b), U), 0.43f);
    const float kf = c.li + floorf(c.lf + d);
    if (us >= 0.07f && V <= c.vr) return kf;
    if (!(kf >= 0.0f) || !isfinite(kf) || (us < 0.013f && V > us)) continue;
    const float lhs = logf(V) + c.log_invalpha - logf(c.a / (us * us) + c.b);
    const int ki = (int)kf - c.k0;
    const float rhs = (ki >= 0 && ki < kPoissonTable) ? tab[ki] : poisson_logpmf(kf, c.lam, c.li, c.lf, c.loglam);
    if (lhs <= rhs) return kf;
  }
  return c.li;
}

// numpy's random_poisson (legacy-distributions.c): multiplication method below 10,
// Hoermann PTRS above, float32.
__device__ __forceinline__ float poisson_draw(float lam, PixelStream& st) {
  if (!(lam > 0.0f)) return 0.0f;
  if (lam < 10.0f) {
    const float enlam = expf(-lam);
    float prod = 1.0f, x = 0.0f;
    for (int it = 0; it < 4096; ++it) {
      prod *= u01(st.next());
      if (prod > enlam) x += 1.0f; else break;
    }
    return x;
  }
  const float slam = sqrtf(lam), loglam = logf(lam);
  const float b = __fadd_rn(0.931f, __fmul_rn(2.53f, slam));
  const float a = __fadd_rn(-0.059f, __fmul_rn(0.02483f, b));
  const float invalpha = 1.1239f + 1.1328f / (b - 3.4f);
  const float vr = 0.9277f - 3.6224f / (b - 2.0f);
  const float li = floorf(lam), lf = lam - li;
  for (int it = 0; it < 512; ++it) {
    const float U = u01(st.next()) - 0.5f;
    const float V = u01(st.next());
    const float us = 0.5f - fabsf(U);
    // explicit roundings (no FMA contraction) so the oracle reproduces k exactly
    const float d = __fadd_rn(__fmul_rn(__fadd_rn(__fdiv_rn(__fmul_rn(2.0f, a), us), b), U), 0.43f);
    const float kf = li + floorf(lf + d);
    if (us >= 0.07f && V <= vr) return kf;
    if (!(kf >= 0.0f) || !isfinite(kf) || (us < 0.013f && V > us)) continue;
    const float lhs = logf(V) + logf(invalpha) - logf(a / (us * us) + b);
    if (lhs <= poisson_logpmf(kf, lam, li, lf, loglam)) return kf;
  }
  return li;
}
