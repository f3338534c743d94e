"""TEST INFRASTRUCTURE ONLY -- noise sources for the render oracle.

Three interchangeable sources feed oracle/render_oracle.py:
  * PhiloxNoise  -- the exact counter streams of the CUDA renderer (csrc/philox.cuh,
                    csrc/render.cu); lets tests compare noisy GPU frames draw-for-draw.
  * NumpyNoise   -- numpy Generator draws; statistically what the reference does with
                    np.random.normal / np.random.poisson
                    (helpers/helpersGeneration.py:300,312,317;
                     Experiments/PSFNoise/trainSettingsPSFNoise.py:279,303,305).
  * MeanNoise    -- every normal returns its mean (z = 0): the deterministic setting
                    used for noise-free goldens (SURVEY.md section 8c smoke values).

Poisson sampler = numpy's own algorithm (legacy-distributions.c random_poisson:
multiplication method for lam < 10, Hoermann's PTRS transformed rejection for
lam >= 10), restated in float32 with the log-pmf evaluated in a cancellation-free
form so that lam ~ 1e6 (PSFNoise: Poisson(x * 100), x ~ 1.5e4) stays accurate in fp32.
"""
import math

import numpy as np

from . import philox as px

_LOGFACT = np.array([math.lgamma(k + 1.0) for k in range(16)], dtype=np.float32)
F32 = np.float32


def poisson_logpmf_f32(k, lam, li, lf, loglam):
    """log(Poisson pmf) in float32, same operation order as csrc/philox.cuh.
    k, lam, li=floor(lam), lf=lam-li, loglam=log(lam) are float32 arrays."""
    k = k.astype(F32)
    small = k < F32(12.0)
    ks = np.clip(k, 0, 15).astype(np.int64)
    out_small = (-lam + k * loglam - _LOGFACT[ks]).astype(F32)
    x = np.maximum(k, F32(12.0))
    dk = ((li - x) + lf).astype(F32)                      # lam - k without cancellation
    t = (x * np.log1p((dk / x).astype(F32)).astype(F32)).astype(F32)
    inv = (F32(1.0) / x).astype(F32)
    corr = (inv * (F32(1.0 / 12.0) - inv * inv * F32(1.0 / 360.0))).astype(F32)
    out_big = (t - dk - F32(0.5) * np.log((F32(6.2831853071795860) * x).astype(F32)).astype(F32) - corr).astype(F32)
    return np.where(small, out_small, out_big).astype(F32)


def poisson_from_uniform_words(lam, word_fn, max_rounds=64):
    """lam: float32 array (>=0).  word_fn(q) -> uint32 array (same shape) giving the
    q-th word of each element's Poisson uniform stream.  Returns float32 counts."""
    lam = np.asarray(lam, dtype=F32)
    out = np.zeros(lam.shape, dtype=F32)
    done = lam <= 0
    # ---- multiplication method, lam < 10 (Knuth)
    small = (~done) & (lam < F32(10.0))
    if small.any():
        enlam = np.exp(-lam.astype(F32)).astype(F32)
        prod = np.ones(lam.shape, dtype=F32)
        cnt = np.zeros(lam.shape, dtype=F32)
        active = small.copy()
        q = 0
        while active.any():
            u = px.u01(word_fn(q))
            prod = np.where(active, (prod * u).astype(F32), prod)
            fin = active & ~(prod > enlam)
            cnt = np.where(active & ~fin, cnt + 1, cnt)
            active &= ~fin
            q += 1
            if q > 4096:
                raise RuntimeError("poisson multiplication method did not terminate")
        out = np.where(small, cnt, out)
        done |= small
    big = ~done
    if big.any():
        lam_b = np.where(big, lam, F32(100.0)).astype(F32)
        slam = np.sqrt(lam_b).astype(F32)
        loglam = np.log(lam_b).astype(F32)
        b = (F32(0.931) + F32(2.53) * slam).astype(F32)
        a = (F32(-0.059) + F32(0.02483) * b).astype(F32)
        invalpha = (F32(1.1239) + F32(1.1328) / (b - F32(3.4))).astype(F32)
        vr = (F32(0.9277) - F32(3.6224) / (b - F32(2.0))).astype(F32)
        li = np.floor(lam_b).astype(F32)
        lf = (lam_b - li).astype(F32)
        active = big.copy()
        res = np.zeros(lam.shape, dtype=F32)
        j = 0
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            while active.any():
                U = (px.u01(word_fn(2 * j)) - F32(0.5)).astype(F32)
                V = px.u01(word_fn(2 * j + 1))
                us = (F32(0.5) - np.abs(U)).astype(F32)
                d = (((F32(2.0) * a / us).astype(F32) + b) * U + F32(0.43)).astype(F32)
                kf = (li + np.floor((lf + d).astype(F32))).astype(F32)
                fast = (us >= F32(0.07)) & (V <= vr)
                rej = (kf < 0) | ((us < F32(0.013)) & (V > us)) | ~np.isfinite(kf)
                lhs = (np.log(V).astype(F32) + np.log(invalpha).astype(F32)
                       - np.log((a / (us * us).astype(F32) + b).astype(F32)).astype(F32)).astype(F32)
                rhs = poisson_logpmf_f32(np.where(np.isfinite(kf) & (kf >= 0), kf, F32(0)), lam_b, li, lf, loglam)
                acc = fast | (~rej & (lhs <= rhs))
                take = active & acc
                res = np.where(take, kf, res)
                active &= ~acc
                j += 1
                if j > max_rounds * 8:
                    raise RuntimeError("PTRS did not terminate")
        out = np.where(big, res, out)
    return out.astype(F32)


ALIAS_ENTRIES = 256
ALIAS_MAX_LAMBDA = 380.0


def poisson_alias_table(lam):
    """Walker / Vose alias table of Poisson(lam) on k in [k0, k0 + 256): the float64 construction of csrc/render.cu
    build_alias_table restated operation by operation (pmf by the recurrence p(k+1) = p(k) lam / (k+1) from the mode).
    Returns (entries uint32[256] = alias << 24 | threshold(24 bit), k0) or None when lam is outside (0, 380]."""
    lam = float(lam)
    if not (lam > 0.0) or lam > ALIAS_MAX_LAMBDA:
        return None
    K = ALIAS_ENTRIES
    mode = int(math.floor(lam))
    k0 = mode - K // 2 if mode > K // 2 else 0
    p = [0.0] * K
    p[mode - k0] = 1.0
    for k in range(mode, k0 + K - 1):
        p[k + 1 - k0] = p[k - k0] * lam / float(k + 1)
    for k in range(mode, k0, -1):
        p[k - 1 - k0] = p[k - k0] * float(k) / lam
    total = 0.0
    for i in range(K):
        total += p[i]
    sc = [0.0] * K
    prob = [1.0] * K
    alias = list(range(K))
    small, large = [], []
    for i in range(K):
        sc[i] = p[i] / total * float(K)
        (small if sc[i] < 1.0 else large).append(i)
    while small and large:
        sm, lg = small.pop(), large.pop()
        prob[sm] = sc[sm]
        alias[sm] = lg
        sc[lg] = (sc[lg] + sc[sm]) - 1.0
        (small if sc[lg] < 1.0 else large).append(lg)
    ent = np.zeros(K, dtype=np.uint32)
    for i in range(K):
        t = math.floor(prob[i] * 16777216.0 + 0.5)
        t = min(max(t, 0.0), 16777215.0)
        ent[i] = (alias[i] << 24) | int(t)
    return ent, k0


def alias_draw(table, words):
    ent, k0 = table
    words = np.asarray(words, dtype=np.uint32)
    j = (words >> np.uint32(24)).astype(np.int64)
    e = ent[j]
    acc = (words & np.uint32(0xFFFFFF)) < (e & np.uint32(0xFFFFFF))
    return (k0 + np.where(acc, j, (e >> np.uint32(24)).astype(np.int64))).astype(F32)


class PhiloxNoise:
    """Counter streams of the CUDA renderer."""

    def __init__(self, seed):
        self.k0, self.k1 = px.seed_key(seed)

    def _words(self, item, block, seq, stream, variant=0):
        return px.philox4x32_10(item, block, seq, px.stream_word(stream, variant), self.k0, self.k1)

    def intensity_z(self, seq, n_items, v1=False):
        """One standard normal per item (V1: item = frame*n+p; PSFNoise: item = frame).  v1=True: the V1 renderer's Box-Muller
        with the angle in (-pi, pi] (csrc/philox.cuh box_muller_fast)."""
        item = np.arange(n_items, dtype=np.uint32)
        w = self._words(item, np.uint32(0), np.uint32(seq), px.STREAM_INTENSITY)
        if v1:
            return px.box_muller_shifted(w[0], w[1])[0]
        z, _ = px.box_muller(w[0], w[1])
        return z

    def pixel(self, seq, n_pixels, variant=0):
        """Returns (z_background[n_pixels], poisson(lam)->counts)."""
        item = np.arange(n_pixels, dtype=np.uint32)
        w0 = self._words(item, np.uint32(0), np.uint32(seq), px.STREAM_PIXEL, variant)
        z, _ = px.box_muller(w0[0], w0[1])
        cache = {0: w0}

        def word_fn(q):
            if q < 2:
                return w0[2 + q]
            blk = 1 + (q - 2) // 4
            if blk not in cache:
                cache[blk] = self._words(item, np.uint32(blk), np.uint32(seq), px.STREAM_PIXEL, variant)
            return cache[blk][(q - 2) % 4]

        def poisson(lam):
            return poisson_from_uniform_words(np.asarray(lam, dtype=F32).reshape(-1), word_fn).reshape(np.shape(lam))

        return z, poisson

    def pixel_v1(self, seq, F, P, lam):
        """"Pair" layout of the V1 renderer (csrc/philox.cuh): one Philox block per pair of horizontally adjacent pixels --
        words x,y -> the two background normals (Box-Muller, angle in (-pi, pi]), words z,w -> the two Poisson(lam) draws
        through the alias table (PTRS on the pixel's own blocks >= 1 when lam > 380).  Returns (z[F,P,P], k[F,P,P] or None)."""
        ppr = (P + 1) // 2
        pairs = ppr * P
        item = np.arange(F * pairs, dtype=np.uint32)
        w = self._words(item, np.uint32(0), np.uint32(seq), px.STREAM_PIXEL)
        zl, zr = px.box_muller_shifted(w[0], w[1])
        z = np.zeros((F, P, 2 * ppr), dtype=F32)
        z[:, :, 0::2] = zl.reshape(F, P, ppr)
        z[:, :, 1::2] = zr.reshape(F, P, ppr)
        z = z[:, :, :P]
        if lam is None or lam == -1:
            return z, None
        table = poisson_alias_table(lam)
        if table is not None:
            k = np.zeros((F, P, 2 * ppr), dtype=F32)
            k[:, :, 0::2] = alias_draw(table, w[2]).reshape(F, P, ppr)
            k[:, :, 1::2] = alias_draw(table, w[3]).reshape(F, P, ppr)
            return z, k[:, :, :P]
        pit = np.arange(F * P * P, dtype=np.uint32)
        cache = {}

        def word_fn(q):            # PixelStream with q starting at 2: blocks 1, 2, ... of item = pixel
            blk = 1 + q // 4
            if blk not in cache:
                cache[blk] = self._words(pit, np.uint32(blk), np.uint32(seq), px.STREAM_PIXEL)
            return cache[blk][q % 4]

        k = poisson_from_uniform_words(np.full(F * P * P, lam, dtype=F32), word_fn)
        return z, k.reshape(F, P, P)


class NumpyNoise:
    """What the reference does: independent numpy draws (statistical parity only)."""

    def __init__(self, seed=None):
        self.rng = np.random.default_rng(seed)

    def intensity_z(self, seq, n_items, v1=False):
        return self.rng.standard_normal(n_items).astype(F32)

    def pixel(self, seq, n_pixels, variant=0):
        z = self.rng.standard_normal(n_pixels).astype(F32)

        def poisson(lam):
            return self.rng.poisson(np.asarray(lam, dtype=np.float64)).astype(F32)

        return z, poisson

    def pixel_v1(self, seq, F, P, lam):
        z = self.rng.standard_normal((F, P, P)).astype(F32)
        if lam is None or lam == -1:
            return z, None
        return z, self.rng.poisson(float(lam), (F, P, P)).astype(F32)


class MeanNoise:
    """Every draw returns its mean: z = 0 and Poisson(lam) -> lam."""

    def intensity_z(self, seq, n_items, v1=False):
        return np.zeros(n_items, dtype=F32)

    def pixel(self, seq, n_pixels, variant=0):
        def poisson(lam):
            return np.asarray(lam, dtype=F32)

        return np.zeros(n_pixels, dtype=F32), poisson

    def pixel_v1(self, seq, F, P, lam):
        z = np.zeros((F, P, P), dtype=F32)
        if lam is None or lam == -1:
            return z, None
        return z, np.full((F, P, P), lam, dtype=F32)
