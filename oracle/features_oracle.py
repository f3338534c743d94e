"""TEST INFRASTRUCTURE ONLY -- numpy/scipy restatement of the reference's trajectory-feature producer
(the ViT's `features` input of the ImagesFeatures experiment):

    helpers/helpersFeatures.py:448-520  compute_diffusion_features   (25 features, order of :7-34)
      :102-132 msd   :135-191 fit_diffusion_scaling (scipy curve_fit, trf, bounds)   :194-218 efficiency
      :221-247 fractal_dim   :250-284 gaussianity   :287-324 kurtosis   :327-347 msd_ratio
      :350-378 trappedness   :381-402 convex_hull_area   :404-446 dot products / step lengths
    helpers/helpersGeneration.py:48-74  average_trajectories_frames

Parity PINNED: tests/test_oracle_features.py checks this file against tests/golden/features_golden.npz, generated
from the unmodified reference by oracle/make_golden.py (incl. the two 25-vectors stored in the reference's
tests/models_tests/FeaturesTests.ipynb cell 2).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this module; the product computes the features in csrc/features.cu.

Vectorised where the reference loops in Python; the power-law fit calls scipy.optimize.curve_fit with the
reference's exact arguments (the reference's result is whatever trf converges to, so the oracle uses trf too)."""
import numpy as np

N_FEATURES = 25


def average_frames(trajectories, n):
    """helpersGeneration.py:48-74: mean over groups of n sub-positions -> (N, T//n, 2)."""
    N, T, d = trajectories.shape
    F = T // n
    return trajectories[:, :F * n].reshape(N, F, n, d).mean(axis=2)


def msd(x, y):
    """helpersFeatures.py:102-132 (frac = 0.5; all lags when the trajectory has <= 20 points)."""
    L = len(x)
    N = int(L * 0.5) if L > 20 else L
    return np.array([np.mean((x[lag:] - x[:-lag]) ** 2 + (y[lag:] - y[:-lag]) ** 2) for lag in range(1, N)])


def fit_power_law(msds, dt):
    """helpersFeatures.py:135-191: MSD = 4 D t^alpha + offset, trf with bounds; returns ((D, alpha, offset), r2)."""
    from scipy.optimize import curve_fit
    t = np.arange(1, len(msds) + 1) * dt

    def model(tt, D, alpha, offset):
        return 4 * D * tt ** alpha + offset

    try:
        p, _ = curve_fit(model, t, msds, p0=[msds[0] / (4 * dt), 1, 0.001], bounds=([1e-5, 1e-5, 0], [np.inf, 10, np.inf]),
                         method="trf", maxfev=10000)
        res = msds - model(t, *p)
        r2 = 1 - np.sum(res ** 2) / np.sum((msds - np.mean(msds)) ** 2)
    except (RuntimeError, ValueError):
        p, r2 = (0, 0), 0
    return p, r2


def features(traj, dt=1.0):
    """helpersFeatures.py:448-520 for one (L, 2) float64 trajectory."""
    x, y = traj[:, 0], traj[:, 1]
    L = len(x)
    if L < 3:
        return np.full(N_FEATURES, np.nan)
    m = msd(x, y)
    dx, dy = np.diff(x), np.diff(y)
    sl = np.sqrt(dx ** 2 + dy ** 2)
    dots = dx[:-1] * dx[1:] + dy[:-1] * dy[1:] if L > 2 else np.array([0.0])
    d2 = (x[:, None] - x[None, :]) ** 2 + (y[:, None] - y[None, :]) ** 2
    maxd = d2.max()
    p, r2 = fit_power_law(m, dt)
    D, alpha = p[0], p[1]
    top, bottom = (x[-1] - x[0]) ** 2 + (y[-1] - y[0]) ** 2, np.sum(sl ** 2)
    if bottom == 0:
        eff_log, eff = -np.inf, 0
    else:
        eff = top / ((L - 1) * bottom)
        with np.errstate(divide="ignore"):
            eff_log = np.log(eff)
    total = sl.sum()
    fractal = 1 if total == 0 else np.log(L) / (np.log(L) + np.log(np.sqrt(maxd) / total))
    gn = []
    for lag in range(1, len(m) + 1):
        if lag >= L:
            break
        r4 = np.mean((x[lag:] - x[:-lag]) ** 4 + (y[lag:] - y[:-lag]) ** 4)
        if m[lag - 1] > 0:
            gn.append(r4 / (2 * m[lag - 1] ** 2))
    gauss = np.mean(gn) if gn else np.nan
    cov = np.cov(x, y)
    lam = 0.5 * (cov[0, 0] + cov[1, 1]) + np.sqrt((0.5 * (cov[0, 0] - cov[1, 1])) ** 2 + cov[0, 1] ** 2)
    v = np.array([cov[0, 1], lam - cov[0, 0]])
    if not np.any(v):
        v = np.array([1.0, 0.0]) if cov[0, 0] >= cov[1, 1] else np.array([0.0, 1.0])
    v = v / np.linalg.norm(v)
    proj = v[0] * x + v[1] * y
    c = proj - proj.mean()
    kurt = np.mean(c ** 4) / np.mean(c ** 2) ** 2
    ratio = np.mean(m[:-1] / m[1:] - np.arange(1, len(m)) / np.arange(2, len(m) + 1)) if len(m) >= 2 else np.nan
    r0 = np.sqrt(maxd) / 2
    trapped = 0 if (r0 == 0 or D == 0) else 1 - np.exp(0.2045 - 0.25117 * (D * L) / r0 ** 2)
    mean_sl = sl.mean()
    return np.array([
        alpha, D, r2, eff_log, eff, fractal, gauss, kurt, ratio, trapped, L, mean_sl, m.mean(),
        dots.mean(), np.mean(np.sign(dots[1:]) == np.sign(dots[:-1])) if len(dots) > 1 else np.nan, np.mean(np.sign(dots) > 0),
        total, sl.min(), sl.max(), sl.max() - sl.min(), total / L,
        np.std(sl, ddof=1) / mean_sl if (mean_sl > 0 and len(sl) > 1) else np.nan,
        np.sum(sl < 0.1) / len(sl), np.sum(sl > 0.4) / len(sl), hull_area(x, y)])


def hull_area(x, y):
    """helpersFeatures.py:381-402 (scipy ConvexHull(...).volume): Andrew's monotone chain + shoelace."""
    pts = sorted(set(zip(x.tolist(), y.tolist())))
    if len(pts) < 3:
        return 0.0

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    h = lower[:-1] + upper[:-1]
    if len(h) < 3:
        return 0.0
    hx, hy = np.array([q[0] for q in h]), np.array([q[1] for q in h])
    return 0.5 * abs(np.dot(hx, np.roll(hy, -1)) - np.dot(hy, np.roll(hx, -1)))


def features_batch(trajs, dt=1.0):
    return np.stack([features(t, dt) for t in trajs])
