"""CPU tests (-m "not gpu"): the render oracle against the committed goldens that
oracle/make_golden.py produced from the unmodified reference, and -- when
/root/reference is present -- against the reference itself."""
import os

import numpy as np
import pytest

from oracle import refshim, render_oracle as ro
from oracle.noise import NumpyNoise, PhiloxNoise
from oracle.make_golden import C3_PROPS

CLEAN = dict(C3_PROPS, background_intensity=[0, 0], poisson_noise=-1)
PSF = [2, 1.75, 1.5, 1.25, 1]
NOISE = [0, 1 / 50, 1 / 25, 1 / 20, 1 / 10, 1 / 5]
PSFNOISE_PROPS = dict(C3_PROPS, particle_intensity=[5000, 500], background_intensity=[5000, 0])
FRAMERATE_PROPS = dict(C3_PROPS, output_size=13)


def relmax(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.fixture(scope="module")
def gold(golden_dir):
    return (np.load(os.path.join(golden_dir, "render_inputs.npz")),
            np.load(os.path.join(golden_dir, "render_golden.npz")),
            np.load(os.path.join(golden_dir, "render_noise_stats.npz")))


def test_literal_is_bit_exact(gold):
    inp, g, _ = gold
    out = ro.render_v1(inp["traj30"][:2], 10, True, CLEAN, mode="literal")
    assert out.dtype == np.float32
    assert np.array_equal(out, g["v1_p9_center"][:2])
    out = ro.render_v1(inp["traj30"][:1], 10, False, CLEAN, mode="literal")
    assert np.array_equal(out, g["v1_p9_nocenter"][:1])


@pytest.mark.parametrize("key,sl,n,center,over", [
    ("v1_p9_center", slice(0, 8), 10, True, {}),
    ("v1_p9_nocenter", slice(0, 2), 10, False, {}),
    ("v1_p13_center", slice(0, 2), 10, True, {"output_size": 13}),
    ("v1_p8_center", slice(0, 1), 10, True, {"output_size": 8}),
    ("v1_p7u10_n15", slice(0, 1), 15, True, {"output_size": 7, "upsampling_factor": 10}),
])
def test_separable_matches_reference(gold, key, sl, n, center, over):
    inp, g, _ = gold
    out = ro.render_v1(inp["traj30"][sl], n, center, dict(CLEAN, **over), mode="separable")
    assert out.shape == g[key].shape
    # layout: the brightest pixel of every frame sits where the reference puts it
    assert np.array_equal(out.reshape(*out.shape[:2], -1).argmax(-1), g[key].reshape(*out.shape[:2], -1).argmax(-1))
    assert relmax(out, g[key]) < 1e-6            # north_star: <= 1e-5 relative in fp32


def test_flip_side_effect_and_sum(gold):
    inp, g, _ = gold
    assert np.allclose(g["v1_flipped_input_y"], -inp["traj30"][:, :4, 1])
    # SURVEY.md 8c smoke values of the reference
    assert abs(float(g["v1_p9_center"].sum(dtype=np.float64)) - 9825153.410454) < 1e-2
    assert abs(float(g["v1_p9_center"][0, 0, 4, 4]) - 3566.737549) < 1e-3


def test_background_and_normalise(gold):
    inp, g, _ = gold
    v = ro.render_v1(inp["traj30"][:2], 10, True, dict(C3_PROPS, poisson_noise=-1))
    vn, meta = ro.normalize_images(v, 1420, 290, 6000)
    assert vn.dtype == np.float32 and meta == (1420, 290, 6000)
    assert np.abs(vn - g["v1_p9_bgmean_norm"]).max() < 2e-6
    with pytest.raises(ValueError):
        ro.normalize_images(v, 10, 0, 10)


def test_psfnoise_and_framerate_mean(gold):
    inp, g, _ = gold
    out = ro.render_psfnoise(inp["traj20"][:1], 10, True, PSFNOISE_PROPS, PSF, NOISE)
    assert out.shape == (1, 5, 6, 20, 9, 9)
    assert relmax(out, g["psfnoise_mean"]) < 1e-6
    # the double-background quirk: noise index >= 1 carries the background twice
    assert abs(out[:, :, 1].mean() - out[:, :, 0].mean() - 5000) < 1.0
    fr = ro.render_framerates(inp["traj30b"][:1], [5, 10, 15, 20, 30, 50], True, FRAMERATE_PROPS)
    assert fr.shape == (1, 6, 60, 13, 13)
    assert np.abs(fr - g["framerate_mean"]).max() < 2e-6
    assert np.all(fr[:, 5, 6:] == 0)             # zero padding beyond T // n frames


def test_errors():
    t = np.zeros((1, 25, 2))
    with pytest.raises(Exception, match="divisble"):
        ro.render_v1(t, 10, True, CLEAN)
    with pytest.raises(Exception, match="No settings given"):
        ro.render_psfnoise(np.zeros((1, 20, 2)), 10, True, PSFNOISE_PROPS, [], [])


def _ks_from_quantiles(sample, qvals):
    """sup |F_sample - F_ref| with F_ref given by its quantile table."""
    q = np.linspace(0, 1, len(qvals))
    s = np.sort(sample.ravel())
    f_s = np.searchsorted(s, qvals, side="right") / s.size
    return float(np.abs(f_s - q).max())


@pytest.mark.parametrize("source", ["philox", "numpy"])
def test_noisy_v1_statistics(gold, source):
    inp, _, st = gold
    R = 6
    outs = []
    for r in range(R):
        nz = PhiloxNoise(1000 + r) if source == "philox" else NumpyNoise(r)
        outs.append(ro.render_v1(inp["traj30"], 10, True, C3_PROPS, noise=nz))
    v = np.stack(outs).astype(np.float64)
    assert abs(v.mean() - st["v1_mean"]) / st["v1_mean"] < 3e-3
    assert abs(v.std() - st["v1_std"]) / st["v1_std"] < 1e-2
    assert np.abs(v.mean(axis=(0, 1, 2)) - st["v1_pix_mean"]).max() / st["v1_pix_mean"].max() < 1e-2
    assert _ks_from_quantiles(v, st["v1_quantiles"]) < 0.01


def test_noisy_psfnoise_statistics(gold):
    inp, _, st = gold
    outs = [ro.render_psfnoise(inp["traj20"], 10, True, PSFNOISE_PROPS, PSF, NOISE, noise=PhiloxNoise(77 + r))
            for r in range(3)]
    v = np.stack(outs).astype(np.float64)
    m = v.mean(axis=(0, 1, 4, 5, 6))
    s = v.std(axis=(0, 1, 4, 5, 6))
    assert np.abs(m / st["psf_mean"] - 1).max() < 5e-3
    assert np.abs(s / st["psf_std"] - 1).max() < 3e-2
    for i in range(5):
        for j in range(6):
            assert _ks_from_quantiles(v[:, :, i, j], st["psf_quantiles"][i, j]) < 0.03


def test_poisson_sampler_chi_square():
    from scipy import stats
    n = 120000
    for lam in (3.0, 10.0, 100.0):
        _, pois = PhiloxNoise(5).pixel(3, n)
        k = pois(np.full(n, lam, dtype=np.float32)).astype(np.int64)
        lo, hi = int(stats.poisson.ppf(1e-4, lam)), int(stats.poisson.ppf(1 - 1e-4, lam))
        obs = np.array([(k == i).sum() for i in range(lo, hi + 1)], dtype=np.float64)
        exp = stats.poisson.pmf(np.arange(lo, hi + 1), lam) * n
        keep = exp > 20
        chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
        assert chi2 < stats.chi2.ppf(1 - 1e-4, keep.sum() - 1), (lam, chi2)
    # lam ~ 1e6 (PSFNoise regime): moments
    _, pois = PhiloxNoise(6).pixel(0, n)
    k = pois(np.full(n, 1.5e6, dtype=np.float32)).astype(np.float64)
    assert abs(k.mean() - 1.5e6) < 5 * np.sqrt(1.5e6 / n)
    assert abs(k.var() / 1.5e6 - 1) < 0.03


def test_alias_table_matches_the_library_and_the_poisson_pmf():
    """V1 renderer noise (round 2): one uniform word per pixel through a 256-entry alias table.  The numpy restatement of the
    table must equal the C library's (host-only entry point, no GPU) bit for bit, and draws from it must follow Poisson(lam)."""
    import ctypes
    from scipy import stats
    from moleculardiffusion_mivit_b200 import _lib
    from oracle.noise import alias_draw, poisson_alias_table
    L = _lib.lib()
    rng = np.random.default_rng(9)
    for lam in (100.0, 0.3, 5.0, 50.0, 129.0, 150.5, 379.9):
        ent, k0 = (ctypes.c_uint32 * 256)(), ctypes.c_int32()
        assert L.mivit_poisson_alias_table(lam, ent, ctypes.byref(k0)) == 0
        tab = poisson_alias_table(lam)
        assert np.array_equal(np.frombuffer(ent, dtype=np.uint32), tab[0]) and k0.value == tab[1], lam
        n = 400000
        k = alias_draw(tab, rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)).astype(np.int64)
        lo, hi = int(stats.poisson.ppf(1e-4, lam)), int(stats.poisson.ppf(1 - 1e-4, lam))
        obs = np.array([(k == i).sum() for i in range(lo, hi + 1)], dtype=np.float64)
        exp = stats.poisson.pmf(np.arange(lo, hi + 1), lam) * n
        keep = exp > 20
        chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
        assert chi2 < stats.chi2.ppf(1 - 1e-4, max(keep.sum() - 1, 1)), (lam, chi2)
    ent, k0 = (ctypes.c_uint32 * 256)(), ctypes.c_int32()
    assert L.mivit_poisson_alias_table(500.0, ent, ctypes.byref(k0)) != 0 and poisson_alias_table(500.0) is None   # PTRS regime
    assert poisson_alias_table(-1.0) is None
    # exact table mass: sum over entries of the accept / alias probabilities reproduces the pmf to the 2^-24 quantisation
    ent_np, k0v = poisson_alias_table(100.0)
    pm = np.zeros(256)
    thr = (ent_np & 0xFFFFFF).astype(np.float64) / 2.0 ** 24
    np.add.at(pm, np.arange(256), thr / 256)
    np.add.at(pm, (ent_np >> 24).astype(np.int64), (1 - thr) / 256)
    assert np.abs(pm - stats.poisson.pmf(k0v + np.arange(256), 100.0)).max() < 1e-7


def test_pair_layout_streams_are_independent_normals_and_poissons():
    """PhiloxNoise.pixel_v1: every pixel gets its own normal and its own Poisson draw (no value shared inside a pair), also when
    P is odd (the last pair of a row has one pixel) and in the PTRS regime (lam > 380)."""
    z, k = PhiloxNoise(3).pixel_v1(7, 40, 13, 100.0)
    assert z.shape == k.shape == (40, 13, 13)
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03
    assert abs(np.corrcoef(z[:, :, 0:12:2].ravel(), z[:, :, 1:13:2].ravel())[0, 1]) < 0.04     # left / right of a pair
    assert abs(k.mean() - 100) < 0.4 and abs(k.var() / 100 - 1) < 0.06
    assert abs(np.corrcoef(k[:, :, 0:12:2].ravel(), k[:, :, 1:13:2].ravel())[0, 1]) < 0.04
    assert abs(np.corrcoef(z.ravel(), k.ravel())[0, 1]) < 0.04
    z2, k2 = PhiloxNoise(3).pixel_v1(7, 6, 9, 1000.0)
    assert abs(k2.mean() - 1000) < 8 and abs(k2.var() / 1000 - 1) < 0.2
    z3, k3 = PhiloxNoise(3).pixel_v1(7, 6, 9, -1)
    assert k3 is None and z3.shape == (6, 9, 9)


@pytest.mark.parametrize("source", ["philox", "numpy"])
def test_noisy_p13_framerate_statistics(golden_dir, gold, source):
    """The props bench.py renders (Framerate experiment, P = 13, n = 10): moments, per-pixel means and KS against statistics of
    the reference's own np.random path (oracle/make_golden_r2.py)."""
    inp, _, _ = gold
    st = np.load(os.path.join(golden_dir, "render_noise_stats_p13.npz"))
    outs = [ro.render_v1(inp["traj30"], 10, True, FRAMERATE_PROPS, noise=PhiloxNoise(2000 + r) if source == "philox" else NumpyNoise(50 + r))
            for r in range(4)]
    v = np.stack(outs).astype(np.float64)
    assert abs(v.mean() - st["mean"]) / st["mean"] < 3e-3
    assert abs(v.std() - st["std"]) / st["std"] < 1e-2
    assert np.abs(v.mean(axis=(0, 1, 2)) - st["pix_mean"]).max() / st["pix_mean"].max() < 1.5e-2
    assert _ks_from_quantiles(v, st["quantiles"]) < 0.01


@pytest.mark.skipif(not refshim.reference_available(), reason="reference checkout not present")
def test_against_live_reference(gold):
    inp, g, _ = gold
    from oracle.make_golden import _Deterministic, _quiet
    gen, _ = refshim.import_reference()
    with _Deterministic():
        t = inp["traj30"][2:4].copy()
        ref = _quiet(gen.trajectories_to_video, t, 10, True, CLEAN)
    assert np.array_equal(ref, g["v1_p9_center"][2:4])
    assert np.array_equal(ro.render_v1(inp["traj30"][2:4], 10, True, CLEAN, mode="literal"), ref)


# ---- trajectories_to_video_multiple_settings (helpersGeneration.py:422-540) -------------------------------
MULTI_CASES = [("p9_center", slice(0, 3), 10, True, {}), ("p13_nocenter_n15", slice(3, 4), 15, False, {"output_size": 13})]


@pytest.mark.parametrize("key,sl,n,center,over", MULTI_CASES)
def test_multi_settings_oracle_matches_reference_golden(golden_dir, key, sl, n, center, over):
    inp = np.load(os.path.join(golden_dir, "render_inputs.npz"))["traj30"]
    g = np.load(os.path.join(golden_dir, "render_multi_golden.npz"))
    props = dict(C3_PROPS, **over)
    lit = ro.render_multi(inp[sl], n, center, props, mode="literal")
    sep = ro.render_multi(inp[sl], n, center, props)
    for name, a, b in zip(("none", "gauss", "poisson", "filter"), lit, sep):
        ref = g["%s/%s" % (key, name)]
        assert a.dtype == np.float32 and a.shape == ref.shape
        assert np.array_equal(a, ref), name                       # literal restatement: bit exact (filter: scipy's algorithm)
        assert relmax(b, ref) < 1e-6, name                        # separable closed form


def test_multi_settings_noisy_moments_vs_reference(golden_dir):
    """Philox streams instead of np.random: first / second moments of the four outputs against the reference's own noisy runs."""
    inp = np.load(os.path.join(golden_dir, "render_inputs.npz"))["traj30"]
    g = np.load(os.path.join(golden_dir, "render_multi_golden.npz"))
    runs = [ro.render_multi(inp[:8], 10, True, C3_PROPS, noise=PhiloxNoise(500 + r)) for r in range(6)]
    for i, name in enumerate(("none", "gauss", "poisson", "filter")):
        v = np.stack([r[i] for r in runs]).astype(np.float64)
        assert abs(v.mean() - float(g["noisy/%s_mean" % name])) < 0.02 * float(g["noisy/%s_mean" % name]), name
        assert abs(v.std() - float(g["noisy/%s_std" % name])) < 0.03 * float(g["noisy/%s_std" % name]), name
        assert abs(v.std(axis=0).mean() - float(g["noisy/%s_pixstd" % name])) < 0.08 * float(g["noisy/%s_pixstd" % name]), name


def test_norm_rl_oracle_matches_reference_golden(golden_dir):
    """trajs_to_vid_norm_rl (helpersGeneration.py:635-658): with the reference's own fftconvolve the restatement is bit exact;
    with the direct float64 sum (what the CUDA kernel computes) the Richardson-Lucy/TV channels agree to the single-precision
    noise of scipy's FFT of the float32 estimate, amplified over 3 / 6 / 11 iterations."""
    inp = np.load(os.path.join(golden_dir, "render_inputs.npz"))["traj30"][:2]
    ref = np.load(os.path.join(golden_dir, "render_norm_rl_golden.npz"))["out"]
    lit = ro.render_norm_rl(inp, 10, True, C3_PROPS, [2, 5, 10], mode="literal", conv="fft")
    assert lit.shape == ref.shape == (2, 7, 30, 9, 9) and lit.dtype == np.float32
    assert np.array_equal(lit, ref)
    direct = ro.render_norm_rl(inp, 10, True, C3_PROPS, [2, 5, 10])
    assert relmax(direct[:, :4], ref[:, :4]) < 1e-6
    for k, tol in ((4, 1e-4), (5, 5e-4), (6, 1.5e-3)):
        assert np.abs(direct[:, k] - ref[:, k]).max() < tol, (k, np.abs(direct[:, k] - ref[:, k]).max())
