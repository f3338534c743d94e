"""GPU parity tests (-m gpu) of the CUDA trajectory-feature producer (csrc/features.cu, through the C ABI and the Python
mirrors) against goldens from the unmodified reference and against the oracle on seeded Brownian trajectories.

Tolerances: every feature is float64 arithmetic and agrees to 1e-9 relative, except the three that depend on the
bounded power-law fit -- alpha, D (and trappedness through D) -- where the REFERENCE's own answer is whatever scipy's
trf stops at (ftol = xtol = 1e-8 on a flat alpha-D-offset valley): 2e-3 relative on the reference's validation
trajectories (goldens), 1e-2 on seeded sub-diffusive-looking short tracks (alpha < 0.5, MSD ~ the 1e-3 starting offset),
1e-6 absolute on r_squared -- and the CUDA fit, being the exact constrained minimiser, never has the lower r_squared."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import features_oracle as fo

FIT = [0, 1, 9]          # alpha, D, trappedness
R2 = 2


def check(got, ref, fit_rtol=2e-3):
    assert got.shape == ref.shape
    exact = [i for i in range(25) if i not in FIT + [R2]]
    assert np.allclose(got[:, exact], ref[:, exact], rtol=1e-9, atol=1e-12, equal_nan=True), \
        np.nanmax(np.abs(got[:, exact] - ref[:, exact]) / (np.abs(ref[:, exact]) + 1e-12), axis=0)
    if fit_rtol is None:      # 5 MSD points for 3 parameters: the fit is not identifiable, only its quality is compared
        assert np.all(got[:, R2] >= ref[:, R2] - 1e-9)
        return
    assert np.allclose(got[:, [1, 9]], ref[:, [1, 9]], rtol=fit_rtol, atol=1e-9), (got[:, [1, 9]], ref[:, [1, 9]])
    ident = ref[:, 1] > 2e-5     # D pinned at its lower bound = a non-increasing MSD curve fitted by the offset alone:
    assert np.allclose(got[ident, 0], ref[ident, 0], rtol=fit_rtol, atol=1e-9)   # alpha is then not identifiable (trf ends anywhere)
    assert np.allclose(got[:, R2], ref[:, R2], rtol=0, atol=1e-6 if fit_rtol < 5e-3 else 1e-4)   # where trf stopped early
    assert np.all(got[:, R2] >= ref[:, R2] - 1e-9)       # the exact constrained minimiser never fits worse than trf


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "features_golden.npz"))


@pytest.fixture(scope="module")
def hf():
    from moleculardiffusion_mivit_b200 import helpersFeatures
    return helpersFeatures


@pytest.mark.parametrize("key,dt", [("feat_dt1", 1.0), ("feat_dt01", 0.1)])
def test_reference_goldens(gold, hf, key, dt):
    import torch
    got = hf.features_device(torch.from_numpy(gold["avg"]).cuda(), dt).cpu().numpy()
    check(got, gold[key])
    one = hf.compute_diffusion_features(gold["avg"][3], dt)                  # the reference's per-trajectory call
    assert one.shape == (25,) and np.allclose(one, got[3], rtol=0, atol=0, equal_nan=True)


def test_notebook_vectors_300_points(gold, hf):
    import torch
    got = hf.features_device(torch.from_numpy(gold["long"]).cuda(), 0.1).cpu().numpy()
    check(got, gold["long_feat_dt01"])
    assert np.allclose(got, gold["notebook"], rtol=3e-3, atol=1e-8)


def test_frame_average_and_batch_wrapper(gold, hf):
    import torch
    avg = hf.average_frames_device(torch.from_numpy(gold["traj"]).cuda(), 10).cpu().numpy()
    assert np.allclose(avg, gold["avg"], rtol=0, atol=1e-15)
    got = hf.compute_features_for_multiple_trajectories(gold["traj"], dt=0.1, nPosPerFrame=10)
    check(got, np.nan_to_num(gold["feat_dt01"], nan=0.0))
    with pytest.raises(ValueError, match="cannot reshape"):
        hf.compute_features_for_multiple_trajectories(gold["traj"], dt=0.1, nPosPerFrame=7)


@pytest.mark.parametrize("L,D", [(30, 1.0), (30, 9.0), (20, 3.0), (6, 5.0), (60, 0.05), (128, 7.0)])
def test_seeded_brownian_vs_oracle(hf, L, D):
    import torch
    rng = np.random.default_rng(L * 31 + int(D * 10))
    tr = np.cumsum(rng.normal(0, np.sqrt(2 * D * 0.1), size=(24, L, 2)), axis=1) * 0.12
    got = hf.features_device(torch.from_numpy(tr).cuda(), 0.1).cpu().numpy()
    check(got, fo.features_batch(tr, 0.1), fit_rtol=1e-2 if L >= 12 else None)


def test_edge_cases(hf):
    import torch
    short = hf.features_device(torch.zeros((2, 2, 2), dtype=torch.float64, device="cuda"), 1.0).cpu().numpy()
    assert np.isnan(short).all()                                               # < 3 points: 25 NaNs (:455-456)
    still = np.zeros((1, 12, 2))
    still[0, :, 0] = np.arange(12) * 0.05                                      # collinear: zero hull area, kurtosis of a ramp
    got = hf.features_device(torch.from_numpy(still).cuda(), 1.0).cpu().numpy()
    ref = fo.features_batch(still, 1.0)
    assert got[0, 24] == 0.0 and ref[0, 24] == 0.0
    exact = [3, 4, 5, 7, 10, 11, 12, 13, 15, 16, 17, 18, 19, 20, 22, 23]
    assert np.allclose(got[0, exact], ref[0, exact], rtol=1e-9, atol=1e-12, equal_nan=True)


def test_create_video_and_feature_pairs(gold, golden_dir):
    from moleculardiffusion_mivit_b200 import helpersGeneration as gen
    from oracle.make_golden import C3_PROPS
    t = gold["traj"][:4].copy()
    vids, feats, (tr_out, avg, avg_err) = gen.create_video_and_feature_pairs(t, 10, True, dict(C3_PROPS), (0, 0), 0.1, seed=5)
    assert vids.shape == (4, 30, 9, 9) and vids.dtype == np.float32 and feats.shape == (4, 25) and feats.dtype == np.float64
    assert tr_out is t and np.array_equal(t, gold["pair_traj_after"])          # in-place y flip, returned as is
    assert np.allclose(avg, gold["pair_avg"], rtol=0, atol=1e-15) and np.array_equal(avg, avg_err)
    check(feats, gold["pair_features"])                                        # features of the FLIPPED trajectories
    assert abs(vids.mean() - gold["pair_video_mean"][0]) < 0.02                 # noisy render vs noise-free golden mean
