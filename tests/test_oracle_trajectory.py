"""CPU tests: statistical contract of the trajectory source restatement (SURVEY.md section 4:
fixtures' step std = sqrt(2D); MSD-recovered D)."""
import os

import numpy as np
import pytest

from oracle import refshim
from oracle.trajectory_oracle import brownian_oracle


def _ks(sample, qvals):
    q = np.linspace(0, 1, len(qvals))
    s = np.sort(np.asarray(sample).ravel())
    return float(np.abs(np.searchsorted(s, qvals, side="right") / s.size - q).max())


def test_statistics_match_the_reference_brownian_motion(golden_dir):
    """the third-party andi_datasets generator is absent (bit-level parity unpinned), but the reference's OWN
    in-repo generator brownian_motion (helpers/helpersGeneration.py:9-45: sigma = sqrt(2 D dt / nposframe) :33, cumsum :42)
    runs: its statistics for the training loops' D groups are committed (oracle/make_golden_r2.py) and the device generator's
    restatement must reproduce them -- step std, the step distribution (KS), the MSD curve (slope 4 D per sub-step) and the
    spread of the end point."""
    g = np.load(os.path.join(golden_dir, "trajectory_stats.npz"))
    lags = g["lags"]
    for i, D in enumerate(g["Ds"]):
        key = "D%g" % D
        traj, Dd = brownian_oracle(600, 300, [float(D)], [1e-12], 1.0, seed=40 + i)
        assert np.allclose(Dd, D, rtol=1e-4)
        steps = np.diff(traj, axis=1)
        assert abs(steps.std() / float(g[key + "/step_std"]) - 1) < 0.01, key
        assert _ks(steps / np.sqrt(2 * D), g[key + "/step_quantiles"]) < 0.01, key
        msd = np.array([((traj[:, l:] - traj[:, :-l]) ** 2).sum(-1).mean() for l in lags])
        assert np.abs(msd / g[key + "/msd"] - 1).max() < 0.12, key                     # 600 trajectories: ~6 % noise at lag 30
        assert abs(np.polyfit(lags, msd, 1)[0] / (4 * D) - 1) < 0.08, key
        assert abs(traj[:, -1].std() / float(g[key + "/end_std"]) - 1) < 0.08, key


@pytest.mark.skipif(not refshim.reference_available(), reason="reference checkout not present")
def test_against_the_live_reference_brownian_motion():
    """Same contract against the reference function itself (run here, where /root/reference exists)."""
    gen, _ = refshim.import_reference()
    np.random.seed(5)
    ref = gen.brownian_motion(800, 30, 10, 5.0, 10.0, startAtZero=True)
    mine, _ = brownian_oracle(800, 300, [5.0], [1e-12], 1.0, seed=77)
    rs, ms = np.diff(ref, axis=1), np.diff(mine, axis=1)
    assert abs(rs.std() / ms.std() - 1) < 0.01
    qs = np.quantile(rs.ravel(), np.linspace(0, 1, 401))
    assert _ks(ms, qs) < 0.01
    assert abs(ref[:, -1].std() / mine[:, -1].std() - 1) < 0.08
    assert np.all(ref[:, 0] == 0) and np.all(mine[:, 0] == 0)          # startAtZero: the first position is the origin


def test_step_statistics_match_fixtures(golden_dir):
    traj, D = brownian_oracle(400, 300, [7.0], [1e-12], 1.0, seed=3)
    steps = np.diff(traj, axis=1)
    assert abs(steps.std() - np.sqrt(14.0)) / np.sqrt(14.0) < 0.01
    assert np.all(traj[:, 0] == 0)
    # the reference's own validation fixtures (val7: D = 7 exactly) have the same step std
    fix = np.load(os.path.join(golden_dir, "render_inputs.npz"))["traj30"] * 100.0
    assert abs(np.diff(fix, axis=1).std() - steps.std()) / steps.std() < 0.03
    # MSD at lag 1 recovers D: <dr^2> = 4 D
    assert abs((steps ** 2).sum(-1).mean() / 4.0 - 7.0) < 0.1


def test_d_groups_and_truncation():
    gm = [1, 3, 5, 7, 9, 10.2]
    traj, D = brownian_oracle(1200, 20, gm, [1.0] * 6, 100.0, seed=11)
    assert np.all(D > 0)
    for g in range(6):
        assert abs(D[g::6].mean() - gm[g]) < (0.2 if g else 0.35)   # group 0 is truncated at 0 -> biased up
    t2, D2 = brownian_oracle(5, 20, gm, [1.0] * 6, 100.0, seed=11, seq_offset=600)
    assert np.array_equal(t2, traj[600:605]) and np.array_equal(D2, D[600:605])   # keyed by global id
