"""CPU tests: statistical contract of the trajectory source restatement (SURVEY.md section 4:
fixtures' step std = sqrt(2D); MSD-recovered D)."""
import os

import numpy as np

from oracle.trajectory_oracle import brownian_oracle


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
