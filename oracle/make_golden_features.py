"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/features_golden.npz from the UNMODIFIED reference
(helpers/helpersFeatures.py, helpers/helpersGeneration.py:48-74,663-719 imported through oracle/refshim.py).
Run in the build container (the GPU box has no /root/reference):

    python -m oracle.make_golden_features

Inputs: trajectories of the reference's validation fixtures (Experiments/validation_trajectories/30/val{1,5,9}.npy,
/100 as load_validation_data does); outputs: what compute_diffusion_features returns for their frame averages
(n = 10) at dt = 1 and dt = 0.1, for two full 300-point trajectories (the inputs of the two 25-vectors stored in the
reference's tests/models_tests/FeaturesTests.ipynb cell 2, copied below as NOTEBOOK_*), and what
create_video_and_feature_pairs returns as features / averaged trajectories."""
import os

import numpy as np

from . import refshim
from .make_golden import C3_PROPS, OUT, _Deterministic, _quiet

# tests/models_tests/FeaturesTests.ipynb cell 2 stored output (raw lines 52-65): val1[0]/100 and val7[0]/100, dt = 0.1
NOTEBOOK_VAL1 = [9.88111557e-01, 9.85268154e-04, 9.96463292e-01, -5.77043334e+00, 3.11840589e-03, 1.64823736e+00, 5.94923256e-01,
                 2.74116863e+00, -2.20270473e-04, 5.39640045e-01, 3.00000000e+02, 1.73477451e-02, 2.87091289e-02, -1.02081710e-05,
                 5.01683502e-01, 4.56375839e-01, 5.18697577e+00, 1.04797098e-03, 5.58786785e-02, 5.48307075e-02, 1.72899192e-02,
                 5.28062082e-01, 1.00000000e+00, 0.00000000e+00, 8.27287153e-02]
NOTEBOOK_VAL7 = [8.86014870e-01, 7.81538506e-03, 9.88494231e-01, -8.39785081e+00, 2.25351126e-04, 1.87461646e+00, 6.72732391e-01,
                 1.59469608e+00, 1.85775295e-03, 9.02085178e-01, 3.00000000e+02, 4.62057560e-02, 1.81901032e-01, 2.85436859e-05,
                 4.98316498e-01, 5.13422819e-01, 1.38155210e+01, 3.27160738e-03, 1.23138671e-01, 1.19867064e-01, 4.60517368e-02,
                 5.22973394e-01, 9.66555184e-01, 0.00000000e+00, 3.72783881e-01]


def main():
    gen, _ = refshim.import_reference()
    import importlib
    hf = importlib.import_module("helpers.helpersFeatures")
    vt = os.path.join(refshim.REFERENCE_ROOT, "Experiments", "validation_trajectories", "30")
    raw = np.concatenate([np.load(os.path.join(vt, "val%d.npy" % d))[:6] for d in (1, 5, 9)]) / 100.0       # (18, 300, 2)
    avg = gen.average_trajectories_frames(raw, 10)
    g = {"traj": raw, "avg": avg,
         "feat_dt1": np.stack([hf.compute_diffusion_features(t, dt=1.0) for t in avg]),
         "feat_dt01": np.stack([hf.compute_diffusion_features(t, dt=0.1) for t in avg])}
    long_in = np.stack([np.load(os.path.join(vt, "val1.npy"))[0], np.load(os.path.join(vt, "val7.npy"))[0]]) / 100.0
    g["long"] = long_in
    g["long_feat_dt01"] = np.stack([hf.compute_diffusion_features(t, dt=0.1) for t in long_in])
    g["notebook"] = np.array([NOTEBOOK_VAL1, NOTEBOOK_VAL7])
    assert np.allclose(g["long_feat_dt01"], g["notebook"], rtol=2e-6, atol=1e-9), "notebook vectors not reproduced"
    # the wrapper: features come from the y-FLIPPED trajectories (trajectories_to_video flips the caller's array in place)
    t = raw[:4].copy()
    with _Deterministic():
        vids, feats, (tr_out, tr_avg, tr_err) = _quiet(gen.create_video_and_feature_pairs, t, 10, True, dict(C3_PROPS), (0, 0), 0.1)
    g["pair_features"], g["pair_avg"], g["pair_traj_after"] = feats, tr_avg, t
    g["pair_video_mean"] = np.array([vids.mean(), vids.std()])
    np.savez_compressed(os.path.join(OUT, "features_golden.npz"), **g)
    print("features golden:", {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
