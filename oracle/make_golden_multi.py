"""TEST INFRASTRUCTURE ONLY -- golden outputs of the UNMODIFIED reference's trajectories_to_video_multiple_settings
(helpers/helpersGeneration.py:422-540) -> tests/golden/render_multi_golden.npz.  Run in the build container:

    python -m oracle.make_golden_multi

Deterministic: np.random.normal returns its mean, np.random.poisson returns lam (oracle/make_golden.py:_Deterministic), so
the four outputs are (signal, signal + bg_mean, the same, Gaussian-filtered).  skimage is not installed here (the reference pins
no version): oracle/refshim.py supplies skimage.filters.gaussian as the scipy.ndimage.gaussian_filter call it wraps -- the
parity of the FILTER output is therefore pinned to scipy's algorithm, not to a skimage binary.  A second, noisy run with the
reference's real np.random path keeps per-output moments only."""
import contextlib
import io
import os
import sys

import numpy as np

from . import refshim
from .make_golden import C3_PROPS, _Deterministic

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    gen, _ = refshim.import_reference()
    inp = np.load(os.path.join(OUT, "render_inputs.npz"))["traj30"]
    rec = {}
    for key, sl, n, center, over in (("p9_center", slice(0, 3), 10, True, {}),
                                     ("p13_nocenter_n15", slice(3, 4), 15, False, {"output_size": 13})):
        props = dict(C3_PROPS, **over)
        with _Deterministic(), contextlib.redirect_stdout(io.StringIO()):
            outs = gen.trajectories_to_video_multiple_settings(inp[sl].copy(), n, center, props)
        for name, a in zip(("none", "gauss", "poisson", "filter"), outs):
            assert a.dtype == np.float32
            rec["%s/%s" % (key, name)] = a
    # noisy statistics from the reference's own np.random path
    np.random.seed(20240518)
    R = 12
    with contextlib.redirect_stdout(io.StringIO()):
        runs = [gen.trajectories_to_video_multiple_settings(inp[:8].copy(), 10, True, dict(C3_PROPS)) for _ in range(R)]
    for i, name in enumerate(("none", "gauss", "poisson", "filter")):
        v = np.stack([r[i] for r in runs]).astype(np.float64)
        rec["noisy/%s_mean" % name] = v.mean()
        rec["noisy/%s_std" % name] = v.std()
        rec["noisy/%s_pixstd" % name] = v.std(axis=0).mean()       # run-to-run spread per pixel = the noise itself
    rec["noisy/repeats"] = R
    np.savez_compressed(os.path.join(OUT, "render_multi_golden.npz"), **rec)
    # trajs_to_vid_norm_rl (:635-658): normalised four outputs + Richardson-Lucy/TV estimates after iterations 2, 5, 10
    with _Deterministic(), contextlib.redirect_stdout(io.StringIO()):
        rl = gen.trajs_to_vid_norm_rl(inp[:2].copy(), 10, True, dict(C3_PROPS), [2, 5, 10])
    assert rl.shape == (2, 7, 30, 9, 9) and rl.dtype == np.float32
    np.savez_compressed(os.path.join(OUT, "render_norm_rl_golden.npz"), out=rl)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in rec.items()})


if __name__ == "__main__":
    sys.exit(main())
