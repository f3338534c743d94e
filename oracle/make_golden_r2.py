"""TEST INFRASTRUCTURE ONLY -- round-2 goldens from the UNMODIFIED reference (imported through oracle/refshim.py).
Run in the build container:

    python -m oracle.make_golden_r2

Writes under tests/golden/:
  render_noise_stats_p13.npz  moments / per-pixel maps / quantiles of the reference's own np.random path for the Framerate
                              experiment's P=13 image_props at n = 10 (what bench.py renders; trainSettingsFramerate.py:62-81)
  trajectory_stats.npz        statistics of the reference's in-repo Brownian generator (helpers/helpersGeneration.py:9-45):
                              step std, step quantiles, MSD curve, for the D groups of the training loops
  vit_deepcnn_n_feat_{early,late}.npz   ImagesFeatures models (trainSettingsImagesFeatures.py:112-188): DeepResNet ViT + 25 features
  vit_linear_s_leaky.npz      F.leaky_relu feed-forward (tests/train_tests/trainSettings.py:89)
  vit_mod_perframe.npz        ModularTransformer(use_regression_token=False, single_prediction=False): per-frame outputs (:585-593)
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from . import refshim
from .make_golden import C3_PROPS, _Deterministic, _quiet

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

FRAMERATE_PROPS = dict(C3_PROPS, output_size=13)


def vit_record(model, x, tgt, feats=None, call=None):
    model.train()
    sd0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    out = call(model) if call is not None else (model(x) if feats is None else model(x, feats))
    loss = F.mse_loss(out, tgt)
    loss.backward()
    rec = {"pred": out.detach().numpy(), "loss": np.float64(loss.item()), "x": x.numpy(), "target": tgt.numpy()}
    if feats is not None:
        rec["features"] = feats.numpy()
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        rec["gradnorm/" + k] = np.float64(p.grad.double().norm().item())
        if p.numel() <= 4096:
            rec["grad/" + k] = p.grad.numpy().copy()
    for k, a in sd0.items():
        rec["sd/" + k] = a
    return rec, float(loss.item())


def main():
    gen, models = refshim.import_reference()
    inp = np.load(os.path.join(OUT, "render_inputs.npz"))
    traj30 = inp["traj30"]

    # ---- Framerate P=13 noisy statistics (the reference's real np.random path)
    np.random.seed(20261019)
    R = 24
    vals = np.stack([_quiet(gen.trajectories_to_video, traj30.copy(), 10, True, FRAMERATE_PROPS) for _ in range(R)])  # (R,8,30,13,13)
    q = np.linspace(0, 1, 2001)
    np.savez_compressed(os.path.join(OUT, "render_noise_stats_p13.npz"),
                        mean=vals.mean(dtype=np.float64), std=vals.std(dtype=np.float64),
                        quantiles=np.quantile(vals.astype(np.float64).ravel(), q),
                        pix_mean=vals.mean(axis=(0, 1, 2), dtype=np.float64), pix_std=vals.std(axis=(0, 1, 2), dtype=np.float64),
                        repeats=R)
    print("p13 stats", vals.mean(), vals.std())

    # ---- the reference's in-repo Brownian generator
    np.random.seed(20261020)
    rec = {}
    Ds = [1.0, 3.0, 5.0, 7.0, 9.0, 10.2]
    qs = np.linspace(0, 1, 401)
    lags = np.arange(1, 31)
    for D in Ds:
        tr = gen.brownian_motion(600, 30, 10, D, 10.0, startAtZero=True)      # sigma^2 = 2 D dt / nposframe = 2 D per sub-step
        steps = np.diff(tr, axis=1)
        msd = np.array([((tr[:, l:] - tr[:, :-l]) ** 2).sum(-1).mean() for l in lags])
        rec["D%g/step_std" % D] = steps.std()
        rec["D%g/step_quantiles" % D] = np.quantile(steps.ravel() / np.sqrt(2 * D), qs)
        rec["D%g/msd" % D] = msd
        rec["D%g/end_std" % D] = tr[:, -1].std()
    rec["lags"] = lags
    rec["Ds"] = np.array(Ds)
    np.savez_compressed(os.path.join(OUT, "trajectory_stats.npz"), **rec)
    print("trajectory stats", {k: float(v) for k, v in rec.items() if k.endswith("step_std")})

    # ---- ViT goldens
    z = np.load(os.path.join(OUT, "vit_deepcnn_n.npz"))
    x, tgt = torch.tensor(z["x"]), torch.tensor(z["target"])
    feats = torch.randn(4, 25, generator=torch.Generator().manual_seed(5))
    for fusion in ("early", "late"):
        torch.manual_seed(11)
        m = models.GeneralTransformer(models.DeepResNetEmbedding, {"patch_size": 9, "embed_dim": 64}, 64, 4, 128, 6, models.MLPHead,
                                      F.relu, 0.0, False, True, True, True, fusion, 25)
        r, loss = vit_record(m, x, tgt, feats)
        np.savez_compressed(os.path.join(OUT, "vit_deepcnn_n_feat_%s.npz" % fusion), **r)
        print("deepcnn_n_feat_" + fusion, sum(p.numel() for p in m.parameters()), loss)
    torch.manual_seed(12)
    m = models.GeneralTransformer(models.LinearProjectionEmbedding, {"patch_size": 9, "embed_dim": 32}, 32, 2, 64, 3, models.MLPHead,
                                  F.leaky_relu, 0.0, True, True, True)
    r, loss = vit_record(m, x, tgt)
    np.savez_compressed(os.path.join(OUT, "vit_linear_s_leaky.npz"), **r)
    print("linear_s_leaky", loss)
    torch.manual_seed(13)
    m = models.ModularTransformer(32, 2, 64, 2, models.MLPHead(input_dim=32), F.relu, 0.0, True, False, False, "images_only",
                                  models.LinearProjectionEmbedding, {"patch_size": 9, "embed_dim": 32})
    tgt_pf = torch.rand(4, 30, 1, generator=torch.Generator().manual_seed(14))
    r, loss = vit_record(m, x, tgt_pf, call=lambda mm: mm(x, None))
    assert r["pred"].shape == (4, 30, 1)
    np.savez_compressed(os.path.join(OUT, "vit_mod_perframe.npz"), **r)
    print("mod_perframe", loss)


if __name__ == "__main__":
    sys.exit(main())
