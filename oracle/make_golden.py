"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED
reference at /root/reference (imported through oracle/refshim.py).  Run in the build
container (the GPU box has no /root/reference):

    python -m oracle.make_golden

Inputs are a few trajectories of the reference's own validation fixtures
(Experiments/validation_trajectories/{20,30}/val*.npy); outputs are what the
reference's functions return for them.  Deterministic goldens patch np.random.normal
to return its mean and np.random.poisson to return lam (no reference file is edited);
noisy goldens use the reference's real np.random path under np.random.seed and keep
only summary statistics (moments, quantiles, per-pixel maps).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from . import refshim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

C3_PROPS = {  # Experiments/Embeddings/trainSettingsEmbeddings.py image_props (P=9)
    "particle_intensity": [6000 - 1420, 500], "NA": 1.46, "wavelength": 500e-9, "psf_division_factor": 1.3,
    "resolution": 100e-9, "output_size": 9, "upsampling_factor": 5, "background_intensity": [1420, 290],
    "poisson_noise": 100, "trajectory_unit": 1200,
}


class _Deterministic:
    """np.random.normal -> loc, np.random.poisson -> lam."""

    def __enter__(self):
        self._n, self._p = np.random.normal, np.random.poisson
        np.random.normal = lambda loc=0.0, scale=1.0, size=None: (
            np.full(size, loc, dtype=np.float64) if size is not None else float(loc))
        np.random.poisson = lambda lam=1.0, size=None: np.asarray(lam)
        return self

    def __exit__(self, *a):
        np.random.normal, np.random.poisson = self._n, self._p


def _quiet(fn, *a, **k):
    """The reference prints 'Particle Left the image' per sub-position; silence it."""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    os.makedirs(OUT, exist_ok=True)
    gen, models = refshim.import_reference()
    ps = refshim.import_experiment_settings("PSFNoise")
    fr = refshim.import_experiment_settings("Framerate")
    vt = os.path.join(refshim.REFERENCE_ROOT, "Experiments", "validation_trajectories")
    traj30 = np.load(os.path.join(vt, "30", "val7.npy"))[:8] / 100.0     # as load_validation_data does (/traj_div_factor)
    traj30b = np.load(os.path.join(vt, "30", "val3.npy"))[:2] / 100.0
    traj20 = np.load(os.path.join(vt, "20", "val5.npy"))[:2] / 100.0
    np.savez_compressed(os.path.join(OUT, "render_inputs.npz"), traj30=traj30, traj30b=traj30b, traj20=traj20)

    clean = dict(C3_PROPS)
    clean["background_intensity"] = [0, 0]
    clean["poisson_noise"] = -1
    g = {}
    with _Deterministic():
        t = traj30.copy()
        g["v1_p9_center"] = _quiet(gen.trajectories_to_video, t, 10, True, clean)
        g["v1_flipped_input_y"] = t[:, :4, 1].copy()      # pins the in-place y flip side effect (:197)
        g["v1_p9_nocenter"] = _quiet(gen.trajectories_to_video, traj30[:2].copy(), 10, False, clean)
        p13 = dict(clean); p13["output_size"] = 13
        g["v1_p13_center"] = _quiet(gen.trajectories_to_video, traj30[:2].copy(), 10, True, p13)
        p8 = dict(clean); p8["output_size"] = 8           # even grid: non-unit linspace step (:90-91)
        g["v1_p8_center"] = _quiet(gen.trajectories_to_video, traj30[:1].copy(), 10, True, p8)
        p7 = dict(clean); p7["output_size"] = 7; p7["upsampling_factor"] = 10
        g["v1_p7u10_n15"] = _quiet(gen.trajectories_to_video, traj30[:1].copy(), 15, True, p7)
        bgp = dict(C3_PROPS); bgp["poisson_noise"] = -1   # mean background, no Poisson
        v = _quiet(gen.trajectories_to_video, traj30[:2].copy(), 10, True, bgp)
        g["v1_p9_bgmean_norm"] = gen.normalize_images(v, 1420, 290, 6000)[0]
        g["psfnoise_mean"] = _quiet(ps.trajs_to_vid_psf_noise, traj20[:1].copy(), 10, center=True, image_props=ps.image_props,
                                    PSF_Settings=ps.PSF_Settings, Noise_Settings=ps.Noise_Settings)
        tt = traj30b[:1].copy()
        g["framerate_mean"] = _quiet(fr.trajs_to_vid_framerates, tt, fr.nPosPerFrame, center=True,
                                     image_props=fr.image_props).numpy()
        assert np.array_equal(tt, traj30b[:1])            # six flips restore the caller's array
    np.savez_compressed(os.path.join(OUT, "render_golden.npz"), **g)

    # ---- noisy statistics from the reference's own np.random path
    np.random.seed(20261018)
    R = 40
    vals = np.stack([_quiet(gen.trajectories_to_video, traj30.copy(), 10, True, C3_PROPS) for _ in range(R)])  # (R,8,30,9,9)
    q = np.linspace(0, 1, 2001)
    s = {"v1_mean": vals.mean(dtype=np.float64), "v1_std": vals.std(dtype=np.float64),
         "v1_quantiles": np.quantile(vals.astype(np.float64).ravel(), q),
         "v1_pix_mean": vals.mean(axis=(0, 1, 2), dtype=np.float64), "v1_pix_std": vals.std(axis=(0, 1, 2), dtype=np.float64),
         "v1_repeats": R}
    Rp = 12
    pv = np.stack([_quiet(ps.trajs_to_vid_psf_noise, traj20.copy(), 10, center=True, image_props=ps.image_props,
                          PSF_Settings=ps.PSF_Settings, Noise_Settings=ps.Noise_Settings) for _ in range(Rp)])  # (R,2,5,6,20,9,9)
    s["psf_mean"] = pv.mean(axis=(0, 1, 4, 5, 6), dtype=np.float64)
    s["psf_std"] = pv.std(axis=(0, 1, 4, 5, 6), dtype=np.float64)
    qq = np.linspace(0, 1, 501)
    s["psf_quantiles"] = np.stack([[np.quantile(pv[:, :, i, j].astype(np.float64).ravel(), qq) for j in range(pv.shape[3])]
                                   for i in range(pv.shape[2])])
    s["psf_repeats"] = Rp
    np.savez_compressed(os.path.join(OUT, "render_noise_stats.npz"), **s)

    # ---- ViT goldens (reference nn.Module, shared random init)
    with _Deterministic():
        v = _quiet(gen.trajectories_to_video, traj30[:4].copy(), 10, True, clean)
    x = torch.tensor(gen.normalize_images(v, 1420, 290, 6000)[0])
    tgt = torch.tensor([[0.1], [0.3], [0.7], [0.9]])

    def run(name, model, feats=None, save_state=True):
        model.train()
        sd0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        opt.zero_grad()
        out = model(x) if feats is None else model(x, feats)
        loss = F.mse_loss(out, tgt)
        loss.backward()
        rec = {"pred": out.detach().numpy(), "loss": np.float64(loss.item()), "x": x.numpy(), "target": tgt.numpy()}
        if feats is not None:
            rec["features"] = feats.numpy()
        for k, p in model.named_parameters():
            rec["gradnorm/" + k] = np.float64(p.grad.double().norm().item())
            if p.numel() <= 4096:
                rec["grad/" + k] = p.grad.numpy().copy()
        opt.step()
        for k, p in model.state_dict().items():
            rec["after_sum/" + k] = np.float64(p.double().sum().item())
            rec["after_abs/" + k] = np.float64(p.double().abs().sum().item())
        if save_state:
            for k, a in sd0.items():
                rec["sd/" + k] = a
        np.savez_compressed(os.path.join(OUT, "vit_%s.npz" % name), **rec)
        print(name, "params", sum(p.numel() for p in model.parameters()), "loss", loss.item())

    torch.manual_seed(0)
    run("deepcnn_n", models.GeneralTransformer(models.DeepResNetEmbedding, {"patch_size": 9, "embed_dim": 64}, 64, 4, 128, 6,
                                               models.MLPHead, F.relu, 0.0, False, True, True))
    torch.manual_seed(1)
    run("linear_s_pos", models.GeneralTransformer(models.LinearProjectionEmbedding, {"patch_size": 9, "embed_dim": 32}, 32, 2, 64, 3,
                                                  models.MLPHead, F.relu, 0.0, True, True, True))
    torch.manual_seed(2)
    run("cnn_s_mean", models.GeneralTransformer(models.CNNEmbedding, {"patch_size": 9, "embed_dim": 32}, 32, 2, 64, 3,
                                                models.MLPHead, F.gelu, 0.0, True, False, True))
    feats = torch.randn(4, 25, generator=torch.Generator().manual_seed(5))
    for fusion in ("early", "late"):
        torch.manual_seed(3)
        run("linear_s_feat_" + fusion,
            models.GeneralTransformer(models.LinearProjectionEmbedding, {"patch_size": 9, "embed_dim": 32}, 32, 2, 64, 3,
                                      models.MLPHead, F.relu, 0.0, False, True, True, True, fusion, 25), feats)
    sz = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden bytes", sz)


if __name__ == "__main__":
    sys.exit(main())
