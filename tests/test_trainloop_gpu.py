"""GPU test (-m gpu) of the experiment-loop mirror (moleculardiffusion_mivit_b200/trainloop.py) of
Experiments/PSFNoise/trainModelsPSFNoise.py:113-251: batch-size schedule, per-model training over the cycle's data,
StepLR per cycle, eval-mode validation losses in the reference's dict layout, results file interchange."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PSF, NOISE = [2, 1], [0, 1 / 20]
PROPS = {"particle_intensity": [5000, 500], "NA": 1.46, "wavelength": 500e-9, "psf_division_factor": 1.3, "resolution": 100e-9,
         "output_size": 9, "upsampling_factor": 5, "background_intensity": [5000, 0], "poisson_noise": 100, "trajectory_unit": 1200}


def make_prediction(model, name, images, eval=True):       # trainSettingsPSFNoise.py:164-172
    prefix, psf_index, noise_index = name.split("_")
    return model(images[:, int(psf_index), int(noise_index)])


def test_two_cycles_psfnoise_layout(golden_dir, tmp_path):
    import torch
    import torch.nn.functional as F
    from moleculardiffusion_mivit_b200 import models as M, experiments as X, trainloop as TL
    torch.manual_seed(0)
    names = ["tr_0_0", "tr_1_1"]
    models = {n: M.GeneralTransformer(M.LinearProjectionEmbedding, {"patch_size": 9, "embed_dim": 32}, 32, 2, 64, 2, M.MLPHead,
                                      F.relu, 0.0, False, True, True).cuda() for n in names}
    render = lambda t: X.trajs_to_vid_psf_noise(t, 10, center=True, image_props=PROPS, PSF_Settings=PSF, Noise_Settings=NOISE, seed=3)
    inp = np.load(os.path.join(golden_dir, "render_inputs.npz"))["traj30"]
    val = [(render(inp[:3].copy()), 1.0), (render(inp[3:6].copy()), 7.0)]
    assert val[0][0].shape == (3, 2, 2, 30, 9, 9)
    prefix = str(tmp_path / "training_results_PSFNoise")
    loop = TL.ExperimentLoop(models, render, make_prediction, val, T=300, N=6, TrainingDs_list=[[1, 1], [10.2, 1]],
                             adaptive_batch_size=1, seed=1, results_prefix=prefix)
    w0 = {n: m.state_dict()["mlp_head.mlp.3.weight"].clone() for n, m in models.items()}
    assert loop.batch_size == 1
    losses = loop.run(2)
    assert loop.batch_size == 2                                   # doubled at cycle 1 (adaptive_batch_size = 1)
    assert len(loop.all_gen_labels) == 2 * (6 + 3) and np.all(loop.all_gen_labels > 0)   # N and N // 2 for the 10.2 group
    for n in names:
        assert set(losses[n]) == {"val_1.0", "val_7.0", "val_avg"} and all(len(v) == 2 for v in losses[n].values())
        assert abs(losses[n]["val_avg"][-1] - 0.5 * (losses[n]["val_1.0"][-1] + losses[n]["val_7.0"][-1])) < 1e-6
        assert not torch.equal(w0[n], models[n].state_dict()["mlp_head.mlp.3.weight"])          # it trained
        assert loop.trainers[n].step_count == 9 + 5 and loop.trainers[n].epoch == 2              # ceil(9/1) + ceil(9/2) steps
    # validation loss = MSE(pred * D_max, D) in eval mode (:206-238)
    m = models["tr_1_1"].eval()
    with torch.no_grad():
        manual = F.mse_loss(m(torch.as_tensor(val[1][0]).float().cuda()[:, 1, 1]) * 10.0, torch.full((3, 1), 7.0, device="cuda")).item()
    assert abs(manual - losses["tr_1_1"]["val_7.0"][-1]) < 1e-5 * max(1.0, manual)
    # results files: last cycles (suffix = cycles remaining) and the final one; the reference's key layout
    for suffix in ("2", "1", ""):
        assert os.path.exists(prefix + suffix + ".pth")
    res = torch.load(prefix + ".pth", weights_only=False)
    assert set(res) == {"validation_losses", "all_labels", "model_weights"} and set(res["model_weights"]) == set(names)
    ref_keys = set(k[3:] for k in np.load(os.path.join(golden_dir, "vit_linear_s_feat_late.npz")).files if k.startswith("sd/"))
    mine = set(res["model_weights"]["tr_0_0"])
    assert {k for k in ref_keys if not k.startswith("feature_projector")} - {"transformer.encoder_layers.2.self_attn.q_proj.weight"} \
        >= {k for k in mine if "encoder_layers.2" not in k} - set()                                # same naming scheme
    assert "embedding.proj.weight" in mine and "mlp_head.mlp.3.bias" in mine and "reg_token" in mine


def test_loop_trains_a_foreign_cnn_baseline_beside_the_vit(golden_dir, tmp_path):
    """The reference experiments train a CNN baseline (MultiImageResNet, out of this package's scope) beside every ViT: any
    nn.Module that is not one of this package's transformers goes through the stock-PyTorch trainer of the same loop."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from moleculardiffusion_mivit_b200 import models as M, experiments as X, trainloop as TL
    torch.manual_seed(0)

    class TinyBaseline(nn.Module):      # stand-in with the baseline's call convention: [B, frames, P, P] -> [B, 1]
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(30, 4, 3, padding=1)
            self.fc = nn.Linear(4, 1)

        def forward(self, x):
            return self.fc(F.relu(self.conv(x)).mean(dim=(2, 3)))

    models = {"tr_0_0": M.GeneralTransformer(M.LinearProjectionEmbedding, {"patch_size": 9, "embed_dim": 32}, 32, 2, 64, 2, M.MLPHead,
                                             F.relu, 0.0, False, True, True).cuda(),
              "res_0_0": TinyBaseline()}
    render = lambda t: X.trajs_to_vid_psf_noise(t, 10, center=True, image_props=PROPS, PSF_Settings=PSF, Noise_Settings=NOISE, seed=3)
    inp = np.load(os.path.join(golden_dir, "render_inputs.npz"))["traj30"]
    val = [(render(inp[:3].copy()), 1.0)]
    loop = TL.ExperimentLoop(models, render, make_prediction, val, T=300, N=4, TrainingDs_list=[[1, 1]], adaptive_batch_size=-1,
                             seed=1, results_prefix=str(tmp_path / "res"))
    assert isinstance(loop.trainers["res_0_0"], TL._TorchTrainer) and not isinstance(loop.trainers["tr_0_0"], TL._TorchTrainer)
    w0 = models["res_0_0"].fc.weight.detach().clone()
    losses = loop.run(1, save=False)
    assert set(losses) == {"tr_0_0", "res_0_0"} and len(losses["res_0_0"]["val_avg"]) == 1
    assert not torch.equal(w0, models["res_0_0"].fc.weight.detach())


def test_cycles_draw_fresh_noise_and_features_travel_with_the_videos(golden_dir, tmp_path):
    """(1) render_fn is handed the global id of its first trajectory, so the noise streams differ between D groups and cycles even
    with a fixed seed (the reference draws fresh np.random noise every cycle); (2) the ImagesFeatures loop
    (trainModelsImagesFeatures.py:155-203): render_fn returns (videos, features), models whose name contains "ft" get the
    features, the others None; validation sets are (images, features, D) triples."""
    import torch
    import torch.nn.functional as F
    from moleculardiffusion_mivit_b200 import models as M, helpersGeneration as G, trainloop as TL
    props = {"particle_intensity": [4580, 500], "NA": 1.46, "wavelength": 500e-9, "psf_division_factor": 1.3, "resolution": 100e-9,
             "output_size": 9, "upsampling_factor": 5, "background_intensity": [1420, 290], "poisson_noise": 100, "trajectory_unit": 1200}
    seen = []

    def render(trajs, seq_offset=0, seed=None):
        vids = G.trajectories_to_video(trajs.copy(), 10, True, props, seed=seed, seq_offset=seq_offset, normalize=(1420, 290, 6000))
        feats = G.create_video_and_feature_pairs(trajs.copy(), 10, True, props, dt=1.0, seed=seed)[1]
        feats = np.nan_to_num(feats, nan=0.0, posinf=0.0, neginf=0.0)
        seen.append((seq_offset, vids.copy()))
        return vids, feats

    def make_prediction(model, name, images, features, *a, **k):          # trainSettingsImagesFeatures.py:303-341
        with torch.no_grad():
            return model(images, features if "ft" in name else None)

    torch.manual_seed(0)
    mk = lambda ft: M.GeneralTransformer(M.LinearProjectionEmbedding, {"patch_size": 9, "embed_dim": 32}, 32, 2, 64, 2, M.MLPHead, F.relu,
                                         0.0, False, True, True, ft, "late", 25 if ft else None).cuda()
    models = {"im_tr": mk(False), "im_ft_late_tr": mk(True)}
    inp = np.load(os.path.join(golden_dir, "render_inputs.npz"))["traj30"]
    v0, f0 = render(inp[:3].copy(), seq_offset=10 ** 6, seed=1)
    seen.clear()
    loop = TL.ExperimentLoop(models, render, make_prediction, [(v0, f0, 7.0)], T=300, N=4, TrainingDs_list=[[3, 1], [7, 1]],
                             adaptive_batch_size=-1, seed=1, results_prefix=str(tmp_path / "res"))
    w0 = models["im_ft_late_tr"].state_dict()["feature_projector.0.weight"].clone()
    losses = loop.run(2, save=False)
    assert [s for s, _ in seen] == [0, 4, 8, 12]                          # global ids: 2 groups x 2 cycles x N = 4
    assert not np.array_equal(seen[0][1], seen[2][1])                     # cycle 2 sees new trajectories AND new noise
    bg = [v[:, :, 0, 0] for _, v in seen]                                 # a corner pixel is background + noise only
    assert not np.allclose(bg[0], bg[2]) and not np.allclose(bg[0], bg[1])
    assert set(losses) == {"im_tr", "im_ft_late_tr"} and len(losses["im_tr"]["val_7.0"]) == 2
    assert not torch.equal(w0, models["im_ft_late_tr"].state_dict()["feature_projector.0.weight"])   # the features were used
