"""Per-parameter gradient / prediction error of the CUDA DeepResNet-ViT vs the fp32 oracle at product shapes (P=13, F=30,
deepcnn_n), on frames rendered by the CUDA renderer from Brownian trajectories (what bench.py trains on).
    python scripts/diag_parity.py [B ...]"""
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from moleculardiffusion_mivit_b200 import models as M
from moleculardiffusion_mivit_b200.helpersGeneration import brownian_motion, derive_render_params, render_device
from oracle import vit_oracle as vo

P, Fr, E, H, HD, L = 13, 30, 64, 4, 128, 6
PROPS = {"particle_intensity": [4580, 500], "NA": 1.46, "wavelength": 500e-9, "psf_division_factor": 1.3, "resolution": 100e-9,
         "output_size": P, "upsampling_factor": 5, "background_intensity": [1420, 290], "poisson_noise": 100, "trajectory_unit": 1200}


def frames(B, seed=3):
    traj = brownian_motion(B, Fr, 10, [1, 3, 5, 7, 9, 10.2], 1.0, seed=seed, D_var=1.0, div=100.0, return_device=True, return_D=True)
    prm = derive_render_params(PROPS, 10, True)
    prm.normalize, prm.norm_sub, prm.norm_div = 1, 1420.0 - 290.0, 6000.0 - 1130.0
    return render_device(traj[0], prm, seed=seed), (traj[1] / 10.0).view(-1, 1)


for B in [int(a) for a in sys.argv[1:]] or [32, 256]:
    torch.manual_seed(1)
    model = M.GeneralTransformer(M.DeepResNetEmbedding, {"patch_size": P, "embed_dim": E}, E, H, HD, L, M.MLPHead, F.relu, 0.0, False, True, True)
    sd = {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}
    cfg = dict(embedding="deepresnet", embed_dim=E, num_heads=H, num_layers=L, activation="relu", use_pos_encoding=False, use_regression_token=True)
    x, tgt = frames(B)
    t0 = time.time()
    ref_pred, ref_loss, ref_g, _ = vo.loss_and_grads(sd, cfg, x.cpu(), tgt.cpu(), None)
    t1 = time.time()
    model.cuda().train()
    pred = model(x)
    loss = F.mse_loss(pred, tgt)
    loss.backward()
    print("B=%d oracle %.1fs  loss %.6f (oracle %.6f)  pred max err %.2e" % (B, t1 - t0, loss.item(), float(ref_loss), (pred.cpu() - ref_pred).abs().max().item()))
    rows = []
    for k, p in model.named_parameters():
        r = ref_g[k]
        if float(r.norm()) > 0:
            rows.append((float((p.grad.cpu() - r).norm() / r.norm()), k, float(r.norm())))
    rows.sort(reverse=True)
    for e, k, n in rows[:14]:
        print("   %-55s rel %.4f  |g| %.3e" % (k, e, n))
    emb = [e for e, k, n in rows if k.startswith("embedding.")]
    rest = [e for e, k, n in rows if not k.startswith("embedding.")]
    print("   worst embedding %.4f, worst transformer/head %.4f, median all %.4f" % (max(emb), max(rest), float(np.median([e for e, _, _ in rows]))))
