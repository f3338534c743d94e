"""GPU parity tests (-m gpu) of the CUDA CNN baselines (moleculardiffusion_mivit_b200/baselines.py -> csrc/resnet.cu through the
C ABI) against the fp32 oracle (oracle/resnet_oracle.py, itself pinned to the reference classes helpers/models.py:600-772 by the
goldens): state_dict keys, prediction, loss, every gradient, BatchNorm running statistics, the fused AdamW step, eval mode.
Everything is fp32 on both sides -> tolerances are summation-order noise (1e-4 relative)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import resnet_oracle as rn
from oracle import vit_oracle as vo
from test_oracle_resnet import CASES, load


def build(name):
    from moleculardiffusion_mivit_b200 import baselines as BL
    if name == "resnet_ft_p9":
        return BL.MultiImageFeatureResNet(9, 25, feature_size=64, hidden_size=128)
    return BL.MultiImageResNet(13 if "p13" in name else 9, single_prediction=CASES[name]["single"])


def relnorm(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("name", list(CASES))
def test_forward_loss_grads_match_oracle(golden_dir, name):
    import torch
    import torch.nn.functional as F
    z, sd, x, tgt, ext = load(golden_dir, name)
    ref_pred, ref_loss, ref_g, ref_stats = rn.loss_and_grads(sd, x, tgt, ext, CASES[name]["single"])
    gmax = max(float(v.norm()) for v in ref_g.values())
    # fp32 on both sides: summation-order noise (measured <= 1.1e-5).  The BatchNorm statistics are accumulated with atomics, so the
    # forward differs from run to run in the last bit, and in resnet_ft_p9 ONE pre-activation of layer2's 3x3 output sits within
    # that bit of zero: in ~25 % of the runs its ReLU mask differs from the CPU oracle's and the gradients that see that element
    # move (BatchNorm affine gradients of that layer, sums over few pixels, by 1.0e-3; scripts/diag_resnet_flaky.py -- poisoning
    # the workspace with NaN changes nothing, i.e. no uninitialised read).  Hence: EVERY run must be within the mask-flip bound
    # (1e-2, still far below what a wrong kernel produces), and one of up to five runs within summation-order noise.
    tight_fail = None
    for attempt in range(5):
        model = build(name)
        assert list(model.state_dict().keys()) == list(sd.keys())                 # reference state_dict keys, same order
        model.load_state_dict(sd)
        model.cuda().train()
        pred = model(x.cuda(), ext.cuda()) if ext is not None else model(x.cuda())
        assert tuple(pred.shape) == z["pred"].shape
        loss = F.mse_loss(pred, tgt.cuda())
        loss.backward()
        assert (pred.cpu() - ref_pred).abs().max().item() < 2e-5 * max(1.0, ref_pred.abs().max().item())
        assert abs(loss.item() - float(z["loss"])) < 1e-5 * max(1.0, float(z["loss"]))
        tight_fail = None
        for k, p in model.named_parameters():
            assert p.grad is not None, k
            r = ref_g[k]
            if float(r.norm()) < 1e-6 * gmax:
                assert float(p.grad.cpu().norm()) < 1e-4 * gmax, k
                continue
            e = relnorm(p.grad.cpu(), r)
            assert e < 1e-2, (k, e)
            if e >= 2e-4 and tight_fail is None:
                tight_fail = (k, e, attempt)
        if tight_fail is None:
            break
    assert tight_fail is None, tight_fail
    msd = model.state_dict()
    for k, v in ref_stats.items():
        assert torch.allclose(msd[k].cpu().float(), v.float(), rtol=1e-4, atol=1e-6), k


@pytest.mark.parametrize("name", ["resnet_p9", "resnet_ft_p9"])
def test_fused_train_step_and_eval_mode(golden_dir, name):
    import torch
    from moleculardiffusion_mivit_b200.baselines import CnnTrainer
    z, sd, x, tgt, ext = load(golden_dir, name)
    model = build(name)
    model.load_state_dict(sd)
    model.cuda().train()
    tr = CnnTrainer(model, lr=1e-4)
    loss = tr.train_step(x.cuda(), tgt.cuda(), ext.cuda() if ext is not None else None)
    assert abs(loss.item() - float(z["loss"])) < 1e-5 * max(1.0, float(z["loss"]))
    _, _, ref_g, ref_stats = rn.loss_and_grads(sd, x, tgt, ext, True)
    gmax = max(float(v.norm()) for v in ref_g.values())
    msd = model.state_dict()
    for k, gr in ref_g.items():
        want = vo.adamw_update(sd[k], gr, torch.zeros_like(gr), torch.zeros_like(gr), 1)[0]
        upd, upd_ref = msd[k].cpu() - sd[k], want - sd[k]
        assert (msd[k].cpu() - want).abs().max().item() <= 2.1e-4, k              # first AdamW step: |update| ~ lr
        if float(gr.norm()) > 1e-5 * gmax:
            differ = int((torch.sign(upd) != torch.sign(upd_ref)).sum())          # near-zero gradients may flip sign (fp32 order)
            assert differ <= max(2, 0.005 * upd.numel()), (k, differ, upd.numel())
    assert int(model.resnet.bn1.num_batches_tracked) == 1
    # eval mode: running statistics; the reference semantics for inference, and gradients are refused
    ref_sd = dict(sd)
    ref_sd.update(ref_stats)
    for k, gr in ref_g.items():
        ref_sd[k] = msd[k].cpu()
    model.eval()
    with torch.no_grad():
        pe = model(x.cuda(), ext.cuda()) if ext is not None else model(x.cuda())
    want = rn.forward(ref_sd, x, ext, True, training=False)
    assert (pe.cpu() - want).abs().max().item() < 1e-4 * max(1.0, want.abs().max().item())
    with pytest.raises(NotImplementedError, match="eval"):
        model(x.cuda(), ext.cuda()) if ext is not None else model(x.cuda())


@pytest.mark.parametrize("P,Fr,B", [(9, 30, 64), (7, 20, 5), (13, 30, 16), (15, 10, 3), (12, 6, 2)])
def test_shape_sweep_matches_oracle(P, Fr, B):
    """Patch sizes of the Embeddings sweep (7 / 9 / 13 / 15) and an even one, frame counts of the Framerate experiment, batch
    sizes up to several tiles of every kernel; weights: the mirror's own random init shared with the oracle."""
    import torch
    import torch.nn.functional as F
    from moleculardiffusion_mivit_b200 import baselines as BL
    torch.manual_seed(P * 100 + Fr)
    model = BL.MultiImageResNet(P)
    sd = {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    x = 0.1 + 0.25 * torch.randn((B, Fr, P, P), generator=g).abs()
    tgt = torch.rand((B, 1), generator=g)
    model.cuda().train()
    pred = model(x.cuda())
    F.mse_loss(pred, tgt.cuda()).backward()
    ref_pred, _, ref_g, _ = rn.loss_and_grads(sd, x, tgt)
    assert (pred.cpu() - ref_pred).abs().max().item() < 5e-5 * max(1.0, ref_pred.abs().max().item())
    gmax = max(float(v.norm()) for v in ref_g.values())
    for k, p in model.named_parameters():
        if float(ref_g[k].norm()) > 1e-6 * gmax:
            assert relnorm(p.grad.cpu(), ref_g[k]) < 1e-3, (k, relnorm(p.grad.cpu(), ref_g[k]))


def test_constructor_errors_and_experiment_loop(golden_dir, tmp_path):
    """Constructor surface of the reference + the experiment loop trains the CUDA baseline beside the ViT with the fused trainers
    (trainModelsPSFNoise.py:177-196: every model sees every batch of the cycle)."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from moleculardiffusion_mivit_b200 import baselines as BL, models as M, helpersGeneration as G, trainloop as TL
    with pytest.raises(NotImplementedError, match="nn.ReLU"):
        BL.MultiImageResNet(9, activation=nn.GELU)
    with pytest.raises(NotImplementedError, match="num_blocks"):
        BL.LightResNet(BL.BasicBlock, [2, 2, 2])
    assert sum(p.numel() for p in BL.MultiImageResNet(9).parameters()) == 315617        # train_resultsEmbeddings.ipynb cell 2
    props = {"particle_intensity": [4580, 500], "NA": 1.46, "wavelength": 500e-9, "psf_division_factor": 1.3, "resolution": 100e-9,
             "output_size": 9, "upsampling_factor": 5, "background_intensity": [1420, 290], "poisson_noise": 100, "trajectory_unit": 1200}

    def render(trajs, seq_offset=0, seed=None):
        return G.trajectories_to_video(trajs, 10, True, props, seed=seed, seq_offset=seq_offset, normalize=(1420, 290, 6000))

    def make_prediction(model, name, images, eval=True):
        return model(images)

    torch.manual_seed(0)
    models = {"tr": M.GeneralTransformer(M.LinearProjectionEmbedding, {"patch_size": 9, "embed_dim": 32}, 32, 2, 64, 2, M.MLPHead, F.relu,
                                         0.0, False, True, True).cuda(),
              "resnet": BL.MultiImageResNet(9)}
    inp = np.load(os.path.join(golden_dir, "render_inputs.npz"))["traj30"]
    val = [(render(inp[:3].copy(), seq_offset=10 ** 6, seed=1), 7.0)]
    loop = TL.ExperimentLoop(models, render, make_prediction, val, T=300, N=6, TrainingDs_list=[[3, 1], [7, 1]], adaptive_batch_size=-1,
                             seed=1, results_prefix=str(tmp_path / "res"))
    assert isinstance(loop.trainers["resnet"], BL.CnnTrainer)
    w0 = models["resnet"].state_dict()["resnet.fc2.weight"].clone()
    losses = loop.run(2)
    assert len(losses["resnet"]["val_7.0"]) == 2 and np.isfinite(losses["resnet"]["val_avg"]).all()
    assert not torch.equal(w0, models["resnet"].state_dict()["resnet.fc2.weight"])
    res = torch.load(str(tmp_path / "res") + ".pth", weights_only=False)
    assert "resnet.layer2.0.shortcut.1.running_var" in res["model_weights"]["resnet"]
