"""CPU tests (-m "not gpu"): the plain-PyTorch ViT oracle against goldens produced by
the unmodified reference nn.Module (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import vit_oracle as vo

from vit_cases import CASES, PARAM_COUNTS, load_case  # noqa: E402


@pytest.mark.parametrize("name", list(CASES))
def test_forward_loss_grads(golden_dir, name):
    z, sd, x, tgt, feats = load_case(golden_dir, name)
    if name in PARAM_COUNTS:
        assert sum(v.numel() for k, v in sd.items() if not vo.is_buffer(k)) == PARAM_COUNTS[name]
    pred, loss, g, _ = vo.loss_and_grads(sd, CASES[name], x, tgt, feats)
    assert np.abs(pred.numpy() - z["pred"]).max() < 2e-6
    assert abs(float(loss) - float(z["loss"])) < 1e-6
    gmax = max(float(z[k]) for k in z.files if k.startswith("gradnorm/"))
    for k in g:
        assert abs(float(g[k].double().norm()) - float(z["gradnorm/" + k])) < 1e-4 * gmax + 1e-4 * float(z["gradnorm/" + k])
        if "grad/" + k in z.files:
            ref = z["grad/" + k]
            # the oracle writes BatchNorm out by hand (the reference calls nn.BatchNorm2d): fp32 summation-order noise of up to 4e-4
            # relative on the small feature-path gradients of the DeepResNet + features models
            rtol = 1e-3 if name.startswith("deepcnn_n_feat") else 1e-5
            assert np.abs(g[k].numpy() - ref).max() <= rtol * max(np.abs(ref).max(), 1e-3 * gmax)


@pytest.mark.parametrize("name", ["deepcnn_n", "linear_s_feat_late"])
def test_train_step(golden_dir, name):
    z, sd, x, tgt, feats = load_case(golden_dir, name)
    st = vo.new_opt_state(sd)
    loss = vo.train_step(sd, st, CASES[name], x, tgt, feats)
    assert abs(loss - float(z["loss"])) < 1e-6
    gmax = max(float(z[k]) for k in z.files if k.startswith("gradnorm/"))
    for k, v in sd.items():
        ref_sum, ref_abs = float(z["after_sum/" + k]), float(z["after_abs/" + k])
        tol = 1e-5 * max(ref_abs, 1.0)
        if "gradnorm/" + k in z.files and float(z["gradnorm/" + k]) < 1e-6 * gmax:
            # mathematically-zero gradient (k_proj.bias: softmax is shift invariant): Adam turns
            # rounding noise into +-lr per element, so only bound the update size
            tol += 2 * 1e-4 * v.numel()
        assert abs(float(v.double().sum()) - ref_sum) < tol, k
        assert abs(float(v.double().abs().sum()) - ref_abs) < tol, k


from vit_cases import MODULAR_CASES  # noqa: E402


@pytest.mark.parametrize("name", list(MODULAR_CASES))
def test_modular_forward_loss_grads(golden_dir, name):
    """ModularTransformer restatement (oracle.vit_oracle.forward_modular) against the reference nn.Module's outputs."""
    z, sd, x, tgt, feats = load_case(golden_dir, name)
    pred, loss, g, _ = vo.loss_and_grads(sd, MODULAR_CASES[name], x, tgt, feats)
    assert np.abs(pred.numpy() - z["pred"]).max() < 2e-6
    assert abs(float(loss) - float(z["loss"])) < 1e-6
    gmax = max(float(z[k]) for k in z.files if k.startswith("gradnorm/"))
    assert set(g) == set(k[9:] for k in z.files if k.startswith("gradnorm/"))
    for k in g:
        assert abs(float(g[k].double().norm()) - float(z["gradnorm/" + k])) < 1e-4 * gmax + 1e-4 * float(z["gradnorm/" + k])
        if "grad/" + k in z.files:
            ref = z["grad/" + k]
            assert np.abs(g[k].numpy() - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-3 * gmax)
