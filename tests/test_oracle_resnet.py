"""CPU tests (-m "not gpu"): the CNN-baseline oracle (oracle/resnet_oracle.py) against goldens produced by the unmodified
reference nn.Modules (helpers/models.py:600-772; oracle/make_golden_resnet.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import resnet_oracle as rn
from oracle.vit_oracle import adamw_update, is_buffer

CASES = {"resnet_p9": dict(single=True), "resnet_p13": dict(single=True), "resnet_p9_perframe": dict(single=False),
         "resnet_ft_p9": dict(single=True)}
PARAMS = {"resnet_p9": 315617, "resnet_p13": 315617, "resnet_p9_perframe": 315617, "resnet_ft_p9": 327201}


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = {k[3:]: torch.tensor(z[k]) for k in z.files if k.startswith("sd/")}
    ext = torch.tensor(z["ext"]) if "ext" in z.files else None
    return z, sd, torch.tensor(z["x"]), torch.tensor(z["target"]), ext


@pytest.mark.parametrize("name", list(CASES))
def test_forward_loss_grads_and_adamw_step(golden_dir, name):
    z, sd, x, tgt, ext = load(golden_dir, name)
    assert sum(v.numel() for k, v in sd.items() if not is_buffer(k)) == PARAMS[name]
    pred, loss, g, stats = rn.loss_and_grads(sd, x, tgt, ext, CASES[name]["single"])
    assert pred.shape == z["pred"].shape and np.abs(pred.numpy() - z["pred"]).max() < 2e-6
    assert abs(float(loss) - float(z["loss"])) < 1e-6
    gmax = max(float(z[k]) for k in z.files if k.startswith("gradnorm/"))
    assert set(g) == set(k[9:] for k in z.files if k.startswith("gradnorm/"))
    for k in g:
        assert abs(float(g[k].double().norm()) - float(z["gradnorm/" + k])) < 1e-4 * gmax + 1e-4 * float(z["gradnorm/" + k]), k
        if "grad/" + k in z.files:
            ref = z["grad/" + k]
            assert np.abs(g[k].numpy() - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1e-3 * gmax), k
    # one AdamW(lr=1e-4) step of the reference's optimizer + the BatchNorm running-statistics update
    new = dict(sd)
    for k, gr in g.items():
        new[k] = adamw_update(sd[k], gr, torch.zeros_like(gr), torch.zeros_like(gr), 1)[0]
    new.update(stats)
    for k, v in new.items():
        ref_sum, ref_abs = float(z["after_sum/" + k]), float(z["after_abs/" + k])
        tol = 1e-5 * max(ref_abs, 1.0) + (2e-4 * v.numel() if (not is_buffer(k) and float(z["gradnorm/" + k]) < 1e-6 * gmax) else 0)
        assert abs(float(v.double().sum()) - ref_sum) < tol, k
        assert abs(float(v.double().abs().sum()) - ref_abs) < tol, k
