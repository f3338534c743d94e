"""TEST INFRASTRUCTURE ONLY -- CNN-baseline goldens (tests/golden/resnet_*.npz) from the UNMODIFIED reference nn.Modules
(helpers/models.py:600-772, imported through oracle/refshim.py).  Run in the build container:

    python -m oracle.make_golden_resnet

Images: the 4 noise-free P = 9 sequences of tests/golden/vit_deepcnn_n.npz (rendered by the reference) and a P = 13 crop-free
variant made of seeded normalised-image-like noise; external features: seeded normals."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from . import refshim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def record(model, x, tgt, ext=None):
    model.train()
    sd0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    opt.zero_grad()
    out = model(x) if ext is None else model(x, ext)
    loss = F.mse_loss(out, tgt)
    loss.backward()
    rec = {"pred": out.detach().numpy(), "loss": np.float64(loss.item()), "x": x.numpy(), "target": tgt.numpy()}
    if ext is not None:
        rec["ext"] = ext.numpy()
    for k, p in model.named_parameters():
        rec["gradnorm/" + k] = np.float64(p.grad.double().norm().item())
        if p.numel() <= 4096:
            rec["grad/" + k] = p.grad.numpy().copy()
    opt.step()
    for k, p in model.state_dict().items():
        rec["after_sum/" + k] = np.float64(p.double().sum().item())
        rec["after_abs/" + k] = np.float64(p.double().abs().sum().item())
    for k, a in sd0.items():
        rec["sd/" + k] = a
    return rec, float(loss.item()), sum(p.numel() for p in model.parameters())


def main():
    _, models = refshim.import_reference()
    z = np.load(os.path.join(OUT, "vit_deepcnn_n.npz"))
    x9, tgt = torch.tensor(z["x"]), torch.tensor(z["target"])
    g = torch.Generator().manual_seed(31)
    x13 = 0.1 + 0.25 * torch.randn((3, 12, 13, 13), generator=g).abs()
    tgt13 = torch.rand((3, 1), generator=g)
    ext = torch.randn(4, 25, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(21)
    r, loss, n = record(models.MultiImageResNet(9), x9, tgt)
    np.savez_compressed(os.path.join(OUT, "resnet_p9.npz"), **r)
    print("resnet_p9", n, loss)
    torch.manual_seed(22)
    r, loss, n = record(models.MultiImageResNet(13), x13, tgt13)
    np.savez_compressed(os.path.join(OUT, "resnet_p13.npz"), **r)
    print("resnet_p13", n, loss)
    torch.manual_seed(23)
    tgt_pf = torch.rand((4, 30, 1), generator=torch.Generator().manual_seed(6))
    r, loss, n = record(models.MultiImageResNet(9, single_prediction=False), x9, tgt_pf)
    np.savez_compressed(os.path.join(OUT, "resnet_p9_perframe.npz"), **r)
    print("resnet_p9_perframe", n, loss)
    torch.manual_seed(24)
    r, loss, n = record(models.MultiImageFeatureResNet(9, 25, feature_size=64, hidden_size=128), x9, tgt, ext)
    np.savez_compressed(os.path.join(OUT, "resnet_ft_p9.npz"), **r)
    print("resnet_ft_p9", n, loss)


if __name__ == "__main__":
    sys.exit(main())
