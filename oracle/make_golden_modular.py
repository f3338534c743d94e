"""TEST INFRASTRUCTURE ONLY -- ModularTransformer goldens (tests/golden/vit_mod_*.npz) from the UNMODIFIED reference
nn.Module (helpers/models.py:366-593, imported through oracle/refshim.py).  Run in the build container:

    python -m oracle.make_golden_modular

Images are the 4 noise-free sequences of tests/golden/vit_deepcnn_n.npz (themselves rendered by the reference); per-frame
features are seeded normals with a few NaNs (the reference zeroes them with torch.nan_to_num)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from . import refshim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from vit_cases import MODULAR_CASES  # noqa: E402


def build_reference(models, c):
    emb = {"deepresnet": models.DeepResNetEmbedding, "linear": models.LinearProjectionEmbedding, "cnn": models.CNNEmbedding,
           None: None}[c.get("embedding")]
    act = {"relu": F.relu, "gelu": F.gelu}[c["activation"]]
    return models.ModularTransformer(
        c["embed_dim"], c["num_heads"], c["hidden_dim"], c["num_layers"], models.MLPHead(input_dim=c["embed_dim"]), act, 0.0,
        c["use_pos_encoding"], c["use_regression_token"], True, c["mode"], emb,
        {"patch_size": 9, "embed_dim": c["embed_dim"]} if emb is not None else None, c.get("features_dim"),
        c.get("feature_embedding_type", "linear"), c.get("fusion_method", "add"))


def main():
    _, models = refshim.import_reference()
    z = np.load(os.path.join(OUT, "vit_deepcnn_n.npz"))
    x, tgt = torch.tensor(z["x"]), torch.tensor(z["target"])
    for seed, (name, c) in enumerate(MODULAR_CASES.items()):
        torch.manual_seed(100 + seed)
        model = build_reference(models, c).train()
        feats = None
        if c["mode"] != "images_only":
            feats = torch.randn(x.shape[0], x.shape[1], c["features_dim"], generator=torch.Generator().manual_seed(7 + seed))
            feats[0, 3, 1] = float("nan")
            feats[2, 11, 0] = float("nan")
        imgs = x if c["mode"] != "features_only" else None
        sd0 = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
        out = model(imgs, feats)
        loss = F.mse_loss(out, tgt)
        loss.backward()
        rec = {"pred": out.detach().numpy(), "loss": np.float64(loss.item()), "target": tgt.numpy()}
        if imgs is not None:
            rec["x"] = imgs.numpy()
        if feats is not None:
            rec["features"] = feats.numpy()
        for k, p in model.named_parameters():
            rec["gradnorm/" + k] = np.float64(p.grad.double().norm().item())
            if p.numel() <= 4096:
                rec["grad/" + k] = p.grad.numpy().copy()
        for k, a in sd0.items():
            rec["sd/" + k] = a
        np.savez_compressed(os.path.join(OUT, "vit_%s.npz" % name), **rec)
        print(name, "params", sum(p.numel() for p in model.parameters()), "loss", loss.item())


if __name__ == "__main__":
    sys.exit(main())
