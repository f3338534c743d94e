"""Per-parameter gradient error of the CUDA DeepResNet-ViT vs the fp32 oracle as a function of batch size."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from moleculardiffusion_mivit_b200 import models as M
from oracle import vit_oracle as vo

P, Fr, E, H, HD, L = 7, 20, 32, 2, 64, 3
for B in (2, 8, 32):
    torch.manual_seed(1)
    model = M.GeneralTransformer(M.DeepResNetEmbedding, {"patch_size": P, "embed_dim": E}, E, H, HD, L, M.MLPHead, F.relu, 0.0, True, True, True)
    sd = {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}
    cfg = dict(embedding="deepresnet", embed_dim=E, num_heads=H, num_layers=L, activation="relu", use_pos_encoding=True, use_regression_token=True)
    g = torch.Generator().manual_seed(17)
    x = 0.1 + 0.25 * torch.randn((B, Fr, P, P), generator=g).abs()
    tgt = torch.rand((B, 1), generator=g)
    _, _, ref_g, _ = vo.loss_and_grads(sd, cfg, x, tgt, None)
    for impl in (1, 0):
        m2 = M.GeneralTransformer(M.DeepResNetEmbedding, {"patch_size": P, "embed_dim": E}, E, H, HD, L, M.MLPHead, F.relu, 0.0, True, True, True)
        m2.load_state_dict(sd)
        m2.cuda().train()
        m2.conv_impl = impl
        pred = m2(x.cuda())
        F.mse_loss(pred, tgt.cuda()).backward()
        errs = []
        for k, p in m2.named_parameters():
            if k.startswith("embedding") and float(ref_g[k].norm()) > 0:
                errs.append((k, float((p.grad.cpu() - ref_g[k]).norm() / ref_g[k].norm())))
        print("B=%d impl=%d " % (B, impl) + " ".join("%s=%.3f" % (k.replace("embedding.", "").replace("res_block", "rb").replace(".weight", ".w").replace(".bias", ".b"), e) for k, e in errs))
