"""Per-parameter gradient deviation of the fused encoder kernels against the unfused path for one test case:
python scripts/diag_fused_encoder.py E H HD L P F B pos reg"""
import os
import sys
sys.path.insert(0, ".")
import torch
import torch.nn.functional as F
from moleculardiffusion_mivit_b200 import models as M

E, H, HD, Lyr, P, Fr, B = [int(v) for v in sys.argv[1:8]]
pos, reg = sys.argv[8] == "1", sys.argv[9] == "1"
torch.manual_seed(E + Fr)
model = M.GeneralTransformer(M.LinearProjectionEmbedding, {"patch_size": P, "embed_dim": E}, E, H, HD, Lyr, M.MLPHead, F.relu, 0.0, pos, reg, True)
g = torch.Generator().manual_seed(5)
x = 0.1 + 0.25 * torch.randn((B, Fr, P, P), generator=g).abs()
tgt = torch.rand((B, 1), generator=g)
model.cuda().train()
outs = []
for no_fwd, no_bwd in (("1", "1"), ("", "1"), ("", ""), ("1", "")):
    for var, val in (("MIVIT_NO_FUSED_ENCODER", no_fwd), ("MIVIT_NO_FUSED_ENCODER_BWD", no_bwd)):
        if val:
            os.environ[var] = val
        else:
            os.environ.pop(var, None)
    model.zero_grad()
    pred = model(x.cuda())
    F.mse_loss(pred, tgt.cuda()).backward()
    outs.append((pred.detach().cpu(), {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters()}))
p0, g0 = outs[0]
for i, name in ((1, "fused fwd / unfused bwd"), (2, "fused / fused"), (3, "unfused fwd / fused bwd")):
    p1, g1 = outs[i]
    worst = sorted(((float((g1[k] - g0[k]).norm() / (g0[k].norm() + 1e-30)), k) for k in g0), reverse=True)[:5]
    print(name, "pred diff %.2e" % (p0 - p1).abs().max().item(), ["%.2e %s" % w for w in worst])
