"""Diagnosis of an intermittent 1e-3 error in a BatchNorm bias gradient of the CNN baselines: runs the golden cases repeatedly with
the workspace poisoned (NaN / large values) before the forward, so a read of uninitialised workspace shows up deterministically."""
import os
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import torch.nn.functional as F
import test_resnet_gpu as T
from moleculardiffusion_mivit_b200 import baselines as BL

poison = sys.argv[1] if len(sys.argv) > 1 else "nan"
orig = BL._CudaResNet._workspace if hasattr(BL, "_CudaResNet") else None
cls = [c for c in vars(BL).values() if isinstance(c, type) and hasattr(c, "_workspace")][0]
orig = cls._workspace
def patched(self, cfg, B):
    ws = orig(self, cfg, B)
    if not getattr(self, "_poisoned", False):
        if poison == "nan":
            ws.view(torch.float32)[: ws.numel() // 4].fill_(float("nan"))
        else:
            ws.view(torch.float32)[: ws.numel() // 4].fill_(1.0e3)
        self._poisoned = True
    return ws
cls._workspace = patched
gd = os.path.join("tests", "golden")
for name in ["resnet_ft_p9", "resnet_p9"]:
    z, sd, x, tgt, ext = T.load(gd, name)
    ref_pred, ref_loss, ref_g, ref_stats = T.rn.loss_and_grads(sd, x, tgt, ext, T.CASES[name]["single"])
    for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
        model = T.build(name); model.load_state_dict(sd); model.cuda().train()
        pred = model(x.cuda(), ext.cuda()) if ext is not None else model(x.cuda())
        loss = F.mse_loss(pred, tgt.cuda()); loss.backward()
        bad = [k for k, p in model.named_parameters() if not torch.isfinite(p.grad).all()]
        worst = max(((T.relnorm(torch.nan_to_num(p.grad.cpu()), ref_g[k]), k) for k, p in model.named_parameters() if float(ref_g[k].norm()) > 1e-6),
                    key=lambda t: t[0])
        perr = (pred.cpu() - ref_pred).abs().max().item()
        print(name, it, "pred err %.2e" % perr, "worst grad %.2e %s" % worst, "non-finite:", bad[:4])
