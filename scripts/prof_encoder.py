"""Runs a few training steps of a Linear-embedding ViT (E64/H4/HD128/L6, P=9, F=30) at B sequences: the transformer kernels only
(for ncu):  python scripts/prof_encoder.py [B] [steps]"""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from moleculardiffusion_mivit_b200 import models as M
from moleculardiffusion_mivit_b200.training import MiViTTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
model = M.GeneralTransformer(M.LinearProjectionEmbedding, {"patch_size": 9, "embed_dim": 64}, 64, 4, 128, 6, M.MLPHead, F.relu, 0.0,
                             False, True, True).cuda().train()
tr = MiViTTrainer(model, lr=1e-4)
x = torch.rand(B, 30, 9, 9, device="cuda")
y = torch.rand(B, 1, device="cuda")
for _ in range(2):
    tr.train_step(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = tr.train_step(x, y)
e1.record()
torch.cuda.synchronize()
print("B=%d: %.3f ms/step, loss %.5f" % (B, e0.elapsed_time(e1) / steps, loss.item()))
