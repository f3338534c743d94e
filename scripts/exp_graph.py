"""Experiment: CUDA-graph replay of the ViT training step (forward + loss + backward captured, AdamW eager)."""
import ctypes, sys, time
import torch, torch.nn.functional as F
sys.path.insert(0, ".")
from moleculardiffusion_mivit_b200 import _lib, models as M
from moleculardiffusion_mivit_b200.training import MiViTTrainer
L = _lib.lib()
B, P, Fr = 1024, 13, 30
torch.manual_seed(0)
model = M.GeneralTransformer(M.DeepResNetEmbedding, {"patch_size": P, "embed_dim": 64}, 64, 4, 128, 6, M.MLPHead, F.relu, 0.0, False, True, True).cuda().train()
tr = MiViTTrainer(model, lr=1e-4)
x = (0.1 + 0.25 * torch.randn(B, Fr, P, P, device="cuda").abs()).contiguous()
y = torch.rand(B, 1, device="cuda")

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("eager   %.3f ms/step" % timeit(lambda: tr.train_step(x, y)))
cfg = model.vit_config(Fr); ws = model._workspace(cfg, B); pred, dpred = tr._buffers(B)
def fwd_bwd():
    _lib.check(L.mivit_vit_train_step(ctypes.byref(cfg), B, _lib.ptr(x), None, _lib.ptr(y), _lib.ptr(model._flat), _lib.ptr(model._grad_flat),
        _lib.ptr(tr.m), _lib.ptr(tr.v), _lib.ptr(model._bn_flat), _lib.ptr(model._bn_nbt), _lib.ptr(ws), _lib.ptr(pred), _lib.ptr(tr.loss),
        _lib.ptr(dpred), tr.lr, 0.9, 0.999, 1e-8, 0.01, 1, 0, _lib.current_stream()))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    fwd_bwd(); fwd_bwd()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    fwd_bwd()
torch.cuda.synchronize()
step = [tr.step_count]
def graphed():
    g.replay()
    step[0] += 1
    _lib.check(L.mivit_adamw_step(_lib.ptr(model._flat), _lib.ptr(model._grad_flat), _lib.ptr(tr.m), _lib.ptr(tr.v), model._n_params,
                                  tr.lr, 0.9, 0.999, 1e-8, 0.01, step[0], 1.0, _lib.current_stream()))
print("graphed %.3f ms/step" % timeit(graphed), "loss", tr.loss.item())
