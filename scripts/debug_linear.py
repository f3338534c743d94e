import sys, torch, numpy as np
sys.path.insert(0, '.')
from moleculardiffusion_mivit_b200 import _lib
L = _lib.lib()
def call(mode, A, W, bias, out, M, fin, fout, relu=0, acc=0):
    _lib.check(L.mivit_linear_tf32(mode, _lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(out), M, fin, fout, relu, acc, _lib.current_stream()))
torch.manual_seed(0)
for (M, fin, fout) in [(512, 64, 64), (31781, 64, 128), (31781, 128, 64)]:
    X = torch.randn(M, fin, device='cuda'); W = torch.randn(fout, fin, device='cuda') / 8; b = torch.randn(fout, device='cuda')
    for relu, bias in [(0, None), (0, b), (1, None), (1, b)]:
        Y = torch.full((M, fout), 7.0, device='cuda')
        call(0, X, W, bias, Y, M, fin, fout, relu=relu)
        ref = X @ W.t() + (bias if bias is not None else 0)
        if relu: ref = torch.relu(ref)
        err = (Y - ref).abs()
        print('fwd', M, fin, fout, 'relu', relu, 'bias', bias is not None, 'maxerr', err.max().item(), 'bad', (err > 0.03).float().mean().item())
    dY = torch.randn(M, fout, device='cuda')
    for acc in (0, 1):
        dX = torch.ones(M, fin, device='cuda')
        call(1, dY, W, None, dX, M, fin, fout, acc=acc)
        ref = dY @ W + (1.0 if acc else 0.0)
        err = (dX - ref).abs()
        print('dgrad', M, fin, fout, 'acc', acc, 'maxerr', err.max().item(), 'bad', (err > 0.03).float().mean().item())
    dW = torch.zeros(fout, fin, device='cuda')
    call(2, dY, X, None, dW, M, fin, fout)
    ref = dY.t() @ X
    err = (dW - ref).abs()
    print('wgrad', M, fin, fout, 'maxerr', err.max().item(), 'refmax', ref.abs().max().item(), 'bad', (err > 0.5).float().mean().item())
