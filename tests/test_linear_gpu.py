"""GPU parity tests (-m gpu) of the tcgen05 kind::tf32 nn.Linear kernels (csrc/linear_tc.cu) against
torch fp64 matmul; tolerance = TF32 operand rounding (10-bit mantissa): 2e-3 of the output scale."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(64, 64), (64, 128), (128, 64), (32, 32), (32, 64), (64, 192), (128, 128)]


def call(mode, A, W, bias, out, M, fin, fout, relu=0, acc=0):
    from moleculardiffusion_mivit_b200 import _lib
    _lib.check(_lib.lib().mivit_linear_tf32(mode, _lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(out), M, fin, fout, relu, acc,
                                            _lib.current_stream()))


@pytest.mark.parametrize("fin,fout", SHAPES)
@pytest.mark.parametrize("M", [512, 31744 + 37])
def test_linear_forward_and_dgrad(fin, fout, M):
    import torch
    g = torch.Generator(device="cuda").manual_seed(fin * 131 + fout + M)
    X = torch.randn((M, fin), device="cuda", generator=g)
    W = torch.randn((fout, fin), device="cuda", generator=g) / np.sqrt(fin)
    b = torch.randn((fout,), device="cuda", generator=g)
    Y = torch.full((M, fout), 7.0, device="cuda")
    call(0, X, W, b, Y, M, fin, fout, relu=1)
    ref = torch.relu(X.double() @ W.double().t() + b.double()).float()
    assert (Y - ref).abs().max().item() < 2e-3 * ref.abs().max().item() + 2e-3
    dY = torch.randn((M, fout), device="cuda", generator=g)
    dX = torch.ones((M, fin), device="cuda")
    if fout > 128:   # the resident weights + two 128-row dY slabs exceed shared memory: the ViT uses its SIMT GEMM there
        from moleculardiffusion_mivit_b200._lib import MivitError
        with pytest.raises(MivitError, match="not supported"):
            call(1, dY, W, None, dX, M, fin, fout, acc=1)
        return
    call(1, dY, W, None, dX, M, fin, fout, acc=1)
    refd = (dY.double() @ W.double()).float() + 1.0
    assert (dX - refd).abs().max().item() < 2e-3 * refd.abs().max().item() + 2e-3


@pytest.mark.parametrize("fin,fout", [(64, 64), (64, 128), (128, 64), (32, 32), (32, 64), (64, 32)])
def test_linear_wgrad(fin, fout):
    import torch
    M = 31744 + 37
    g = torch.Generator(device="cuda").manual_seed(fin + 7 * fout)
    X = torch.randn((M, fin), device="cuda", generator=g)
    dY = torch.randn((M, fout), device="cuda", generator=g)
    dW = torch.zeros((fout, fin), device="cuda")
    db = torch.full((fout,), 2.0, device="cuda")
    call(2, dY, X, db, dW, M, fin, fout)
    ref = (dY.double().t() @ X.double()).float()
    assert (dW - ref).abs().max().item() < 2e-3 * ref.abs().max().item() + 0.3   # sum of 31k tf32 products
    refb = dY.double().sum(0).float() + 2.0                                      # bias gradient rides along (accumulated)
    assert (db - refb).abs().max().item() < 2e-3 * refb.abs().max().item() + 0.3
    dW2 = torch.zeros((fout, fin), device="cuda")
    call(2, dY, X, None, dW2, M, fin, fout)                                       # without the bias column
    assert (dW2 - ref).abs().max().item() < 2e-3 * ref.abs().max().item() + 0.3
