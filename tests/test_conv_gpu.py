"""GPU parity tests (-m gpu) of the shifted-row implicit-GEMM convolution (csrc/conv_tc.cu,
tcgen05) against torch.nn.functional.conv2d in fp32 on the same bf16-rounded operands."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GUARD = 128


def to_rows(x):
    """[NF,C,P,P] fp32 -> pitched-rows bf16 buffer [GUARD + rows_pad + GUARD, C]; returns (buf, rows)."""
    import torch
    NF, C, P, _ = x.shape
    pit = P + 1
    t = torch.zeros((NF, pit, pit, C), dtype=torch.float32, device=x.device)
    t[:, :P, :P, :] = x.permute(0, 2, 3, 1)
    rows = NF * pit * pit
    rows_pad = (rows + 127) // 128 * 128
    buf = torch.zeros((GUARD + rows_pad + GUARD, C), dtype=torch.bfloat16, device=x.device)
    buf[GUARD:GUARD + rows] = t.reshape(rows, C).to(torch.bfloat16)
    return buf, rows


def from_rows(buf, NF, C, P):
    pit = P + 1
    rows = NF * pit * pit
    t = buf[GUARD:GUARD + rows].float().reshape(NF, pit, pit, C)
    return t[:, :P, :P, :].permute(0, 3, 1, 2).contiguous(), t


def run_conv(x, W, mirrored, impl, want_stats=True):
    import torch
    from moleculardiffusion_mivit_b200 import _lib
    L = _lib.lib()
    NF, _, P, _ = x.shape
    Cout, Cin, k, _ = W.shape
    xb, rows = to_rows(x)                                   # mirrored: x is dY with Cout channels
    gin, gout = (Cout, Cin) if mirrored else (Cin, Cout)
    wp = torch.empty(k * k * Cin * Cout, dtype=torch.bfloat16, device=x.device)
    _lib.check(L.mivit_conv_pack_weights(_lib.ptr(W.contiguous()), _lib.ptr(wp), Cout, Cin, k, int(mirrored),
                                         _lib.current_stream()))
    yb = torch.full((xb.shape[0], gout), 7.0, dtype=torch.bfloat16, device=x.device)
    stats = torch.zeros((2, gout), dtype=torch.float32, device=x.device)
    row0 = GUARD * xb.shape[1] * 2
    row0y = GUARD * gout * 2
    import ctypes
    _lib.check(L.mivit_conv_rows(ctypes.c_void_p(xb.data_ptr() + row0), _lib.ptr(wp), ctypes.c_void_p(yb.data_ptr() + row0y),
                                 _lib.ptr(stats) if want_stats else None, rows, P, gin, gout, k, int(mirrored), impl,
                                 _lib.current_stream()))
    torch.cuda.synchronize()
    y, padded = from_rows(yb, NF, gout, P)
    return y, padded, stats


@pytest.mark.parametrize("impl", [0, 1, 2])
@pytest.mark.parametrize("cin,cout,k", [(32, 64, 3), (64, 64, 3), (64, 128, 3), (128, 128, 3), (32, 64, 1), (64, 128, 1)])
@pytest.mark.parametrize("P", [9, 13])
def test_conv_forward(impl, cin, cout, k, P):
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(cin * 1000 + cout + P)
    NF = 37
    x = torch.randn((NF, cin, P, P), device="cuda", generator=g).to(torch.bfloat16).float()
    W = (torch.randn((cout, cin, k, k), device="cuda", generator=g) / np.sqrt(cin * k * k)).to(torch.bfloat16).float()
    y, padded, stats = run_conv(x, W, False, impl)
    ref = F.conv2d(x.double(), W.double(), padding=k // 2).float()
    err = (y - ref).abs().max().item()
    assert err < 2e-2 * ref.abs().max().item(), err            # bf16 output rounding (2^-9 relative)
    assert (y - ref).abs().mean().item() < 3e-3 * ref.abs().mean().item() + 1e-6
    # pad column / pad line of every frame must be written as exact zeros
    assert torch.all(padded[:, P, :, :] == 0) and torch.all(padded[:, :, P, :] == 0)
    # BatchNorm statistics of the stored values
    yb = y.to(torch.bfloat16).float()
    assert torch.allclose(stats[0], yb.sum(dim=(0, 2, 3)), rtol=1e-3, atol=1e-2 * yb.abs().sum(dim=(0, 2, 3)).max().item() * 1e-2)
    assert torch.allclose(stats[1], (yb * yb).sum(dim=(0, 2, 3)), rtol=2e-3)


@pytest.mark.parametrize("impl", [0, 1, 2])
@pytest.mark.parametrize("cin,cout,k", [(32, 64, 3), (64, 64, 3), (64, 128, 3), (128, 128, 3), (64, 128, 1), (32, 64, 1)])
def test_conv_dgrad(impl, cin, cout, k):
    import torch
    import torch.nn.functional as F
    P, NF = 9, 29
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    dy = torch.randn((NF, cout, P, P), device="cuda", generator=g).to(torch.bfloat16).float()
    W = (torch.randn((cout, cin, k, k), device="cuda", generator=g) / np.sqrt(cout * k * k)).to(torch.bfloat16).float()
    dx, _, _ = run_conv(dy, W, True, impl, want_stats=False)
    ref = F.conv_transpose2d(dy.double(), W.double(), padding=k // 2).float()
    assert (dx - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


def test_conv_tc_matches_simt_large():
    """Many tiles per CTA (persistent loop), ragged last tile."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(5)
    NF, P = 1200, 13
    x = torch.randn((NF, 64, P, P), device="cuda", generator=g).to(torch.bfloat16).float()
    W = (torch.randn((64, 64, 3, 3), device="cuda", generator=g) / 24.0).to(torch.bfloat16).float()
    y1, _, s1 = run_conv(x, W, False, 1)
    y0, _, s0 = run_conv(x, W, False, 0)
    assert (y1 - y0).abs().max().item() <= 2e-2 * y0.abs().max().item()
    assert torch.allclose(s1, s0, rtol=2e-3, atol=1.0)


@pytest.mark.parametrize("impl", [0, 1, 2])
@pytest.mark.parametrize("cin,cout,k", [(32, 64, 3), (64, 64, 3), (64, 128, 3), (128, 128, 3), (32, 64, 1), (64, 128, 1)])
def test_conv_wgrad(impl, cin, cout, k):
    import ctypes
    import torch
    import torch.nn.functional as F
    from moleculardiffusion_mivit_b200 import _lib
    P, NF = 13, (41 if impl == 0 else 900)     # 900 frames: several stages per CTA (both mbarrier phases)
    g = torch.Generator(device="cuda").manual_seed(cin * 7 + cout + k)
    x = torch.randn((NF, cin, P, P), device="cuda", generator=g).to(torch.bfloat16).float()
    dy = torch.randn((NF, cout, P, P), device="cuda", generator=g).to(torch.bfloat16).float()
    xb, rows = to_rows(x)
    db, _ = to_rows(dy)
    dW = torch.zeros((cout, cin, k, k), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().mivit_conv_rows_wgrad(ctypes.c_void_p(xb.data_ptr() + GUARD * cin * 2),
                                                ctypes.c_void_p(db.data_ptr() + GUARD * cout * 2), _lib.ptr(dW), rows, P, cin,
                                                cout, k, impl, _lib.current_stream()))
    torch.cuda.synchronize()
    W = torch.zeros((cout, cin, k, k), dtype=torch.float64, device="cuda", requires_grad=True)
    y = F.conv2d(x.double(), W, padding=k // 2)
    ref, = torch.autograd.grad(y, W, dy.double())
    ref = ref.float()
    assert (dW - ref).abs().max().item() < 2e-3 * ref.abs().max().item() + 1e-3, (dW - ref).abs().max().item()


@pytest.mark.parametrize("cin,cout", [(32, 64), (64, 128)])
@pytest.mark.parametrize("NF", [3, 700])
def test_conv_fused_skip(cin, cout, NF):
    """conv1 (3x3) + skip (1x1) of a ResidualBlock in one pipelined launch; NF = 700 gives every
    persistent CTA several tiles (double-buffer wrap-around, both mbarrier phases)."""
    import ctypes
    import torch
    import torch.nn.functional as F
    from moleculardiffusion_mivit_b200 import _lib
    L = _lib.lib()
    P = 13
    g = torch.Generator(device="cuda").manual_seed(cin + NF)
    x = torch.randn((NF, cin, P, P), device="cuda", generator=g).to(torch.bfloat16).float()
    W = (torch.randn((cout, cin, 3, 3), device="cuda", generator=g) / np.sqrt(cin * 9)).to(torch.bfloat16).float()
    Ws = (torch.randn((cout, cin, 1, 1), device="cuda", generator=g) / np.sqrt(cin)).to(torch.bfloat16).float()
    xb, rows = to_rows(x)
    wp = torch.empty(9 * cin * cout, dtype=torch.bfloat16, device="cuda")
    ws = torch.empty(cin * cout, dtype=torch.bfloat16, device="cuda")
    _lib.check(L.mivit_conv_pack_weights(_lib.ptr(W), _lib.ptr(wp), cout, cin, 3, 0, _lib.current_stream()))
    _lib.check(L.mivit_conv_pack_weights(_lib.ptr(Ws), _lib.ptr(ws), cout, cin, 1, 0, _lib.current_stream()))
    yb = torch.full((xb.shape[0], cout), 7.0, dtype=torch.bfloat16, device="cuda")
    sb = torch.full((xb.shape[0], cout), 7.0, dtype=torch.bfloat16, device="cuda")
    st = torch.zeros((2, 2, cout), dtype=torch.float32, device="cuda")
    _lib.check(L.mivit_conv_rows_fused(ctypes.c_void_p(xb.data_ptr() + GUARD * cin * 2), _lib.ptr(wp), _lib.ptr(ws),
                                       ctypes.c_void_p(yb.data_ptr() + GUARD * cout * 2),
                                       ctypes.c_void_p(sb.data_ptr() + GUARD * cout * 2), _lib.ptr(st[0]), _lib.ptr(st[1]),
                                       rows, P, cin, cout, 1, _lib.current_stream()))
    torch.cuda.synchronize()
    for buf, w, pad, s in ((yb, W, 1, st[0]), (sb, Ws, 0, st[1])):
        y, padded = from_rows(buf, NF, cout, P)
        ref = F.conv2d(x.double(), w.double(), padding=pad).float()
        assert (y - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
        assert torch.all(padded[:, P, :, :] == 0) and torch.all(padded[:, :, P, :] == 0)
        assert torch.allclose(s[0], y.sum(dim=(0, 2, 3)), rtol=2e-3, atol=0.05 * NF)
        assert torch.allclose(s[1], (y * y).sum(dim=(0, 2, 3)), rtol=2e-3)
