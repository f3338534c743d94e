#!/usr/bin/env python
"""Micro-benchmark of the tf32 nn.Linear kernels at the bench.py token count (1024 x 31 tokens)."""
import sys
import torch
sys.path.insert(0, ".")
from moleculardiffusion_mivit_b200 import _lib  # noqa: E402
L = _lib.lib()
import os
M = int(os.environ.get("NSEQ", "1024")) * 31
REPS = 20


def call(mode, A, W, bias, out, fin, fout, relu=0, acc=0):
    _lib.check(L.mivit_linear_tf32(mode, _lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(out), M, fin, fout, relu, acc, _lib.current_stream()))


def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS * 1e3


for fin, fout in [(64, 64), (64, 128), (128, 64)]:
    X = torch.randn(M, fin, device="cuda"); W = torch.randn(fout, fin, device="cuda") / 8; b = torch.randn(fout, device="cuda")
    Y = torch.empty(M, fout, device="cuda"); dY = torch.randn(M, fout, device="cuda"); dX = torch.empty(M, fin, device="cuda")
    dW = torch.zeros(fout, fin, device="cuda")
    byts = M * (fin + fout) * 4
    for name, fn in [("fwd", lambda: call(0, X, W, b, Y, fin, fout)), ("dgrad", lambda: call(1, dY, W, None, dX, fin, fout)),
                     ("wgrad", lambda: call(2, dY, X, None, dW, fin, fout))]:
        us = timeit(fn)
        print("%-6s %3d->%3d  %7.1f us  %7.1f GB/s" % (name, fin, fout, us, byts / us / 1e3))
