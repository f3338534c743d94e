#!/usr/bin/env python
"""Micro-benchmark of the shifted-row convolution kernels at the bench.py workload size
(P=13, 1024 x 30 frames -> 6.02 M rows): CUDA-event time, algorithmic TFLOP/s and the HBM GB/s
of the operand/result streams.   python scripts/bench_conv.py [reps]"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from moleculardiffusion_mivit_b200 import _lib  # noqa: E402

L = _lib.lib()
GUARD = 128
import os
P, NF = 13, int(os.environ.get("NSEQ", "1024")) * 30
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 5
pit = P + 1
rows = NF * pit * pit
rows_pad = (rows + 127) // 128 * 128
valid = NF * P * P


def rows_tensor(C, fill=True):
    t = torch.zeros((GUARD + rows_pad + GUARD, C), dtype=torch.bfloat16, device="cuda")
    if fill:
        v = t[GUARD:GUARD + rows].view(NF, pit, pit, C)
        v[:, :P, :P, :] = torch.randn((NF, P, P, C), device="cuda", dtype=torch.bfloat16)
    return t


def row0(t):
    return ctypes.c_void_p(t.data_ptr() + GUARD * t.shape[1] * 2)


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS


def pack(W, dgrad):
    co, ci, k, _ = W.shape
    wp = torch.empty(k * k * ci * co, dtype=torch.bfloat16, device="cuda")
    _lib.check(L.mivit_conv_pack_weights(_lib.ptr(W), _lib.ptr(wp), co, ci, k, dgrad, _lib.current_stream()))
    return wp


st = _lib.current_stream()
bufs = {c: [rows_tensor(c) for _ in range(2)] for c in (32, 64, 128)}
outs = {c: [rows_tensor(c, False) for _ in range(2)] for c in (32, 64, 128)}
print("%-28s %8s %10s %10s" % ("kernel", "ms", "TFLOP/s", "GB/s"))
for (ci, co, k, skip) in [(32, 64, 3, True), (64, 64, 3, False), (64, 128, 3, True), (128, 128, 3, False),
                          (128, 64, 3, False), (64, 32, 3, False), (128, 64, 1, False), (64, 32, 1, False)]:
    # forward-form GEMM: X[rows, ci] -> Y[rows, co]  (dgrad kernels are the same launch with a mirrored pack)
    W = torch.randn((co, ci, k, k), device="cuda") * 0.05
    wp = pack(W, 0)
    X, Y, Ys = bufs[ci][0], outs[co][0], outs[co][1]
    stats = torch.zeros((2, co), device="cuda")
    stats_s = torch.zeros((2, co), device="cuda")
    for with_stats in (True, False):
        if skip:
            Wk = torch.randn((co, ci, 1, 1), device="cuda") * 0.05
            wk = pack(Wk, 0)
            fn = lambda: _lib.check(L.mivit_conv_rows_fused(row0(X), _lib.ptr(wp), _lib.ptr(wk), row0(Y), row0(Ys),
                                                            _lib.ptr(stats) if with_stats else None,
                                                            _lib.ptr(stats_s) if with_stats else None, rows, P, ci, co, 1, st))
            flops = 2.0 * valid * (k * k + 1) * ci * co
            byts = rows * 2 * (ci + 2 * co)
        else:
            fn = lambda: _lib.check(L.mivit_conv_rows(row0(X), _lib.ptr(wp), row0(Y), _lib.ptr(stats) if with_stats else None,
                                                      rows, P, ci, co, k, 0, 1, st))
            flops = 2.0 * valid * k * k * ci * co
            byts = rows * 2 * (ci + co)
        ms = timeit(fn)
        print("%-28s %8.3f %10.1f %10.1f" % ("rows %dx%dx%d%s%s" % (ci, co, k * k, "+skip" if skip else "", "" if with_stats else " nostats"),
                                            ms, flops / ms / 1e9, byts / ms / 1e6))
for (ci, co, k) in [(32, 64, 3), (64, 64, 3), (64, 128, 3), (128, 128, 3), (32, 64, 1), (64, 128, 1)]:
    X, dY = bufs[ci][0], bufs[co][1]
    dW = torch.zeros((co, ci, k, k), device="cuda")
    fn = lambda: _lib.check(L.mivit_conv_rows_wgrad(row0(X), row0(dY), _lib.ptr(dW), rows, P, ci, co, k, 1, st))
    ms = timeit(fn)
    print("%-28s %8.3f %10.1f %10.1f" % ("wgrad %dx%dx%d" % (ci, co, k * k), ms, 2.0 * valid * k * k * ci * co / ms / 1e9,
                                        rows * 2 * (ci + co) / ms / 1e6))
