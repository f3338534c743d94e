"""Synchronised BatchNorm (the data-parallel PARITY mode): two ranks with B/2 sequences each must
reproduce the single-process training step on the B sequences -- loss, every gradient, the BatchNorm running statistics.
Both ranks share the one GPU of the test box; the process group is gloo (its all-reduce accepts CUDA tensors), so the test
exercises the real path: C library -> mivit_set_allreduce_hook callback -> torch.distributed -> back into the kernels.
(The same test passes with NCCL on two GPUs: MIVIT_TEST_BACKEND=nccl under `gpurun --gpus 2`.)"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from vit_cases import load_case

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAME = "deepcnn_n"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch():
    """8 sequences: the 4 golden ones and a perturbed copy (an even split must not give both ranks the same data)."""
    import torch
    _, sd, x, tgt, _ = load_case(GOLDEN, NAME)
    g = torch.Generator().manual_seed(11)
    x2 = x.flip(0) + 0.05 * torch.randn(x.shape, generator=g)
    t2 = torch.rand(tgt.shape, generator=g)
    return sd, torch.cat([x, x2]).contiguous(), torch.cat([tgt, t2]).contiguous()


def _step(sd, x, tgt, sync_bn, fused=False, graph=False, steps=1):
    import torch
    from test_vit_gpu import build
    from moleculardiffusion_mivit_b200.training import MiViTTrainer
    model = build(NAME)
    model.load_state_dict(sd)
    model.cuda().train()
    tr = MiViTTrainer(model, lr=1e-4 if steps == 1 else 0.0, sync_bn=sync_bn, fused_allreduce=fused, cuda_graph=graph)
    for _ in range(steps):       # (lr = 0 for the multi-step graph variant: every step sees the same weights)
        loss = tr.train_step(x.cuda(), tgt.cuda())
    torch.cuda.synchronize()
    return tr, float(loss.item()), tr.reduced_gradient().detach().cpu().clone(), model._bn_flat.detach().cpu().clone()


def _worker(rank, world, port, backend, out_dir, mode):
    import torch
    import torch.distributed as dist
    from moleculardiffusion_mivit_b200.parallel import shard_slices
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank if backend == "nccl" else 0)
    dist.init_process_group(backend, rank=rank, world_size=world)
    sd, x, tgt = _batch()
    sl = shard_slices(x.shape[0], world)[rank]
    # "hook": statistics through the host all-reduce callback (torch.distributed), kernel by kernel; "peer": through the peer
    # segments (plain kernels); "peer_graph": the same captured -- step 1 runs eagerly, steps 2 and 3 replay ONE graph that
    # contains the 12 small all-reduces and the fused gradient all-reduce + AdamW
    tr, loss, grad, bn = _step(sd, x[sl], tgt[sl], True, fused=mode != "hook", graph=mode == "peer_graph",
                               steps=3 if mode == "peer_graph" else 1)
    assert tr.sync_bn and tr.world == world and (tr.comm is not None) == (mode != "hook")
    # the reduced gradient is the SUM of the per-rank mean-loss gradients; AdamW applied 1/W
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), loss=loss, grad=(grad / world).numpy(), bn=bn.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["hook", "peer", "peer_graph"])
def test_two_rank_sync_bn_step_equals_single_process(tmp_path, mode):
    import torch
    import torch.multiprocessing as mp
    backend = os.environ.get("MIVIT_TEST_BACKEND", "gloo")
    if backend == "nccl" and torch.cuda.device_count() < 2:
        pytest.skip("nccl needs two GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), backend, str(tmp_path), mode), nprocs=world, join=True)
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % i)) for i in range(world)]
    sd, x, tgt = _batch()
    _, loss, grad, bn = _step(sd, x, tgt, False, steps=3 if mode == "peer_graph" else 1)
    # same all-reduced gradient and the same (global) running statistics on both ranks
    assert np.array_equal(r[0]["grad"], r[1]["grad"])
    assert np.allclose(r[0]["bn"], r[1]["bn"], rtol=0, atol=0)
    # global loss = mean of the two half-batch means
    assert abs(0.5 * (float(r[0]["loss"]) + float(r[1]["loss"])) - loss) < 2e-3 * max(1.0, abs(loss))
    # running statistics of the GLOBAL batch (per-rank statistics would differ at the 1e-1 level on half batches)
    assert np.allclose(r[0]["bn"], bn.numpy(), rtol=2e-3, atol=2e-4)
    g_dp, g_1 = torch.from_numpy(r[0]["grad"]), grad
    rel = float((g_dp - g_1).norm() / g_1.norm())
    assert rel < 3e-2, rel          # same bf16 arithmetic, different tile / summation order (and bf16 rounding of the halves)


def test_sync_bn_differs_from_per_rank_statistics(tmp_path):
    """Guard against a silently ignored hook: the default per-rank-statistics step on a half batch must NOT give the global
    running statistics."""
    sd, x, tgt = _batch()
    _, _, _, bn_half = _step(sd, x[:4], tgt[:4], False)
    _, _, _, bn_full = _step(sd, x, tgt, False)
    assert not np.allclose(bn_half.numpy(), bn_full.numpy(), rtol=2e-3, atol=2e-4)
