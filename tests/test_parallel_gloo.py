"""CPU test of the N>1 path (gloo, world_size 2): per-rank sharding by global sequence id plus one
SUM all-reduce with grad_scale = 1/W reproduces the single-process gradient at the same global batch.
The per-rank gradient comes from the fp32 oracle (no GPU here); the host logic under test is
moleculardiffusion_mivit_b200/parallel.py, which the GPU trainer uses unchanged."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import vit_oracle as vo
from vit_cases import CASES, load_case

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from moleculardiffusion_mivit_b200.parallel import allreduce_sum_, shard_offset, shard_slices
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    name = "linear_s_pos"
    z, sd, x, tgt, feats = load_case(GOLDEN, name)
    sl = shard_slices(x.shape[0], world)[rank]
    assert shard_offset(3, world, rank, 2) == (3 * world + rank) * 2
    _, _, g, _ = vo.loss_and_grads(sd, CASES[name], x[sl], tgt[sl], None)
    keys = sorted(g)
    flat = torch.cat([g[k].reshape(-1) for k in keys])
    scale = allreduce_sum_(flat)
    flat *= scale
    if rank == 0:
        np.save(os.path.join(out_dir, "dp_grad.npy"), flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    dp = np.load(os.path.join(str(tmp_path), "dp_grad.npy"))
    name = "linear_s_pos"
    z, sd, x, tgt, _ = load_case(GOLDEN, name)
    _, _, g, _ = vo.loss_and_grads(sd, CASES[name], x, tgt, None)
    ref = torch.cat([g[k].reshape(-1) for k in sorted(g)]).numpy()
    assert np.abs(dp - ref).max() < 1e-6 * max(1.0, np.abs(ref).max())


def test_shard_slices_cover_batch():
    from moleculardiffusion_mivit_b200.parallel import shard_slices
    for n, w in [(8, 2), (7, 2), (352, 8), (5, 8)]:
        s = shard_slices(n, w)
        assert sum(x.stop - x.start for x in s) == n and s[0].start == 0 and s[-1].stop == n
        assert max(x.stop - x.start for x in s) - min(x.stop - x.start for x in s) <= 1
