"""Data-parallel step with the gradient all-reduce overlapped with the backward: the backward is issued
in two parts (include/mivit.h: mivit_vit_backward_part), the bucket of everything but the image embedding is reduced while the
image-embedding backward runs.  Two ranks share the test box's GPU over gloo (its all-reduce accepts CUDA tensors and
async_op); the overlapped trainer must follow the plain one (one all-reduce after the whole backward) step for step, kernel
by kernel and through the two-graph replay.
Round 2: the same comparison for the PEER-MEMORY path (csrc/peer_comm.cu: each rank maps the other's segment through CUDA IPC;
the gradient all-reduce fused with AdamW is one kernel per bucket, the whole data-parallel step one CUDA graph) -- the two ranks
sharing the GPU exchange their gradients through each other's memory exactly as two GPUs would over NVLink."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from vit_cases import load_case

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from test_vit_gpu import build
    from moleculardiffusion_mivit_b200.training import MiViTTrainer
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {}
    for name in ("deepcnn_n", "linear_s_pos"):
        _, sd, x, tgt, _ = load_case(GOLDEN, name)
        g = torch.Generator().manual_seed(100 + rank)                  # different data on the two ranks
        xs = [(x + 0.05 * torch.randn(x.shape, generator=g)).cuda() for _ in range(3)]
        ts = [torch.rand(tgt.shape, generator=g).cuda() for _ in range(3)]
        for tag, overlap, graph, fused in (("plain", False, False, False), ("overlap", True, False, False),
                                           ("overlap_graph", True, True, False), ("fused", False, False, True),
                                           ("fused_overlap", True, False, True), ("fused_overlap_graph", True, True, True)):
            model = build(name)
            model.load_state_dict(sd)
            model.cuda().train()
            tr = MiViTTrainer(model, lr=1e-4, overlap_allreduce=overlap, cuda_graph=graph, fused_allreduce=fused)
            assert tr.world == world and (tr.comm is not None) == fused
            losses = [float(tr.train_step(a, b).item()) for a, b in zip(xs, ts)]
            torch.cuda.synchronize()
            n = model._n_params
            res["%s/%s/grad" % (name, tag)] = tr.reduced_gradient().detach().cpu().numpy()     # all-reduced SUM of the last step
            res["%s/%s/w" % (name, tag)] = model._flat[:n].detach().cpu().numpy()
            res["%s/%s/loss" % (name, tag)] = np.asarray(losses)
            res["%s/%s/ne" % (name, tag)] = np.int64(tr._n_embedding(model.vit_config(x.shape[1])))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **res)
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_allreduce_follows_the_plain_data_parallel_step(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % i)) for i in range(world)]
    for name, tol in (("deepcnn_n", 3e-2), ("linear_s_pos", 2e-3)):
        ne = int(r[0]["%s/overlap/ne" % name])
        assert 0 < ne < r[0]["%s/plain/grad" % name].size
        for tag in ("overlap", "overlap_graph", "fused", "fused_overlap", "fused_overlap_graph"):
            for k in range(world):
                g0, g1 = r[k]["%s/plain/grad" % name], r[k]["%s/%s/grad" % (name, tag)]
                for lo, hi in ((0, ne), (ne, g0.size)):                    # both buckets were reduced
                    rel = np.linalg.norm(g1[lo:hi] - g0[lo:hi]) / np.linalg.norm(g0[lo:hi])
                    assert rel < tol, (name, tag, k, lo, rel)              # 3 AdamW steps of atomics-order noise in between
                assert np.allclose(r[k]["%s/plain/loss" % name], r[k]["%s/%s/loss" % (name, tag)], rtol=2e-2, atol=1e-4)
            # the reduced gradient is the same on both ranks
            assert np.array_equal(r[0]["%s/%s/grad" % (name, tag)], r[1]["%s/%s/grad" % (name, tag)])
            if tag.startswith("fused"):      # rank-ordered sums + identical AdamW arithmetic: the replicas stay BIT-identical
                assert np.array_equal(r[0]["%s/%s/w" % (name, tag)], r[1]["%s/%s/w" % (name, tag)])
            # three AdamW steps moved the weights like the plain data-parallel trainer's (lr 1e-4: |update| <= ~3e-4 per weight)
            assert np.abs(r[0]["%s/%s/w" % (name, tag)] - r[0]["%s/plain/w" % name]).max() < 7e-4
