"""Host-side logic of the data-parallel path (one process per GPU).  The
reference is single-process, so nothing here has a reference counterpart; the contract is that an
N-rank run trains on exactly the data set and (for BatchNorm-free models) with exactly the gradient
of the 1-rank run at the same global batch:

  * sequences are keyed by GLOBAL id (Philox counter), rank r of W takes ids
    [ (step*W + r)*B, (step*W + r + 1)*B ) -- `shard_offset`;
  * every rank computes the gradient of ITS mean loss; the flat 2 MB gradient buffer is summed
    with one all-reduce and AdamW applies grad_scale = 1/W -- `allreduce_sum_` + grad_scale.

Pure torch.distributed (NCCL on GPUs, gloo in the CPU tests); no CUDA kernels are called here."""
import torch


def shard_offset(step, world, rank, batch_per_rank):
    """Global id of the first sequence rank `rank` processes in step `step`."""
    return (int(step) * int(world) + int(rank)) * int(batch_per_rank)


def shard_slices(global_batch, world):
    """Even split of a global batch into `world` contiguous slices (remainder to the first ranks)."""
    base, rem = divmod(int(global_batch), int(world))
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append(slice(start, start + n))
        start += n
    return out


def allreduce_sum_(flat, group=None):
    """In-place SUM all-reduce of the flat gradient buffer; returns the grad_scale (1/world) that the
    optimizer must apply to turn the sum of per-rank mean-loss gradients into the global-batch gradient."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def broadcast_state_(tensors, group=None, src=0):
    """Rank `src`'s copy of every tensor replaces the other ranks' (what DistributedDataParallel does with parameters and buffers
    at construction): ranks built from different seeds or checkpoints start the data-parallel run from identical state."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) <= 1:
        return
    root = dist.get_global_rank(group, src) if group is not None else src
    for t in tensors:
        dist.broadcast(t, src=root, group=group)
