"""Host-side logic of the data-parallel path (one process per GPU).  The
reference is single-process, so nothing here has a reference counterpart; the contract is that an
N-rank run trains on exactly the data set and (for BatchNorm-free models) with exactly the gradient
of the 1-rank run at the same global batch:

  * sequences are keyed by GLOBAL id (Philox counter), rank r of W takes ids
    [ (step*W + r)*B, (step*W + r + 1)*B ) -- `shard_offset`;
  * every rank computes the gradient of ITS mean loss; the flat 2 MB gradient buffer is summed
    with one all-reduce and AdamW applies grad_scale = 1/W -- `allreduce_sum_` + grad_scale.

`shard_offset` / `shard_slices` / `allreduce_sum_` / `broadcast_state_` are pure torch.distributed (NCCL on GPUs, gloo in the CPU
tests).  `PeerCommunicator` sets up the peer-memory segments of csrc/peer_comm.cu (include/mivit.h: mivit_comm_*): the
gradient exchange itself is then ONE kernel fused with AdamW, without NCCL on the data path; torch.distributed only carries
the 64-byte IPC handles at start-up."""
import ctypes

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


class _DevicePointer:
    """Zero-copy view of raw device memory as a torch tensor (through __cuda_array_interface__)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerCommunicator:
    """One peer-memory SEGMENT per rank (header | small-exchange slots | flat gradient buffer), allocated by the C library,
    exported with a CUDA IPC handle and mapped by every peer of the group -- the set-up of mivit_allreduce_adamw /
    mivit_allreduce_small (csrc/peer_comm.cu).  All ranks must live on one node (NVLink / NVSwitch or PCIe peer access; two
    ranks sharing one GPU work too, which is what the single-GPU tests do).  `group`: torch.distributed process group used only
    for the handle exchange and a barrier."""

    def __init__(self, n_grad_floats, group=None):
        import torch.distributed as dist
        from . import _lib
        L = _lib.lib()
        self._lib, self._L = _lib, L
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("PeerCommunicator covers one NVSwitch domain: at most 8 ranks")
        self.n = int(n_grad_floats)
        nbytes = int(L.mivit_comm_segment_bytes(self.n))
        seg = ctypes.c_void_p()
        _lib.check(L.mivit_comm_alloc(nbytes, ctypes.byref(seg)))
        self._own = seg.value
        handle = (ctypes.c_uint8 * 64)()
        _lib.check(L.mivit_comm_ipc_handle(ctypes.c_void_p(self._own), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.comm = _lib.PeerComm()
        self.comm.rank, self.comm.world = self.rank, self.world
        self._opened = []
        for r, hb in enumerate(handles):
            if r == self.rank:
                self.comm.segment[r] = self._own
                continue
            p = ctypes.c_void_p()
            _lib.check(L.mivit_comm_ipc_open((ctypes.c_uint8 * 64).from_buffer_copy(hb), ctypes.byref(p)))
            self.comm.segment[r] = p.value
            self._opened.append(p.value)
        goff = int(L.mivit_comm_grad_offset_bytes())
        dev = torch.device("cuda", torch.cuda.current_device())
        self.grad = torch.as_tensor(_DevicePointer(self._own + goff, ((nbytes - goff) // 4,), "<f4"), device=dev)
        self.grad.zero_()
        torch.cuda.synchronize()
        dist.barrier(group=group)          # every segment is mapped and zeroed before the first exchange

    def ref(self):
        return ctypes.byref(self.comm)

    def set_lr(self, lr):
        self._lib.check(self._L.mivit_comm_set_lr(self.ref(), float(lr), self._lib.current_stream()))

    def set_step(self, steps_done):
        self._lib.check(self._L.mivit_comm_set_step(self.ref(), int(steps_done), self._lib.current_stream()))

    def allreduce_small_(self, buf, call=0):
        """In-place SUM all-reduce of a CUDA float32 tensor of <= 512 elements (synchronised-BatchNorm statistics)."""
        self._lib.check(self._L.mivit_allreduce_small(self.ref(), int(call), self._lib.ptr(buf), buf.numel(), self._lib.current_stream()))
        return buf

    def close(self):
        if self._own is None:
            return
        torch.cuda.synchronize()
        for p in self._opened:
            self._L.mivit_comm_ipc_close(ctypes.c_void_p(p))
        self._opened = []
        self.grad = None
        self._L.mivit_comm_free(ctypes.c_void_p(self._own))
        self._own = None
