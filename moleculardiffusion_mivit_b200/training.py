"""Fused training step of the reference loop body
(Experiments/PSFNoise/trainModelsPSFNoise.py:182-196: zero_grad, model(x), MSELoss, backward,
AdamW.step; StepLR(5, 0.9) stepped once per cycle) as ONE C-ABI call on flat buffers, plus the
data-parallel variant: per-rank generation, one NCCL all-reduce of the 2 MB flat gradient, AdamW
with grad_scale = 1/world_size.  Nothing here exists in the reference (it is single-process)."""
import ctypes

import torch

from . import _lib
from .models import _CudaViT, TrajectorySource
from .parallel import PeerCommunicator, allreduce_sum_, broadcast_state_

__all__ = ["MiViTTrainer"]


class MiViTTrainer:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, step_size=5, gamma=0.9,
                 process_group=None, distributed=None, cuda_graph=False, sync_bn=False, overlap_allreduce=True,
                 fused_allreduce=True):
        if not isinstance(model, _CudaViT):
            raise TypeError("MiViTTrainer drives a moleculardiffusion_mivit_b200.models.GeneralTransformer / ModularTransformer")
        self.model = model
        self.base_lr, self.lr = float(lr), float(lr)
        self.betas, self.eps, self.weight_decay = betas, float(eps), float(weight_decay)
        self.step_size, self.gamma, self.epoch = int(step_size), float(gamma), 0
        self.step_count = 0
        model._ensure_flat()
        n = model._n_params
        dev = model._flat.device
        self.m = torch.zeros(n + 4, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n + 4, dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self._scratch = {}
        # cuda_graph=True: the ~270 launches of forward + loss + backward are captured once per (batch shape, lr) and replayed
        # (a dependent in-stream launch costs ~4 us on B200; the replay saves ~0.45 ms of a 21 ms step); AdamW stays eager
        # because its bias corrections change every step.
        self.cuda_graph = bool(cuda_graph)
        self._graphs = {}
        self._graph_ws = self._graph_flat = None
        import torch.distributed as dist
        self.dist = dist if (distributed if distributed is not None else (dist.is_available() and dist.is_initialized())) else None
        self.group = process_group
        self.world = self.dist.get_world_size(process_group) if self.dist else 1
        # sync_bn=True (data-parallel parity mode): BatchNorm statistics and their backward sums are all-reduced
        # over the group inside the step (include/mivit.h: mivit_set_allreduce_hook), so W ranks with B/W sequences each
        # reproduce the single-process step on B sequences.  Default is per-rank statistics (stock DDP semantics).
        self.sync_bn = bool(sync_bn) and self.world > 1
        self._hook = _lib.ALLREDUCE_FN(self._allreduce_hook) if self.sync_bn else None
        self._hook_ws = None
        self._hook_error = None
        # overlap_allreduce (data-parallel): the backward is issued in two parts; the gradients of everything but the
        # image embedding (head, encoder layers, tokens: 42 % of the 2 MB) are all-reduced on the collective's own stream while the
        # image-embedding backward (~45 % of the step) runs, the embedding's bucket follows, AdamW waits for both.
        self.overlap_allreduce = bool(overlap_allreduce)
        self.fused_allreduce = bool(fused_allreduce)
        self._ne = None
        if any(not p.requires_grad for p in model.parameters()):
            # the fused AdamW walks the whole flat buffer; torch.optim.AdamW would skip frozen parameters
            raise NotImplementedError("MiViTTrainer updates every parameter: requires_grad=False parameters are not supported")
        self.comm = None
        self._side = None
        if self.world > 1:
            # like DistributedDataParallel: every rank starts from rank 0's parameters, BatchNorm statistics and (zero) moments
            bufs = [model._flat]
            if model._is_deep():
                bufs += [model._bn_flat, model._bn_nbt]
            broadcast_state_(bufs, self.group)
            if self.fused_allreduce:
                # peer-memory data path (csrc/peer_comm.cu): the model's flat gradient buffer moves into this rank's segment, the
                # gradient all-reduce + AdamW become one kernel per bucket and -- being ordinary kernels -- part of the step's graph
                self.comm = PeerCommunicator(n + 4, self.group)
                model._grad_flat = self.comm.grad[:n + 4]
                self._gsum = torch.zeros(n + 4, dtype=torch.float32, device=dev)
                self._side = torch.cuda.Stream()
                self.comm.set_lr(self.lr)
                self.comm.set_step(self.step_count)

    def _allreduce_hook(self, buf, n_floats, stream, user):
        """C callback (mivit_allreduce_fn): SUM all-reduce of n_floats fp32 at device address `buf`, which always lies inside
        the step's workspace tensor; enqueued by torch.distributed in order with the current stream."""
        try:
            ws = self._hook_ws
            off = int(buf) - ws.data_ptr()
            if off < 0 or off + 4 * n_floats > ws.numel() * ws.element_size() or off % 4:
                raise ValueError("all-reduce buffer outside the workspace")
            view = ws.view(torch.uint8).reshape(-1)[off:off + 4 * n_floats].view(torch.float32)
            self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, group=self.group)
            return 0
        except Exception as e:      # an exception must not unwind through the C frames
            self._hook_error = e
            return 1

    def launch_description(self):
        if self.comm is not None:
            return ("ONE CUDA-graph replay per step: forward+loss+backward+peer-memory all-reduce+AdamW" if self.cuda_graph
                    else "kernel by kernel")
        if self.sync_bn and self.model._is_deep():
            return "kernel by kernel (synchronised BatchNorm through the host all-reduce hook)"
        return "CUDA-graph replay of forward+loss+backward, eager AdamW" if self.cuda_graph else "kernel by kernel"

    def allreduce_description(self):
        if self.world == 1:
            return None
        if self.comm is not None:
            two = self.overlap_allreduce and self.model._image_embedding() is not None
            return ("fused peer-memory all-reduce + AdamW kernel (no NCCL on the data path), " +
                    ("two buckets, the non-embedding one overlapped with the image-embedding backward" if two else "one bucket"))
        if self.sync_bn and self.model._is_deep():
            return "one NCCL all-reduce after the backward"
        if self.overlap_allreduce and self.model._image_embedding() is not None:
            return "NCCL, two buckets, the non-embedding one overlapped with the image-embedding backward"
        return "one NCCL all-reduce after the backward"

    # StepLR(step_size, gamma): lr = base * gamma ** (epoch // step_size), stepped once per cycle
    def scheduler_step(self):
        self.epoch += 1
        self.lr = self.base_lr * self.gamma ** (self.epoch // self.step_size)
        if self.comm is not None:
            self.comm.set_lr(self.lr)        # the fused all-reduce + AdamW kernel reads lr from the segment (graph replays too)

    def reduced_gradient(self):
        """The all-reduced SUM of the per-rank gradients of the last step, [n_params] (data-parallel runs)."""
        n = self.model._n_params
        return self._gsum[:n] if self.comm is not None else self.model._grad_flat[:n]

    def _buffers(self, rows):
        s = self._scratch.get(rows)
        if s is None:
            dev = self.model._flat.device
            s = (torch.empty((rows, 1), dtype=torch.float32, device=dev), torch.empty((rows, 1), dtype=torch.float32, device=dev))
            if self.cuda_graph:
                self._scratch[rows] = s      # captured graphs hold these addresses: keep every size alive
            else:
                self._scratch = {rows: s}
        return s

    def train_step(self, x, target, features=None):
        """x: CUDA float32 [B,F,P,P] (None for a 'features_only' ModularTransformer); target: CUDA float32 [B,1] ([B,F,1] for a
        per-frame ModularTransformer); features: [B,feat_dim] (GeneralTransformer) or [B,F,features_dim] (ModularTransformer).
        Enqueues the whole step on the current stream and returns the (device) loss tensor without synchronising.  The returned
        tensor is the trainer's ONE loss buffer (every step overwrites it): read it (`.item()`) or `.clone()` it before the next
        step if the values are collected."""
        model = self.model
        model._ensure_flat()
        if not model.training:
            model.train()
        x, features = model._check_inputs(x, features)
        B, Fr = model._batch_frames(x, features)
        return self._step(x, None, target, features, B, Fr)

    def train_step_from_trajectories(self, src, target, features=None):
        """The same step starting from TRAJECTORIES (BASELINE north_star (1), include/mivit.h: mivit_vit_train_step_traj): `src` is a
        models.TrajectorySource (model.trajectory_source(trajs, nPosPerFrame, center, image_props, seed=..., normalize=...) applies
        the reference's argument handling, or build one around a device-resident float64 [B,T,2] tensor).  Linear / CNN embeddings:
        the frames exist only inside the fused render->embedding kernel and its re-rendering weight-gradient twin; DeepResNet: they
        are rendered into src.frames first."""
        model = self.model
        model._ensure_flat()
        if not model.training:
            model.train()
        if not isinstance(src, TrajectorySource):
            raise TypeError("src must be a models.TrajectorySource (see GeneralTransformer.trajectory_source)")
        features = model._check_features_only(features, src.B, src.F)
        return self._step(None, src, target, features, src.B, src.F)

    def _step(self, x, src, target, features, B, Fr):
        model = self.model
        cfg = model.vit_config(Fr)
        ws = model._workspace(cfg, B)
        rows = B * Fr if cfg.per_frame else B
        target = target.to(device=model._flat.device, dtype=torch.float32).reshape(-1, 1).contiguous()
        if target.shape[0] != rows:
            raise ValueError("target has %d rows, the model predicts %d" % (target.shape[0], rows))
        pred, dpred = self._buffers(rows)
        deep = model._is_deep()
        if src is not None and deep and src.frames is None:
            src.frames = self._frames_buffer(B, Fr, cfg.P)
        self.step_count += 1
        L = _lib.lib()
        overlap = self.overlap_allreduce and self.world > 1 and model._image_embedding() is not None
        if self.comm is not None:
            return self._fused_dp_step(x, src, target, features, cfg, B, ws, pred, dpred, deep, overlap)
        if self.sync_bn and deep:
            return self._sync_bn_step(x, src, target, features, cfg, B, ws, pred, dpred)
        ent = self._graph_entry(x, src, target, features, cfg, B, ws, pred, dpred, deep, overlap) if self.cuda_graph else None
        n, ne = model._n_params, self._n_embedding(cfg)
        grad = model._grad_flat
        if ent is not None:                       # replay the captured forward + loss + backward (one or two graphs)
            graphs, gx, gt, gf, counts, gsrc = ent
            if gx is not None:
                gx.copy_(x)
            if gsrc is not None:                  # trajectories and the global sequence id go through the graph's static buffers
                gsrc.traj.copy_(src.traj)
                gsrc.seq_offset_dev.fill_(src.seq_offset)
            gt.copy_(target)
            if gf is not None:
                gf.copy_(features)
            handles = []
            for i, (g, cnt) in enumerate(zip(graphs, counts)):
                g.replay()
                L.mivit_add_launch_count(cnt)     # the graph's kernel nodes are launches of this library too
                if overlap:                       # bucket 1 (everything but the image embedding) goes out while graph 2 runs
                    part = grad[ne:n] if i == 0 else grad[:ne]
                    handles.append(self.dist.all_reduce(part, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))
        elif overlap:                             # kernel by kernel, same split
            self._forward_loss_backward_part(1, cfg, B, x, src, features, target, ws, pred, dpred, deep)
            handles = [self.dist.all_reduce(grad[ne:n], op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)]
            self._forward_loss_backward_part(2, cfg, B, x, src, features, target, ws, pred, dpred, deep)
            handles.append(self.dist.all_reduce(grad[:ne], op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            handles = []
            self._fused_step_call(cfg, B, x, src, features, target, ws, pred, dpred, deep, int(self.world == 1))
            if self.world == 1:                   # AdamW ran inside the call
                model._gen += 1
                self.last_pred = pred
                return self.loss
        model._gen += 1
        scale = 1.0
        if overlap:
            for h in handles:
                h.wait()                          # the current stream waits for the reductions; the host does not block
            scale = 1.0 / self.world
        elif self.world > 1:
            scale = allreduce_sum_(grad[:n], self.group)
        _lib.check(L.mivit_adamw_step(_lib.ptr(model._flat), _lib.ptr(grad), _lib.ptr(self.m), _lib.ptr(self.v), n, self.lr,
                                      self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count, scale,
                                      _lib.current_stream()))
        self.last_pred = pred
        return self.loss

    # ---- data-parallel step on the peer-memory path: everything is a kernel of this library, so the whole step is one graph ----
    def _fused_dp_step(self, x, src, target, features, cfg, B, ws, pred, dpred, deep, overlap):
        model = self.model
        L = _lib.lib()
        bn_sync = self.sync_bn and deep
        if bn_sync:
            _lib.check(L.mivit_set_bn_sync_comm(self.comm.ref()))
        try:
            ent = self._graph_entry(x, src, target, features, cfg, B, ws, pred, dpred, deep, overlap, fused=True) if self.cuda_graph else None
            if ent is not None:
                graphs, gx, gt, gf, counts, gsrc = ent
                if gx is not None:
                    gx.copy_(x)
                if gsrc is not None:
                    gsrc.traj.copy_(src.traj)
                    gsrc.seq_offset_dev.fill_(src.seq_offset)
                gt.copy_(target)
                if gf is not None:
                    gf.copy_(features)
                graphs[0].replay()
                L.mivit_add_launch_count(counts[0])
            else:
                self._fused_dp_body(cfg, B, x, src, features, target, ws, pred, dpred, deep, overlap)
        finally:
            if bn_sync:
                L.mivit_set_bn_sync_comm(None)
        model._gen += 1
        self.last_pred = pred
        return self.loss

    def _allreduce_adamw(self, bucket, lo, hi, advance, max_ctas=0):
        model = self.model
        _lib.check(_lib.lib().mivit_allreduce_adamw(self.comm.ref(), bucket, lo, hi, _lib.ptr(model._flat), _lib.ptr(self.m),
                                                    _lib.ptr(self.v), self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                                    int(advance), _lib.ptr(self._gsum), int(max_ctas), _lib.current_stream()))

    def _fused_dp_body(self, cfg, B, x, src, features, target, ws, pred, dpred, deep, overlap):
        """forward, MSE, backward and the fused all-reduce + AdamW kernels; with `overlap` the bucket of everything but the image
        embedding goes out on a side stream while the image-embedding backward runs (fork / join by events, capturable)."""
        n = self.model._n_params
        if not overlap:
            self._forward_loss_backward_part(0, cfg, B, x, src, features, target, ws, pred, dpred, deep)
            self._allreduce_adamw(0, 0, n, True)
            return
        ne4 = (self._n_embedding(cfg) + 3) // 4 * 4      # bucket boundary on a 16-byte boundary (the late bucket takes the slack)
        cur = torch.cuda.current_stream()
        self._forward_loss_backward_part(1, cfg, B, x, src, features, target, ws, pred, dpred, deep)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            self._allreduce_adamw(0, ne4, n, False, max_ctas=24)     # shares the GPU with the image-embedding backward
        self._forward_loss_backward_part(2, cfg, B, x, src, features, target, ws, pred, dpred, deep)
        cur.wait_stream(self._side)                       # bucket 0 is complete (and read the step count) before bucket 1 advances it
        self._allreduce_adamw(1, 0, ne4, True)

    def _n_embedding(self, cfg):
        """floats of the image-embedding block at the head of the flat parameter / gradient buffer"""
        if self._ne is None:
            self._ne = int(_lib.lib().mivit_vit_embedding_param_count(ctypes.byref(cfg)))
        return self._ne

    def _frames_buffer(self, B, Fr, P):
        key = ("frames", B, Fr, P)
        buf = self._scratch.get(key)
        if buf is None:
            buf = self._scratch[key] = torch.empty((B, Fr, P, P), dtype=torch.float32, device=self.model._flat.device)
        return buf

    def _fused_step_call(self, cfg, B, x, src, features, target, ws, pred, dpred, deep, apply_update):
        """ONE C-ABI call: forward, MSE, backward (+ AdamW) -- mivit_vit_train_step or its from-trajectories twin."""
        model = self.model
        L = _lib.lib()
        tail = (_lib.ptr(features), _lib.ptr(target), _lib.ptr(model._flat), _lib.ptr(model._grad_flat), _lib.ptr(self.m),
                _lib.ptr(self.v), _lib.ptr(model._bn_flat) if deep else None, _lib.ptr(model._bn_nbt) if deep else None, _lib.ptr(ws),
                _lib.ptr(pred), _lib.ptr(self.loss), _lib.ptr(dpred), self.lr, self.betas[0], self.betas[1], self.eps,
                self.weight_decay, self.step_count, int(apply_update), _lib.current_stream())
        if src is not None:
            return _lib.check(L.mivit_vit_train_step_traj(ctypes.byref(cfg), B, *src.args(), *tail))
        return _lib.check(L.mivit_vit_train_step(ctypes.byref(cfg), B, _lib.ptr(x), *tail))

    def _forward_loss_backward_part(self, part, cfg, B, x, src, features, target, ws, pred, dpred, deep):
        """part 1: forward, MSE, backward of everything but the image embedding; part 2: backward of the image embedding;
        part 0: forward, MSE and the whole backward."""
        model = self.model
        L = _lib.lib()
        st = _lib.current_stream()
        rows = pred.shape[0]
        bn = (_lib.ptr(model._bn_flat) if deep else None, _lib.ptr(model._bn_nbt) if deep else None)
        if part != 2:
            if src is not None:
                _lib.check(L.mivit_vit_forward_traj(ctypes.byref(cfg), B, *src.args(), _lib.ptr(features), _lib.ptr(model._flat),
                                                    bn[0], bn[1], _lib.ptr(ws), _lib.ptr(pred), 1, st))
            else:
                _lib.check(L.mivit_vit_forward(ctypes.byref(cfg), B, _lib.ptr(x), _lib.ptr(features), _lib.ptr(model._flat),
                                               bn[0], bn[1], _lib.ptr(ws), _lib.ptr(pred), 1, st))
            _lib.check(L.mivit_mse_loss(_lib.ptr(pred), _lib.ptr(target), rows, _lib.ptr(self.loss), _lib.ptr(dpred), st))
        if src is not None:
            _lib.check(L.mivit_vit_backward_traj(ctypes.byref(cfg), B, *src.args(), _lib.ptr(features), _lib.ptr(dpred),
                                                 _lib.ptr(model._flat), _lib.ptr(model._grad_flat), _lib.ptr(ws), part, st))
        else:
            _lib.check(L.mivit_vit_backward_part(ctypes.byref(cfg), B, _lib.ptr(x), _lib.ptr(features), _lib.ptr(dpred),
                                                 _lib.ptr(model._flat), _lib.ptr(model._grad_flat), _lib.ptr(ws), part, st))

    def _sync_bn_step(self, x, src, target, features, cfg, B, ws, pred, dpred):
        """Training step with synchronised BatchNorm: the library calls back into `_allreduce_hook` 12 times per step (7 forward
        statistics, 5 backward sums); launched kernel by kernel (the collectives are not captured into a graph)."""
        model = self.model
        L = _lib.lib()
        self._hook_ws, self._hook_error = ws, None
        L.mivit_set_allreduce_hook(self._hook, None, self.world)
        try:
            self._fused_step_call(cfg, B, x, src, features, target, ws, pred, dpred, True, 0)
        finally:
            L.mivit_set_allreduce_hook(_lib.ALLREDUCE_FN(), None, 1)
            if self._hook_error is not None:
                raise self._hook_error
        model._gen += 1
        scale = allreduce_sum_(model._grad_flat[:model._n_params], self.group)
        _lib.check(L.mivit_adamw_step(_lib.ptr(model._flat), _lib.ptr(model._grad_flat), _lib.ptr(self.m), _lib.ptr(self.v),
                                      model._n_params, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                      self.step_count, scale, _lib.current_stream()))
        self.last_pred = pred
        return self.loss

    def _graph_entry(self, x, src, target, features, cfg, B, ws, pred, dpred, deep, overlap, fused=False):
        """CUDA graphs of forward + loss + backward for this batch shape: one graph, or two when the gradient all-reduce is
        overlapped (graph 1 ends where the non-embedding gradients are final, graph 2 is the image-embedding backward).
        Returns None on the first call of a shape: that step runs kernel by kernel -- it also performs every lazy
        initialisation -- and the graphs are captured on the second call."""
        model = self.model
        if self._graph_ws is not ws or self._graph_flat is not model._flat:
            # the model keeps ONE live workspace: a new batch shape replaced it (or the parameters moved), and every captured
            # graph points into the old buffers
            self._graphs.clear()
            self._graph_ws, self._graph_flat = ws, model._flat
        key = (B, None if x is None else tuple(x.shape[1:]), None if features is None else tuple(features.shape[1:]), bool(overlap),
               None if src is None else (src.T, bytes(src.prm), src.seed), pred.shape[0], bool(fused),
               bool(self.sync_bn))   # render scalars are captured by value
        ent = self._graphs.get(key)
        if ent is None:
            self._graphs[key] = "pending"
            return None
        L = _lib.lib()
        if ent == "pending":
            gx, gt = (torch.empty_like(x) if x is not None else None), torch.empty_like(target)
            gf = torch.empty_like(features) if features is not None else None
            gsrc = None
            if src is not None:
                gsrc = TrajectorySource(torch.empty_like(src.traj), src.prm, src.seed, 0,
                                        torch.zeros(1, dtype=torch.int64, device=src.traj.device), src.frames)
            graphs, counts = [], []
            cap = torch.cuda.Stream()
            cap.wait_stream(torch.cuda.current_stream())
            if fused:       # the whole data-parallel step, collectives included, is one graph
                n0 = L.mivit_launch_count()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap):
                    self._fused_dp_body(cfg, B, gx, gsrc, gf, gt, ws, pred, dpred, deep, overlap)
                n_cap = int(L.mivit_launch_count() - n0)
                L.mivit_add_launch_count(-n_cap)
                graphs.append(g)
                counts.append(n_cap)
            for part in (() if fused else (1, 2) if overlap else (0,)):
                n0 = L.mivit_launch_count()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap):
                    self._forward_loss_backward_part(part, cfg, B, gx, gsrc, gf, gt, ws, pred, dpred, deep)
                n_cap = int(L.mivit_launch_count() - n0)
                L.mivit_add_launch_count(-n_cap)          # capturing enqueued nothing: only replays count
                graphs.append(g)
                counts.append(n_cap)
            ent = self._graphs[key] = (graphs, gx, gt, gf, counts, gsrc)
        return ent
