"""Drop-in mirror of the ViT part of the reference's helpers/models.py:
same class names, constructor signatures, sub-module / state_dict key names and initialisation
order (so `torch.manual_seed(s)` gives the same random-init weights as the reference), but the
arithmetic of GeneralTransformer.forward / backward runs in the CUDA library (include/mivit.h:
mivit_vit_forward / mivit_vit_backward / mivit_vit_train_step).  There is no PyTorch fallback.

  MultiHeadAttention               helpers/models.py:11-59
  FeedForward                      :61-77
  TransformerEncoderLayerWithSkip  :81-108
  Transformer                      :111-141
  LinearProjectionEmbedding        :146-167
  CNNEmbedding                     :170-199
  ResidualBlock / DeepResNetEmbedding  :202-257
  MLPHead                          :260-276
  GeneralTransformer               :278-361
  ModularTransformer               :366-593  (per-frame features: 'images_only' | 'features_only' | 'both' with
                                              'add' | 'concat_proj' | 'concat_features' fusion)
  ImageDataset / ImageFeatureDataset   :781-803

Not supported (raise at construction): dropout > 0 (every reference experiment uses 0.0), MLPHead activations other
than nn.ReLU.  `single_prediction=False` is stored and ignored by GeneralTransformer exactly like the reference does
(helpers/models.py:303, forward :328-361); ModularTransformer without a regression token then returns per-frame
predictions [B, F, 1] (:585-593).

Additions (no reference counterpart as ONE call): `forward_from_trajectories` -- the training loops'
`model(torch.Tensor(normalize_images(trajectories_to_video(trajs, ...))))` with the renderer fused into the frame embedding.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.data import Dataset

from . import _lib

MAX_TOKENS = 128  # helpers/models.py:8

__all__ = ["MAX_TOKENS", "MultiHeadAttention", "FeedForward", "TransformerEncoderLayerWithSkip", "Transformer",
           "LinearProjectionEmbedding", "CNNEmbedding", "ResidualBlock", "DeepResNetEmbedding", "MLPHead",
           "GeneralTransformer", "ModularTransformer", "ImageDataset", "ImageFeatureDataset", "VitConfig", "TrajectorySource"]


class VitConfig(ctypes.Structure):
    """struct mivit_vit_config (include/mivit.h)."""
    _fields_ = [("embedding", ctypes.c_int32), ("P", ctypes.c_int32), ("F", ctypes.c_int32), ("E", ctypes.c_int32),
                ("H", ctypes.c_int32), ("HD", ctypes.c_int32), ("L", ctypes.c_int32), ("activation", ctypes.c_int32),
                ("use_pos", ctypes.c_int32), ("use_reg", ctypes.c_int32), ("use_feat", ctypes.c_int32),
                ("fusion", ctypes.c_int32), ("feat_dim", ctypes.c_int32), ("head_hidden", ctypes.c_int32),
                ("conv_impl", ctypes.c_int32), ("bn_eps", ctypes.c_float), ("bn_momentum", ctypes.c_float),
                ("ln_eps", ctypes.c_float), ("modular", ctypes.c_int32), ("mod_mode", ctypes.c_int32),
                ("mod_fembed", ctypes.c_int32), ("mod_fusion", ctypes.c_int32), ("per_frame", ctypes.c_int32)]


def _no_direct_forward(self, *a, **k):
    raise NotImplementedError(
        "%s is a parameter container in moleculardiffusion_mivit_b200: its arithmetic runs inside "
        "GeneralTransformer.forward (CUDA library); call the GeneralTransformer." % type(self).__name__)


class MultiHeadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        self.dropout = nn.Dropout(dropout)
        nn.init.xavier_uniform_(self.q_proj.weight)
        nn.init.xavier_uniform_(self.k_proj.weight)
        nn.init.xavier_uniform_(self.v_proj.weight)
        nn.init.xavier_uniform_(self.out_proj.weight)

    forward = _no_direct_forward


class FeedForward(nn.Module):
    def __init__(self, embed_dim, hidden_dim, activation_fct, dropout=0.0):
        super().__init__()
        self.fc1 = nn.Linear(embed_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, embed_dim)
        self.dropout = nn.Dropout(dropout)
        if not callable(activation_fct):
            raise ValueError("activation_fct must be a callable function from torch.nn.functional or a custom function.")
        self.activation = activation_fct

    forward = _no_direct_forward


class TransformerEncoderLayerWithSkip(nn.Module):
    def __init__(self, embed_dim, num_heads, hidden_dim, activation_fct, dropout=0.0):
        super().__init__()
        self.self_attn = MultiHeadAttention(embed_dim, num_heads, dropout)
        self.norm1 = nn.LayerNorm(embed_dim)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.feed_forward = FeedForward(embed_dim, hidden_dim, activation_fct, dropout)
        self.dropout = nn.Dropout(dropout)

    forward = _no_direct_forward


class Transformer(nn.Module):
    def __init__(self, embed_dim, num_heads, hidden_dim, num_layers, dropout, use_pos_encoding, activation_fct):
        super().__init__()
        self.embed_dim = embed_dim
        self.use_pos_encoding = use_pos_encoding
        if self.use_pos_encoding:
            self.pos_embedding = nn.Parameter(torch.randn(1, MAX_TOKENS, embed_dim))
        self.encoder_layers = nn.ModuleList([
            TransformerEncoderLayerWithSkip(embed_dim=embed_dim, num_heads=num_heads, hidden_dim=hidden_dim,
                                            activation_fct=activation_fct, dropout=dropout)
            for _ in range(num_layers)])
        self.norm = nn.LayerNorm(embed_dim)

    forward = _no_direct_forward


class LinearProjectionEmbedding(nn.Module):
    def __init__(self, patch_size, embed_dim):
        super().__init__()
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.proj = nn.Linear(patch_size * patch_size, embed_dim)

    forward = _no_direct_forward


class CNNEmbedding(nn.Module):
    def __init__(self, patch_size, embed_dim):
        super().__init__()
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.conv = nn.Conv2d(in_channels=1, out_channels=embed_dim, kernel_size=(patch_size, patch_size))

    forward = _no_direct_forward


class ResidualBlock(nn.Module):
    def __init__(self, in_channels, out_channels, downsample=False):
        super().__init__()
        if downsample:
            raise NotImplementedError("downsample=True is not used by DeepResNetEmbedding and is not supported")
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1, stride=1, bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1, stride=1, bias=False)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.skip = nn.Sequential()
        if in_channels != out_channels or downsample:
            self.skip = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, bias=False),
                                      nn.BatchNorm2d(out_channels))

    forward = _no_direct_forward


class DeepResNetEmbedding(nn.Module):
    def __init__(self, patch_size=7, embed_dim=128):
        super().__init__()
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.initial_conv = nn.Conv2d(1, 32, kernel_size=3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        self.relu = nn.ReLU(inplace=True)
        self.res_block1 = ResidualBlock(32, 64)
        self.res_block2 = ResidualBlock(64, 128)
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(128, embed_dim)

    forward = _no_direct_forward


class MLPHead(nn.Module):
    def __init__(self, input_dim, hidden_dim=128, output_dim=1, dropout=0.0, activation=nn.ReLU):
        super().__init__()
        if dropout > 0 or output_dim != 1 or activation is not nn.ReLU:
            raise NotImplementedError("MLPHead: only dropout=0, output_dim=1, activation=nn.ReLU run on the CUDA path")
        self.mlp = nn.Sequential(nn.Linear(input_dim, hidden_dim), activation(),
                                 nn.Dropout(dropout) if dropout > 0 else nn.Identity(),
                                 nn.Linear(hidden_dim, output_dim))

    forward = _no_direct_forward


_ACTIVATIONS = {F.relu: 0, F.gelu: 1, F.leaky_relu: 2}
_EMBEDDINGS = {LinearProjectionEmbedding: 0, CNNEmbedding: 1, DeepResNetEmbedding: 2}
_BN_CHANNELS = (32, 64, 64, 64, 128, 128, 128)


class _VitFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, features, src, *params):
        if not model.training and model._is_deep():
            # mivit_vit_backward implements the BATCH-statistics BatchNorm backward only; with running statistics
            # (eval mode) its gradients would be wrong, so refuse instead of returning them silently
            raise NotImplementedError(
                "backward through a DeepResNetEmbedding model in eval() mode is not implemented on the CUDA path "
                "(BatchNorm backward uses batch statistics): call model.train(), or wrap the call in torch.no_grad()")
        pred, gen = model._run_forward(x, features, model.training, src)
        ctx.model, ctx.gen, ctx.src = model, gen, src
        ctx.has_x, ctx.has_feat = x is not None, features is not None
        ctx.save_for_backward(*[t for t in (x, features) if t is not None])
        return pred

    @staticmethod
    def backward(ctx, dpred):
        saved = list(ctx.saved_tensors)
        x = saved.pop(0) if ctx.has_x else None
        feats = saved.pop(0) if ctx.has_feat else None
        grads = ctx.model._run_backward(x, feats, dpred.contiguous(), ctx.gen, ctx.src)
        return (None, None, None, None) + tuple(grads)


class TrajectorySource:
    """Image source of the from-trajectories entry points (include/mivit.h: mivit_vit_*_traj): device float64 trajectories
    [B,T,2] (y already flipped), the renderer's scalar parameters, the Philox seed and the global id of sequence 0, an optional
    DEVICE uint64 counter added to that id (graph replay), and -- DeepResNet only -- the [B,F,P,P] frame buffer."""

    def __init__(self, traj, prm, seed, seq_offset=0, seq_offset_dev=None, frames=None):
        self.traj, self.prm, self.seed, self.seq_offset = traj, prm, int(seed), int(seq_offset)
        self.seq_offset_dev, self.frames = seq_offset_dev, frames
        self.B, self.T = int(traj.shape[0]), int(traj.shape[1])
        self.F = self.T // int(prm.n)

    def args(self):
        return (_lib.ptr(self.traj), self.T, ctypes.byref(self.prm), self.seed, self.seq_offset, _lib.ptr(self.seq_offset_dev),
                _lib.ptr(self.frames))


def _embedding_keys(prefix, emb):
    """state_dict keys of an embedding module's parameters in the flat-buffer order of mivit_vit_config."""
    kind = _EMBEDDINGS[type(emb)]
    if kind == 2:
        k = ["initial_conv.weight", "bn1.weight", "bn1.bias"]
        for b in ("res_block1", "res_block2"):
            k += ["%s.%s" % (b, s) for s in ("conv1.weight", "bn1.weight", "bn1.bias", "conv2.weight", "bn2.weight", "bn2.bias",
                                             "skip.0.weight", "skip.1.weight", "skip.1.bias")]
        k += ["fc.weight", "fc.bias"]
    elif kind == 0:
        k = ["proj.weight", "proj.bias"]
    else:
        k = ["conv.weight", "conv.bias"]
    return [prefix + s for s in k]


def _transformer_keys(num_layers, use_reg, use_pos):
    k = ["norm.weight", "norm.bias"]
    if use_reg:
        k.append("reg_token")
    if use_pos:
        k.append("transformer.pos_embedding")
    for i in range(num_layers):
        p = "transformer.encoder_layers.%d." % i
        for s in ("q_proj", "k_proj", "v_proj", "out_proj"):
            k += [p + "self_attn.%s.weight" % s, p + "self_attn.%s.bias" % s]
        k += [p + "norm1.weight", p + "norm1.bias", p + "feed_forward.fc1.weight", p + "feed_forward.fc1.bias",
              p + "feed_forward.fc2.weight", p + "feed_forward.fc2.bias", p + "norm2.weight", p + "norm2.bias"]
    return k + ["transformer.norm.weight", "transformer.norm.bias"]


class _CudaViT(nn.Module):
    """Plumbing shared by GeneralTransformer and ModularTransformer: ONE flat fp32 CUDA buffer behind all Parameters (canonical
    order = param_keys()), the BatchNorm running-statistics buffer, the workspace and the forward / backward calls into the C
    library.  Subclasses provide param_keys(), vit_config(n_frames), _image_embedding() and _check_inputs()."""

    def _init_cuda_state(self):
        self._flat = self._grad_flat = self._bn_flat = self._bn_nbt = None
        self._ws, self._gen = {}, 0
        self.conv_impl = 1   # 1 = tcgen05 convolutions; 0 = SIMT cross-check kernels (tests only)

    def _is_deep(self):
        emb = self._image_embedding()
        return emb is not None and _EMBEDDINGS[type(emb)] == 2

    def _bn_modules(self):
        e = self._image_embedding()
        return [e.bn1, e.res_block1.bn1, e.res_block1.bn2, e.res_block1.skip[1],
                e.res_block2.bn1, e.res_block2.bn2, e.res_block2.skip[1]]

    # ------------------------------------------------------------------ flat buffers --------
    def _ensure_flat(self):
        """(Re)build the flat fp32 CUDA parameter buffer and re-point every Parameter / BatchNorm buffer at a
        view of it.  Parameter objects keep their identity, so optimizers created earlier stay valid."""
        dev = _lib.require_cuda()
        named = dict(self.named_parameters())
        keys = self.param_keys()
        assert set(keys) == set(named), "parameter set does not match the canonical layout"
        ok = self._flat is not None and self._flat.device == dev
        if ok:
            off = 0
            for k in keys:
                p = named[k]
                if p.data_ptr() != self._flat.data_ptr() + 4 * off or p.dtype != torch.float32:
                    ok = False
                    break
                off += p.numel()
        if not ok:
            cfg = self.vit_config(1)
            n = _lib.lib().mivit_vit_param_count(ctypes.byref(cfg))
            if n < 0:
                raise _lib.MivitError(_lib.lib().mivit_last_error().decode())
            sizes = (ctypes.c_int64 * n)()
            _lib.check(_lib.lib().mivit_vit_param_sizes(ctypes.byref(cfg), sizes, n))
            assert n == len(keys) and [int(s) for s in sizes] == [named[k].numel() for k in keys], \
                "C-ABI parameter layout does not match the module"
            total = sum(int(s) for s in sizes)
            flat = torch.empty(total + 4, dtype=torch.float32, device=dev)
            off = 0
            for k in keys:
                p = named[k]
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data.to(device=dev, dtype=torch.float32))
                p.data = view
                off += p.numel()
            self._flat, self._n_params = flat, total
            self._grad_flat = torch.zeros(total + 4, dtype=torch.float32, device=dev)
        if self._is_deep():
            bns = self._bn_modules()
            okb = self._bn_flat is not None and self._bn_flat.device == dev
            if okb:
                off = 0
                for bn in bns:
                    C = bn.num_features
                    if bn.running_mean.data_ptr() != self._bn_flat.data_ptr() + 4 * off:
                        okb = False
                        break
                    off += 2 * C
            if not okb:
                flat = torch.empty(2 * sum(_BN_CHANNELS), dtype=torch.float32, device=dev)
                nbt = torch.empty(7, dtype=torch.int64, device=dev)
                off = 0
                for i, bn in enumerate(bns):
                    C = bn.num_features
                    for name, o in (("running_mean", off), ("running_var", off + C)):
                        buf = getattr(bn, name)
                        view = flat[o:o + C]
                        view.copy_(buf.data.to(device=dev, dtype=torch.float32))
                        buf.data = view
                    nv = nbt[i:i + 1].view(())
                    nv.copy_(bn.num_batches_tracked.data.to(dev))
                    bn.num_batches_tracked.data = nv
                    off += 2 * C
                self._bn_flat, self._bn_nbt = flat, nbt
        return dev

    def _workspace(self, cfg, B):
        key = (B, cfg.F, cfg.conv_impl)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = _lib.lib().mivit_vit_workspace_bytes(ctypes.byref(cfg), B)
            if nbytes < 0:
                raise _lib.MivitError(_lib.lib().mivit_last_error().decode())
            self._ws.clear()       # one live workspace: activations of the last forward
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self._flat.device)
            self._ws[key] = ws
        return ws

    @staticmethod
    def _batch_frames(x, features):
        t = x if x is not None else features
        return int(t.shape[0]), int(t.shape[1])

    def _pred_shape(self, cfg, B):
        """[B,1], or [B,F,1] for per-frame outputs (ModularTransformer, single_prediction=False, no regression token)."""
        return (B, cfg.F, 1) if cfg.per_frame else (B, 1)

    def _run_forward(self, x, features, training, src=None):
        B, Fr = (src.B, src.F) if src is not None else self._batch_frames(x, features)
        cfg = self.vit_config(Fr)
        ws = self._workspace(cfg, B)
        pred = torch.empty(self._pred_shape(cfg, B), dtype=torch.float32, device=self._flat.device)
        is_deep = self._is_deep()
        bn = (_lib.ptr(self._bn_flat) if is_deep else None, _lib.ptr(self._bn_nbt) if is_deep else None)
        if src is not None:
            if is_deep and src.frames is None:
                src.frames = torch.empty((B, Fr, cfg.P, cfg.P), dtype=torch.float32, device=self._flat.device)
            _lib.check(_lib.lib().mivit_vit_forward_traj(
                ctypes.byref(cfg), B, *src.args(), _lib.ptr(features), _lib.ptr(self._flat), bn[0], bn[1],
                _lib.ptr(ws), _lib.ptr(pred), int(bool(training)), _lib.current_stream()))
        else:
            _lib.check(_lib.lib().mivit_vit_forward(
                ctypes.byref(cfg), B, _lib.ptr(x), _lib.ptr(features), _lib.ptr(self._flat), bn[0], bn[1],
                _lib.ptr(ws), _lib.ptr(pred), int(bool(training)), _lib.current_stream()))
        self._gen += 1
        return pred, self._gen

    def _run_backward(self, x, features, dpred, gen, src=None):
        if gen != self._gen:
            raise RuntimeError("backward() must follow the forward() that produced the output (one live workspace per model)")
        B, Fr = (src.B, src.F) if src is not None else self._batch_frames(x, features)
        cfg = self.vit_config(Fr)
        ws = self._workspace(cfg, B)
        if src is not None:
            _lib.check(_lib.lib().mivit_vit_backward_traj(ctypes.byref(cfg), B, *src.args(), _lib.ptr(features), _lib.ptr(dpred),
                                                          _lib.ptr(self._flat), _lib.ptr(self._grad_flat), _lib.ptr(ws), 0,
                                                          _lib.current_stream()))
        else:
            _lib.check(_lib.lib().mivit_vit_backward(ctypes.byref(cfg), B, _lib.ptr(x), _lib.ptr(features), _lib.ptr(dpred),
                                                     _lib.ptr(self._flat), _lib.ptr(self._grad_flat), _lib.ptr(ws),
                                                     _lib.current_stream()))
        named = dict(self.named_parameters())
        grads, off = [], 0
        for k in self.param_keys():
            p = named[k]
            grads.append(self._grad_flat[off:off + p.numel()].view(p.shape).clone())
            off += p.numel()
        return grads

    def _forward_cuda(self, x, features, src=None):
        named = dict(self.named_parameters())
        params = [named[k] for k in self.param_keys()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _VitFunction.apply(self, x, features, src, *params)
        pred, _ = self._run_forward(x, features, self.training, src)
        return pred

    def trajectory_source(self, trajectories, nPosPerFrame, center=False, image_props={}, *, seed=None, seq_offset=0,
                          normalize=None, seq_offset_dev=None, _mean_noise=False):
        """Argument handling of trajectories_to_video (helpers/helpersGeneration.py:194-247: in-place y flip of the caller's
        array, the T % nPosPerFrame check after it, image_props defaults) -> TrajectorySource on this model's device.
        normalize=(background_mean, background_sigma, theoretical_max) fuses normalize_images (:389-395)."""
        from . import helpersGeneration as _g
        dev = self._ensure_flat()
        trajectories[:, :, 1] *= -1
        if trajectories.shape[1] % nPosPerFrame != 0:
            raise Exception("T is not divisble by posPerFrame")
        prm = _g.derive_render_params(image_props, nPosPerFrame, center, "v1")
        prm.flip_y = 0                      # already applied to the caller's array
        prm.mean_noise = int(bool(_mean_noise))
        if normalize is not None:
            m, sg, mx = normalize
            den = mx - (m - sg)
            if den == 0:
                raise ValueError("Denominator in normalization is zero. Check your inputs.")
            prm.normalize, prm.norm_sub, prm.norm_div = 1, float(m - sg), float(den)
        emb = self._image_embedding()
        if emb is None:
            raise ValueError("this model has no image embedding to render for")
        if prm.P != emb.patch_size:
            raise AssertionError("Patch size mismatch")
        t_dev, _ = _g._to_device_f64(trajectories, dev)
        return TrajectorySource(t_dev, prm, _g._draw_seed(seed), seq_offset, seq_offset_dev)

    def forward_from_trajectories(self, trajectories, nPosPerFrame, center=False, image_props={}, features=None, *, seed=None,
                                  seq_offset=0, normalize=None, _mean_noise=False):
        """`self(torch.Tensor(normalize_images(trajectories_to_video(trajectories, nPosPerFrame, center, image_props))), features)`
        of the reference training loops (Experiments/Embeddings/trainModelsEmbeddings.py:150-186) as ONE call into the CUDA
        library: for LinearProjectionEmbedding / CNNEmbedding the renderer is fused with the frame embedding and the frames never
        reach HBM, forward or backward (the weight gradient re-renders them from the same Philox streams); DeepResNetEmbedding
        renders into a frame buffer first (BatchNorm needs all frames).  Same side effect (in-place y flip), errors and noise
        streams as trajectories_to_video(..., seed=seed, seq_offset=seq_offset, normalize=normalize).  Differentiable w.r.t. the
        parameters."""
        self._ensure_flat()
        src = self.trajectory_source(trajectories, nPosPerFrame, center, image_props, seed=seed, seq_offset=seq_offset,
                                     normalize=normalize, _mean_noise=_mean_noise)
        features = self._check_features_only(features, src.B, src.F)
        return self._forward_cuda(None, features, src)

    # ------------------------------------------------------------------ fused step ----------
    def flat_parameters(self):
        self._ensure_flat()
        return self._flat[:self._n_params]

    def flat_gradients(self):
        self._ensure_flat()
        return self._grad_flat[:self._n_params]


class GeneralTransformer(_CudaViT):
    def __init__(self, embedding_cls, embed_kwargs, embed_dim, num_heads, hidden_dim, num_layers, mlp_head,
                 tr_activation_fct, dropout=0, use_pos_encoding=False, use_regression_token=False,
                 single_prediction=True, use_global_features=False, fusion_type='early', global_feature_dim=None):
        super().__init__()
        if dropout != 0:
            raise NotImplementedError("dropout > 0 is not implemented on the CUDA path (the reference experiments use 0.0)")
        if embedding_cls not in _EMBEDDINGS:
            raise ValueError("embedding_cls must be LinearProjectionEmbedding, CNNEmbedding or DeepResNetEmbedding")
        if tr_activation_fct not in _ACTIVATIONS:
            raise ValueError("tr_activation_fct must be F.relu, F.gelu or F.leaky_relu")
        self.embed_dim = embed_dim
        self.embedding = embedding_cls(**embed_kwargs)
        self.norm = nn.LayerNorm(embed_dim)
        self.use_regression_token = use_regression_token
        self.single_prediction = single_prediction
        self.use_global_features = use_global_features
        self.fusion_type = fusion_type
        if use_regression_token:
            self.reg_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.transformer = Transformer(embed_dim, num_heads, hidden_dim, num_layers, dropout,
                                       use_pos_encoding=use_pos_encoding, activation_fct=tr_activation_fct)
        if use_global_features:
            assert global_feature_dim is not None, "Must provide global_feature_dim if using global features"
            self.feature_projector = nn.Sequential(nn.Linear(global_feature_dim, embed_dim), nn.ReLU(),
                                                   nn.Linear(embed_dim, embed_dim))
        if fusion_type == 'late' and use_global_features:
            self.mlp_head = mlp_head(input_dim=embed_dim * 2)
        else:
            self.mlp_head = mlp_head(input_dim=embed_dim)
        # ---- CUDA-path bookkeeping (not part of the reference surface)
        self._num_heads, self._hidden_dim, self._num_layers = num_heads, hidden_dim, num_layers
        self._activation = _ACTIVATIONS[tr_activation_fct]
        self._use_pos = bool(use_pos_encoding)
        self._feat_dim = int(global_feature_dim) if use_global_features else 0
        self._head_hidden = self.mlp_head.mlp[0].out_features
        self._init_cuda_state()

    def _image_embedding(self):
        return self.embedding

    # ------------------------------------------------------------------ canonical order -----
    def param_keys(self):
        """state_dict keys of the parameters in the flat-buffer order of mivit_vit_config."""
        k = _embedding_keys("embedding.", self.embedding)
        k += _transformer_keys(self._num_layers, self.use_regression_token, self._use_pos)
        if self.use_global_features:
            k += ["feature_projector.0.weight", "feature_projector.0.bias", "feature_projector.2.weight",
                  "feature_projector.2.bias"]
        k += ["mlp_head.mlp.0.weight", "mlp_head.mlp.0.bias", "mlp_head.mlp.3.weight", "mlp_head.mlp.3.bias"]
        return k

    def vit_config(self, n_frames):
        c = VitConfig()
        c.embedding = _EMBEDDINGS[type(self.embedding)]
        c.P, c.F, c.E = int(self.embedding.patch_size), int(n_frames), int(self.embed_dim)
        c.H, c.HD, c.L = int(self._num_heads), int(self._hidden_dim), int(self._num_layers)
        c.activation, c.use_pos, c.use_reg = self._activation, int(self._use_pos), int(self.use_regression_token)
        c.use_feat, c.fusion, c.feat_dim = int(self.use_global_features), int(self.fusion_type == 'late'), self._feat_dim
        c.head_hidden, c.conv_impl = int(self._head_hidden), int(self.conv_impl)
        c.bn_eps, c.bn_momentum, c.ln_eps = 1e-5, 0.1, 1e-5
        return c

    def _check_features_only(self, features, B, Fr):
        if self.use_global_features:
            assert features is not None, "Global features required for %s fusion" % self.fusion_type
            return features.to(device=self._flat.device, dtype=torch.float32).contiguous()
        return None

    def _check_inputs(self, x, features):
        if x.dim() != 4 or x.shape[2] != self.embedding.patch_size or x.shape[3] != self.embedding.patch_size:
            raise AssertionError("Patch size mismatch")
        x = x.to(device=self._flat.device, dtype=torch.float32).contiguous()
        return x, self._check_features_only(features, x.shape[0], x.shape[1])

    def forward(self, x, features=None):
        """x: [batch_size, num_images, image_size, image_size]; features: [batch_size, num_features] or None.
        Returns [batch_size, 1] (CUDA tensor).  Gradients flow to the parameters, not to x."""
        self._ensure_flat()
        x, features = self._check_inputs(x, features)
        return self._forward_cuda(x, features)


class ModularTransformer(_CudaViT):
    """helpers/models.py:366-593.  `features` are PER FRAME, [batch_size, num_images, features_dim]; `mlp_head` is a module
    INSTANCE here (the reference's GeneralTransformer takes the class).  Same constructor errors as the reference."""

    def __init__(self, embed_dim, num_heads, hidden_dim, num_layers, mlp_head, tr_activation_fct, dropout=0,
                 use_pos_encoding=False, use_regression_token=False, single_prediction=True, mode='images_only',
                 image_embedding_cls=None, image_embed_kwargs=None, features_dim=None, feature_embedding_type='linear',
                 fusion_method='add'):
        super().__init__()
        if dropout != 0:
            raise NotImplementedError("dropout > 0 is not implemented on the CUDA path (the reference experiments use 0.0)")
        if tr_activation_fct not in _ACTIVATIONS:
            raise ValueError("tr_activation_fct must be F.relu, F.gelu or F.leaky_relu")
        self.embed_dim = embed_dim
        self.mode = mode
        self.use_regression_token = use_regression_token
        self.single_prediction = single_prediction
        self.fusion_method = fusion_method
        self.features_dim = features_dim
        if mode not in ['images_only', 'features_only', 'both']:
            raise ValueError("mode must be one of: 'images_only', 'features_only', 'both'")
        if mode == 'both' and fusion_method not in ['add', 'concat_proj', 'concat_features']:
            raise ValueError("fusion_method must be one of: 'add', 'concat_proj', 'concat_features'")
        if mode == 'both' and fusion_method == 'concat_features':
            image_embed_dim = embed_dim - features_dim
            if image_embed_dim <= 0:
                raise ValueError(f"embed_dim ({embed_dim}) must be greater than features_dim ({features_dim}) when using "
                                 f"'concat_features' fusion")
            if image_embed_kwargs is None:
                image_embed_kwargs = {}
            image_embed_kwargs['embed_dim'] = image_embed_dim      # mutates the caller's dict, like the reference (:427)
        self.image_embedding = None
        if mode in ['images_only', 'both']:
            if image_embedding_cls is None:
                raise ValueError("image_embedding_cls must be provided when using images")
            if image_embedding_cls not in _EMBEDDINGS:
                raise ValueError("image_embedding_cls must be LinearProjectionEmbedding, CNNEmbedding or DeepResNetEmbedding")
            if image_embed_kwargs is None:
                image_embed_kwargs = {}
            self.image_embedding = image_embedding_cls(**image_embed_kwargs)
        self.feature_embedding = None
        if mode in ['features_only', 'both'] and fusion_method != 'concat_features':
            if features_dim is None:
                raise ValueError("features_dim must be provided when using features")
            if feature_embedding_type == 'linear':
                self.feature_embedding = nn.Linear(features_dim, embed_dim)
            elif feature_embedding_type == 'mlp':
                self.feature_embedding = nn.Sequential(nn.Linear(features_dim, embed_dim * 2), nn.LayerNorm(embed_dim * 2),
                                                       nn.GELU(), nn.Linear(embed_dim * 2, embed_dim))
            else:
                raise ValueError(f"Unknown feature_embedding_type: {feature_embedding_type}")
        self.fusion_layer = None
        if mode == 'both' and fusion_method == 'concat_proj':
            self.fusion_layer = nn.Linear(embed_dim * 2, embed_dim)
        self.norm = nn.LayerNorm(embed_dim)
        if use_regression_token:
            self.reg_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.transformer = Transformer(embed_dim, num_heads, hidden_dim, num_layers, dropout,
                                       use_pos_encoding=use_pos_encoding, activation_fct=tr_activation_fct)
        if not isinstance(mlp_head, MLPHead):
            raise NotImplementedError("ModularTransformer: mlp_head must be an MLPHead instance on the CUDA path")
        self.mlp_head = mlp_head
        # ---- CUDA-path bookkeeping (not part of the reference surface)
        self._num_heads, self._hidden_dim, self._num_layers = num_heads, hidden_dim, num_layers
        self._activation = _ACTIVATIONS[tr_activation_fct]
        self._use_pos = bool(use_pos_encoding)
        self._fembed_mlp = feature_embedding_type == 'mlp'
        self._head_hidden = self.mlp_head.mlp[0].out_features
        self._init_cuda_state()

    def _image_embedding(self):
        return self.image_embedding

    def param_keys(self):
        k = _embedding_keys("image_embedding.", self.image_embedding) if self.image_embedding is not None else []
        if self.feature_embedding is not None:
            if self._fembed_mlp:
                k += ["feature_embedding.%s" % s for s in ("0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias")]
            else:
                k += ["feature_embedding.weight", "feature_embedding.bias"]
        if self.fusion_layer is not None:
            k += ["fusion_layer.weight", "fusion_layer.bias"]
        k += _transformer_keys(self._num_layers, self.use_regression_token, self._use_pos)
        k += ["mlp_head.mlp.0.weight", "mlp_head.mlp.0.bias", "mlp_head.mlp.3.weight", "mlp_head.mlp.3.bias"]
        return k

    def vit_config(self, n_frames):
        c = VitConfig()
        emb = self.image_embedding
        c.embedding = _EMBEDDINGS[type(emb)] if emb is not None else 0
        c.P, c.F, c.E = int(emb.patch_size) if emb is not None else 1, int(n_frames), int(self.embed_dim)
        c.H, c.HD, c.L = int(self._num_heads), int(self._hidden_dim), int(self._num_layers)
        c.activation, c.use_pos, c.use_reg = self._activation, int(self._use_pos), int(self.use_regression_token)
        c.use_feat, c.fusion = 0, 0
        c.feat_dim = int(self.features_dim) if (self.mode != 'images_only' and self.features_dim is not None) else 0
        c.head_hidden, c.conv_impl = int(self._head_hidden), int(self.conv_impl)
        c.bn_eps, c.bn_momentum, c.ln_eps = 1e-5, 0.1, 1e-5
        c.modular = 1
        c.mod_mode = {'images_only': 0, 'features_only': 1, 'both': 2}[self.mode]
        c.mod_fembed = int(self._fembed_mlp)
        c.mod_fusion = {'add': 0, 'concat_proj': 1, 'concat_features': 2}.get(self.fusion_method, 0)
        # helpers/models.py:585-593: without a regression token and with single_prediction=False the head sees every token
        c.per_frame = int(not self.use_regression_token and not self.single_prediction)
        return c

    def _check_features_only(self, features, B, Fr):
        if self.mode == 'images_only':
            return None
        if features is None:
            raise ValueError("Both images and features are required for 'both' mode")
        if features.dim() != 3 or features.shape[2] != self.features_dim:
            raise ValueError("features must be [batch_size, num_images, features_dim]")
        if features.shape[0] != B or features.shape[1] != Fr:
            raise ValueError("Images and features must have the same batch size and sequence length")
        return features.to(device=self._flat.device, dtype=torch.float32).contiguous()

    def _check_inputs(self, images, features):
        if self.mode == 'images_only' and images is None:
            raise ValueError("Images are required for 'images_only' mode")
        if self.mode == 'features_only' and features is None:
            raise ValueError("Features are required for 'features_only' mode")
        if self.mode == 'both' and (images is None or features is None):
            raise ValueError("Both images and features are required for 'both' mode")
        dev = self._flat.device
        if self.mode == 'features_only':
            images = None
        else:
            ps = self.image_embedding.patch_size
            if images.dim() != 4 or images.shape[2] != ps or images.shape[3] != ps:
                raise AssertionError("Patch size mismatch")
            images = images.to(device=dev, dtype=torch.float32).contiguous()
        if self.mode == 'images_only':
            features = None
        else:
            if self.mode == 'both' and (images.shape[0] != features.shape[0] or images.shape[1] != features.shape[1]):
                raise ValueError("Images and features must have the same batch size and sequence length")
            if features.dim() != 3 or features.shape[2] != self.features_dim:
                raise ValueError("features must be [batch_size, num_images, features_dim]")
            features = features.to(device=dev, dtype=torch.float32).contiguous()
        return images, features

    def forward(self, images=None, features=None):
        """images: [batch_size, num_images, image_size, image_size] or None; features: [batch_size, num_images, features_dim]
        or None (NaNs are replaced by zeros).  Returns [batch_size, 1], or [batch_size, num_images, 1] for
        use_regression_token=False, single_prediction=False."""
        self._ensure_flat()
        images, features = self._check_inputs(images, features)
        return self._forward_cuda(images, features)


class ImageDataset(Dataset):  # helpers/models.py:781-790
    def __init__(self, images, labels):
        self.images = images
        self.labels = labels

    def __len__(self):
        return len(self.images)

    def __getitem__(self, idx):
        return self.images[idx], self.labels[idx]


class ImageFeatureDataset(Dataset):  # helpers/models.py:793-803
    def __init__(self, images, features, labels):
        self.images = images
        self.features = features
        self.labels = labels

    def __len__(self):
        return len(self.images)

    def __getitem__(self, idx):
        return self.images[idx], self.features[idx], self.labels[idx]
