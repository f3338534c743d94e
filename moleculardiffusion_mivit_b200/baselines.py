"""Drop-in mirror of the CNN baselines of the reference's helpers/models.py -- the ResNet every
experiment trains beside its ViTs (Experiments/PSFNoise/trainSettingsPSFNoise.py:114, trainSettingsImagesFeatures.py:170-173):

  BasicBlock                  helpers/models.py:600-635
  LightResNet                 :638-683
  MultiImageResNet            :686-701
  LightImagesFeaturesResNet   :704-747
  MultiImageFeatureResNet     :749-772

Same class names, constructor signatures, sub-module / state_dict key names and initialisation order (`torch.manual_seed(s)`
gives the reference's random-init weights; reference `.pth` files load), but forward / backward run in the CUDA library
(include/mivit.h: mivit_resnet_forward / backward / train_step, csrc/resnet.cu).  There is no PyTorch fallback.
Not supported (raise at construction): num_blocks other than [1, 1, 1] and num_classes other than 1 (the only values the
reference instantiates), activations other than nn.ReLU.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib

__all__ = ["BasicBlock", "LightResNet", "MultiImageResNet", "LightImagesFeaturesResNet", "MultiImageFeatureResNet", "CnnTrainer"]


def _no_direct_forward(self, *a, **k):
    raise NotImplementedError(
        "%s is a parameter container in moleculardiffusion_mivit_b200: its arithmetic runs inside MultiImageResNet / "
        "MultiImageFeatureResNet.forward (CUDA library)." % type(self).__name__)


def _check_activation(activation):
    if activation is not nn.ReLU:
        raise NotImplementedError("CNN baselines: only activation=nn.ReLU runs on the CUDA path")


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, in_channels, out_channels, stride=1, activation=nn.ReLU):
        super().__init__()
        _check_activation(activation)
        self.activation = activation(inplace=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.act1 = activation(inplace=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.act2 = activation(inplace=True)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_channels != out_channels:
            self.shortcut = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=stride, bias=False),
                                          nn.BatchNorm2d(out_channels))

    forward = _no_direct_forward


class _Trunk(nn.Module):
    """conv1 .. fc_act of LightResNet / LightImagesFeaturesResNet (identical in both, helpers/models.py:643-659 and :709-725)."""

    def _build_trunk(self, block, num_blocks, feature_size, activation):
        _check_activation(activation)
        if block is not BasicBlock or list(num_blocks) != [1, 1, 1]:
            raise NotImplementedError("CNN baselines: block=BasicBlock, num_blocks=[1, 1, 1] (what the reference instantiates)")
        self.in_channels = 32
        self.activation = activation
        self.conv1 = nn.Conv2d(1, 32, kernel_size=5, stride=2, padding=2, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        self.act = activation(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 32, num_blocks[0], stride=1)
        self.layer2 = self._make_layer(block, 64, num_blocks[1], stride=2)
        self.layer3 = self._make_layer(block, 128, num_blocks[2], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.feature_size = feature_size
        self.fc1 = nn.Linear(128 * block.expansion, self.feature_size)
        self.fc_act = activation(inplace=True)

    def _make_layer(self, block, out_channels, num_blocks, stride):
        strides = [stride] + [1] * (num_blocks - 1)
        layers = []
        for stride in strides:
            layers.append(block(self.in_channels, out_channels, stride, activation=self.activation))
            self.in_channels = out_channels * block.expansion
        return nn.Sequential(*layers)

    forward = _no_direct_forward


class LightResNet(_Trunk):
    def __init__(self, block, num_blocks, num_classes=1, feature_size=64, activation=nn.ReLU):
        super().__init__()
        if num_classes != 1:
            raise NotImplementedError("CNN baselines: num_classes=1 (the regression head of the reference experiments)")
        self._build_trunk(block, num_blocks, feature_size, activation)
        self.fc2 = nn.Linear(self.feature_size, num_classes)


class LightImagesFeaturesResNet(_Trunk):
    def __init__(self, block, num_blocks, feature_size=64, activation=nn.ReLU):
        super().__init__()
        self._build_trunk(block, num_blocks, feature_size, activation)


def _trunk_keys():
    k = ["conv1.weight", "bn1.weight", "bn1.bias"]
    for layer, sc in (("layer1.0", False), ("layer2.0", True), ("layer3.0", True)):
        k += ["%s.%s" % (layer, s) for s in ("conv1.weight", "bn1.weight", "bn1.bias", "conv2.weight", "bn2.weight", "bn2.bias")]
        if sc:
            k += ["%s.%s" % (layer, s) for s in ("shortcut.0.weight", "shortcut.1.weight", "shortcut.1.bias")]
    return ["resnet." + s for s in k + ["fc1.weight", "fc1.bias"]]


class _ResnetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, ext, *params):
        if not model.training:
            raise NotImplementedError("backward through a CNN baseline in eval() mode is not implemented on the CUDA path "
                                      "(BatchNorm backward uses batch statistics): call model.train() or torch.no_grad()")
        pred, gen = model._run_forward(x, ext, True)
        ctx.model, ctx.gen, ctx.has_ext = model, gen, ext is not None
        ctx.save_for_backward(*([x] + ([ext] if ext is not None else [])))
        return pred

    @staticmethod
    def backward(ctx, dpred):
        saved = list(ctx.saved_tensors)
        x = saved[0]
        ext = saved[1] if ctx.has_ext else None
        return (None, None, None) + tuple(ctx.model._run_backward(x, ext, dpred.contiguous(), ctx.gen))


class _CudaResNet(nn.Module):
    """Plumbing shared by MultiImageResNet and MultiImageFeatureResNet: one flat fp32 CUDA buffer behind all Parameters (canonical
    order = param_keys()), the flat BatchNorm running-statistics buffer, the workspace and the calls into the C library."""

    def _init_cuda_state(self):
        self._flat = self._grad_flat = self._bn_flat = self._bn_nbt = None
        self._ws, self._gen = {}, 0

    def _bn_modules(self):
        r = self.resnet
        return [r.bn1, r.layer1[0].bn1, r.layer1[0].bn2, r.layer2[0].bn1, r.layer2[0].bn2, r.layer2[0].shortcut[1],
                r.layer3[0].bn1, r.layer3[0].bn2, r.layer3[0].shortcut[1]]

    def resnet_config(self, n_frames, P):
        c = _lib.ResnetConfig()
        c.P, c.F, c.feature_size = int(P), int(n_frames), int(self.resnet.feature_size)
        c.ext_dim, c.hidden = int(self._ext_dim), int(self._hidden)
        c.single_prediction, c.activation = int(self._single_prediction), 0
        c.bn_eps, c.bn_momentum = 1e-5, 0.1
        return c

    def _ensure_flat(self):
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
            cfg = self.resnet_config(1, 9)
            L = _lib.lib()
            n = L.mivit_resnet_param_count(ctypes.byref(cfg))
            if n < 0:
                raise _lib.MivitError(L.mivit_last_error().decode())
            sizes = (ctypes.c_int64 * n)()
            _lib.check(L.mivit_resnet_param_sizes(ctypes.byref(cfg), sizes, n))
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
            self._ws.clear()
        bns = self._bn_modules()
        okb = self._bn_flat is not None and self._bn_flat.device == dev
        if okb:
            off = 0
            for bn in bns:
                if bn.running_mean.data_ptr() != self._bn_flat.data_ptr() + 4 * off:
                    okb = False
                    break
                off += 2 * bn.num_features
        if not okb:
            flat = torch.empty(2 * sum(bn.num_features for bn in bns), dtype=torch.float32, device=dev)
            nbt = torch.empty(len(bns), dtype=torch.int64, device=dev)
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
        key = (B, cfg.F, cfg.P)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = _lib.lib().mivit_resnet_workspace_bytes(ctypes.byref(cfg), B)
            if nbytes < 0:
                raise _lib.MivitError(_lib.lib().mivit_last_error().decode())
            self._ws.clear()
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self._flat.device)
            self._ws[key] = ws
        return ws

    def _check_inputs(self, x, ext):
        if x.dim() != 4 or x.shape[2] != x.shape[3]:
            raise ValueError("expected images of shape [batch_size, num_images, size, size]")
        dev = self._flat.device
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        if self._ext_dim:
            if ext is None or ext.dim() != 2 or ext.shape[0] != x.shape[0] or ext.shape[1] != self._ext_dim:
                raise ValueError("external_features must be [batch_size, %d]" % self._ext_dim)
            ext = ext.to(device=dev, dtype=torch.float32).contiguous()
        else:
            ext = None
        return x, ext

    def _pred_shape(self, cfg, B):
        return (B, cfg.F, 1) if (not cfg.ext_dim and not cfg.single_prediction) else (B, 1)

    def _run_forward(self, x, ext, training):
        B, Fr, P = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        cfg = self.resnet_config(Fr, P)
        ws = self._workspace(cfg, B)
        pred = torch.empty(self._pred_shape(cfg, B), dtype=torch.float32, device=self._flat.device)
        _lib.check(_lib.lib().mivit_resnet_forward(ctypes.byref(cfg), B, _lib.ptr(x), _lib.ptr(ext), _lib.ptr(self._flat),
                                                   _lib.ptr(self._bn_flat), _lib.ptr(self._bn_nbt), _lib.ptr(ws), _lib.ptr(pred),
                                                   int(bool(training)), _lib.current_stream()))
        self._gen += 1
        return pred, self._gen

    def _run_backward(self, x, ext, dpred, gen):
        if gen != self._gen:
            raise RuntimeError("backward() must follow the forward() that produced the output (one live workspace per model)")
        B, Fr, P = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        cfg = self.resnet_config(Fr, P)
        ws = self._workspace(cfg, B)
        _lib.check(_lib.lib().mivit_resnet_backward(ctypes.byref(cfg), B, _lib.ptr(x), _lib.ptr(ext), _lib.ptr(dpred),
                                                    _lib.ptr(self._flat), _lib.ptr(self._grad_flat), _lib.ptr(ws),
                                                    _lib.current_stream()))
        named = dict(self.named_parameters())
        grads, off = [], 0
        for k in self.param_keys():
            p = named[k]
            grads.append(self._grad_flat[off:off + p.numel()].view(p.shape).clone())
            off += p.numel()
        return grads

    def _forward_cuda(self, x, ext):
        self._ensure_flat()
        x, ext = self._check_inputs(x, ext)
        named = dict(self.named_parameters())
        params = [named[k] for k in self.param_keys()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _ResnetFunction.apply(self, x, ext, *params)
        pred, _ = self._run_forward(x, ext, self.training)
        return pred


class MultiImageResNet(_CudaResNet):
    def __init__(self, image_size, num_classes=1, single_prediction=True, activation=nn.ReLU):
        super().__init__()
        self.single_prediction = single_prediction
        self.resnet = LightResNet(BasicBlock, [1, 1, 1], num_classes, activation=activation)
        self._ext_dim, self._hidden, self._single_prediction = 0, 0, bool(single_prediction)
        self._init_cuda_state()

    def param_keys(self):
        return _trunk_keys() + ["resnet.fc2.weight", "resnet.fc2.bias"]

    def forward(self, x):
        """x: [batch_size, num_images, height, width] -> [batch_size, 1] (mean of the per-frame predictions) or
        [batch_size, num_images, 1] with single_prediction=False."""
        return self._forward_cuda(x, None)


class MultiImageFeatureResNet(_CudaResNet):
    def __init__(self, image_size, external_dim, feature_size=64, hidden_size=128, activation=nn.ReLU):
        super().__init__()
        self.resnet = LightImagesFeaturesResNet(BasicBlock, [1, 1, 1], feature_size, activation=activation)
        self.feature_size = feature_size
        self.external_dim = external_dim
        self.mlp = nn.Sequential(nn.Linear(feature_size + external_dim, hidden_size), activation(inplace=True),
                                 nn.Linear(hidden_size, 1))
        self._ext_dim, self._hidden, self._single_prediction = int(external_dim), int(hidden_size), True
        self._init_cuda_state()

    def param_keys(self):
        return _trunk_keys() + ["mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias"]

    def forward(self, x, external_features):
        """x: [batch_size, num_images, height, width], external_features: [batch_size, external_dim] -> [batch_size, 1]."""
        return self._forward_cuda(x, external_features)


class CnnTrainer:
    """MiViTTrainer's interface for the CNN baselines: the reference loop body (zero_grad, model(x), MSELoss, backward,
    AdamW.step; StepLR(5, 0.9) once per cycle, Experiments/PSFNoise/trainModelsPSFNoise.py:182-196) as ONE C-ABI call."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, step_size=5, gamma=0.9):
        if not isinstance(model, _CudaResNet):
            raise TypeError("CnnTrainer drives a moleculardiffusion_mivit_b200.baselines.MultiImageResNet / MultiImageFeatureResNet")
        self.model = model
        self.base_lr, self.lr = float(lr), float(lr)
        self.betas, self.eps, self.weight_decay = betas, float(eps), float(weight_decay)
        self.step_size, self.gamma, self.epoch, self.step_count = int(step_size), float(gamma), 0, 0
        model._ensure_flat()
        n, dev = model._n_params, model._flat.device
        self.m = torch.zeros(n + 4, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n + 4, dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self._scratch = {}

    def scheduler_step(self):
        self.epoch += 1
        self.lr = self.base_lr * self.gamma ** (self.epoch // self.step_size)

    def train_step(self, x, target, features=None):
        """x: [B,F,P,P]; target: [B,1] ([B,F,1] for single_prediction=False); features: [B,external_dim] for
        MultiImageFeatureResNet.  Returns the trainer's (device) loss buffer without synchronising."""
        model = self.model
        model._ensure_flat()
        if not model.training:
            model.train()
        x, ext = model._check_inputs(x, features)
        B, Fr, P = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        cfg = model.resnet_config(Fr, P)
        ws = model._workspace(cfg, B)
        rows = int(_lib.lib().mivit_resnet_pred_rows(ctypes.byref(cfg), B))
        target = target.to(device=model._flat.device, dtype=torch.float32).reshape(-1, 1).contiguous()
        if target.shape[0] != rows:
            raise ValueError("target has %d rows, the model predicts %d" % (target.shape[0], rows))
        buf = self._scratch.get(rows)
        if buf is None:
            dev = model._flat.device
            buf = self._scratch[rows] = (torch.empty((rows, 1), dtype=torch.float32, device=dev),
                                         torch.empty((rows, 1), dtype=torch.float32, device=dev))
        pred, dpred = buf
        self.step_count += 1
        _lib.check(_lib.lib().mivit_resnet_train_step(
            ctypes.byref(cfg), B, _lib.ptr(x), _lib.ptr(ext), _lib.ptr(target), _lib.ptr(model._flat), _lib.ptr(model._grad_flat),
            _lib.ptr(self.m), _lib.ptr(self.v), _lib.ptr(model._bn_flat), _lib.ptr(model._bn_nbt), _lib.ptr(ws), _lib.ptr(pred),
            _lib.ptr(self.loss), _lib.ptr(dpred), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
            self.step_count, 1, _lib.current_stream()))
        model._gen += 1
        self.last_pred = pred
        return self.loss
