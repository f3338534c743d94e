"""TEST INFRASTRUCTURE ONLY -- plain-PyTorch fp32 restatement of the reference's CNN baselines over a flat
{state_dict key: tensor} mapping (the checker of csrc/resnet.cu; never imported by the product package).

Parity status: PINNED against the unmodified reference classes (helpers/models.py:600-772, imported through oracle/refshim.py)
by tests/golden/resnet_*.npz (oracle/make_golden_resnet.py) in tests/test_oracle_resnet.py.

Reference code restated (paths under /root/reference/helpers/models.py):
  :600-635  BasicBlock                 -> _block
  :638-683  LightResNet                -> _trunk (+ fc2)
  :686-701  MultiImageResNet           -> forward (external_dim = 0)
  :704-747  LightImagesFeaturesResNet  -> _trunk
  :749-772  MultiImageFeatureResNet    -> forward (external_dim > 0)
"""
import torch
import torch.nn.functional as F

from .vit_oracle import _bn, is_buffer


def _block(x, sd, pre, stride, training, stats_out):
    out = F.conv2d(x, sd[pre + ".conv1.weight"], stride=stride, padding=1)
    out = F.relu(_bn(out, sd, pre + ".bn1", training, stats_out))
    out = F.conv2d(out, sd[pre + ".conv2.weight"], stride=1, padding=1)
    out = _bn(out, sd, pre + ".bn2", training, stats_out)
    if pre + ".shortcut.0.weight" in sd:
        idt = F.conv2d(x, sd[pre + ".shortcut.0.weight"], stride=stride)
        idt = _bn(idt, sd, pre + ".shortcut.1", training, stats_out)
    else:
        idt = x
    return F.relu(out + idt)


def _trunk(x, sd, training, stats_out, pre="resnet"):
    """[N,1,H,W] -> act(fc1(avgpool(layer3(layer2(layer1(maxpool(act(bn1(conv1 x)))))))))  [N, feature_size]"""
    out = F.conv2d(x, sd[pre + ".conv1.weight"], stride=2, padding=2)
    out = F.relu(_bn(out, sd, pre + ".bn1", training, stats_out))
    out = F.max_pool2d(out, kernel_size=3, stride=2, padding=1)
    out = _block(out, sd, pre + ".layer1.0", 1, training, stats_out)
    out = _block(out, sd, pre + ".layer2.0", 2, training, stats_out)
    out = _block(out, sd, pre + ".layer3.0", 2, training, stats_out)
    out = torch.flatten(F.adaptive_avg_pool2d(out, (1, 1)), 1)
    return F.relu(F.linear(out, sd[pre + ".fc1.weight"], sd[pre + ".fc1.bias"]))


def forward(sd, x, external_features=None, single_prediction=True, training=True, stats_out=None):
    """MultiImageResNet.forward (:692-701) when the state dict has resnet.fc2.*, MultiImageFeatureResNet.forward (:763-772) when
    it has mlp.*.  x: [B, F, H, W]."""
    B, Fr, h, w = x.shape
    f = _trunk(x.reshape(B * Fr, 1, h, w), sd, training, stats_out)
    if "mlp.0.weight" in sd:
        feats = f.view(B, Fr, -1).mean(dim=1)
        comb = torch.cat([feats, external_features], dim=1)
        hid = F.relu(F.linear(comb, sd["mlp.0.weight"], sd["mlp.0.bias"]))
        return F.linear(hid, sd["mlp.2.weight"], sd["mlp.2.bias"])
    y = F.linear(f, sd["resnet.fc2.weight"], sd["resnet.fc2.bias"]).view(B, Fr, 1)
    return y.mean(dim=1) if single_prediction else y


def loss_and_grads(sd, x, target, external_features=None, single_prediction=True):
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if not is_buffer(k)}
    full = dict(sd)
    full.update(params)
    stats = {}
    pred = forward(full, x, external_features, single_prediction, training=True, stats_out=stats)
    loss = F.mse_loss(pred, target)
    keys = list(params.keys())
    grads = torch.autograd.grad(loss, [params[k] for k in keys])
    return pred.detach(), loss.detach(), dict(zip(keys, grads)), stats
