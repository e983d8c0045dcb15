"""Oracle: Reduced-ResNet18 / SupConResNet forward, backward, SGD and MIR scores.

A torch-CPU fp32 (or fp64) functional restatement of the only model on the
replay-step path: reference models/resnet.py:14-37 (BasicBlock), :69-109
(ResNet.features/logits/forward), :112-116 (Reduced_ResNet18, nf=20),
:140-168 (SupConResNet) and the per-dataset linear replacement of
utils/setup_elements.py:46-68.  Parameters live in a flat ``dict name->tensor``
whose keys equal the reference ``state_dict()`` keys, so reference weights load
unchanged.  Test infrastructure only -- see oracle/__init__.py.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm2d default (models/resnet.py:20)
BN_MOMENTUM = 0.1


class Spec:
    """Static description of one network instance."""

    def __init__(self, in_hw=32, nf=20, num_classes=100, head=None, feat_dim=128):
        self.in_hw = in_hw
        self.nf = nf
        self.num_classes = num_classes
        self.head = head                      # None | 'linear' | 'mlp' | 'None'  (SupConResNet)
        self.feat_dim = feat_dim
        hw = in_hw
        for _ in range(3):                    # three stride-2 stages, 3x3 pad 1
            hw = (hw + 2 - 3) // 2 + 1
        self.final_hw = hw
        self.pooled_hw = hw // 4              # avg_pool2d(out, 4) floors (resnet.py:97)
        self.dim_in = nf * 8 * self.pooled_hw * self.pooled_hw
        self.prefix = 'encoder.' if head is not None else ''

    @property
    def is_supcon(self):
        return self.head is not None


def block_plan(spec):
    """[(name, cin, cout, stride, has_shortcut)] for the 8 BasicBlocks (resnet.py:78-88)."""
    plan, cin = [], spec.nf
    for li, mult in enumerate((1, 2, 4, 8), start=1):
        cout = spec.nf * mult
        for bi in range(2):
            stride = (1 if li == 1 else 2) if bi == 0 else 1
            plan.append(('layer%d.%d' % (li, bi), cin, cout, stride, stride != 1 or cin != cout))
            cin = cout
    return plan


def param_shapes(spec):
    """Ordered name->shape of every learnable tensor, in ``model.parameters()`` order."""
    p, out = spec.prefix, OrderedDict()
    out[p + 'conv1.weight'] = (spec.nf, 3, 3, 3)
    out[p + 'bn1.weight'] = (spec.nf,)
    out[p + 'bn1.bias'] = (spec.nf,)
    for name, cin, cout, stride, sc in block_plan(spec):
        out[p + name + '.conv1.weight'] = (cout, cin, 3, 3)
        out[p + name + '.bn1.weight'] = (cout,)
        out[p + name + '.bn1.bias'] = (cout,)
        out[p + name + '.conv2.weight'] = (cout, cout, 3, 3)
        out[p + name + '.bn2.weight'] = (cout,)
        out[p + name + '.bn2.bias'] = (cout,)
        if sc:
            out[p + name + '.shortcut.0.weight'] = (cout, cin, 1, 1)
            out[p + name + '.shortcut.1.weight'] = (cout,)
            out[p + name + '.shortcut.1.bias'] = (cout,)
    if spec.is_supcon:
        # SupConResNet keeps the encoder's own (unused) classifier: Reduced_ResNet18(100), resnet.py:144
        out[p + 'linear.weight'] = (100, spec.nf * 8)
        out[p + 'linear.bias'] = (100,)
        if spec.head == 'linear':
            out['head.weight'] = (spec.feat_dim, spec.dim_in)
            out['head.bias'] = (spec.feat_dim,)
        elif spec.head == 'mlp':
            out['head.0.weight'] = (spec.dim_in, spec.dim_in)
            out['head.0.bias'] = (spec.dim_in,)
            out['head.2.weight'] = (spec.feat_dim, spec.dim_in)
            out['head.2.bias'] = (spec.feat_dim,)
    else:
        out['linear.weight'] = (spec.num_classes, spec.dim_in)
        out['linear.bias'] = (spec.num_classes,)
    return out


def bn_names(spec):
    p, out = spec.prefix, [spec.prefix + 'bn1']
    for name, cin, cout, stride, sc in block_plan(spec):
        out += [p + name + '.bn1', p + name + '.bn2']
        if sc:
            out.append(p + name + '.shortcut.1')
    return out


def seeded_state(spec, seed, dtype=torch.float32):
    """Deterministic, non-trivial parameters AND running statistics from a numpy
    stream (so fixtures need not store 1.1 M weights).  Returns (params, bn_state)."""
    rs = np.random.RandomState(seed)
    params = OrderedDict()
    for name, shape in param_shapes(spec).items():
        if len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            a = rs.standard_normal(shape) * math.sqrt(2.0 / fan_in)
        elif len(shape) == 2:
            a = rs.standard_normal(shape) / math.sqrt(shape[1])
        elif name.endswith('.weight'):
            a = rs.uniform(0.5, 1.5, shape)
        else:
            a = rs.standard_normal(shape) * 0.1
        params[name] = torch.tensor(a.astype(np.float32)).to(dtype)
    bn = OrderedDict()
    for name in bn_names(spec):
        c = params[name + '.weight'].shape[0]
        bn[name + '.running_mean'] = torch.tensor((rs.standard_normal(c) * 0.1).astype(np.float32)).to(dtype)
        bn[name + '.running_var'] = torch.tensor(rs.uniform(0.5, 1.5, c).astype(np.float32)).to(dtype)
        bn[name + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)
    return params, bn


def _bn(x, params, bn, name, train):
    if train and bn is not None:
        bn[name + '.num_batches_tracked'] += 1
    return F.batch_norm(x, bn[name + '.running_mean'], bn[name + '.running_var'],
                        params[name + '.weight'], params[name + '.bias'],
                        training=train, momentum=BN_MOMENTUM, eps=BN_EPS)


def features(spec, params, bn, x, train):
    """Encoder features before the classifier (resnet.py:90-99).  In train mode the
    running statistics in ``bn`` are updated in place (momentum 0.1, unbiased var)."""
    p = spec.prefix
    out = F.relu(_bn(F.conv2d(x, params[p + 'conv1.weight'], padding=1), params, bn, p + 'bn1', train))
    for name, cin, cout, stride, sc in block_plan(spec):
        n = p + name
        h = F.relu(_bn(F.conv2d(out, params[n + '.conv1.weight'], stride=stride, padding=1),
                       params, bn, n + '.bn1', train))
        h = _bn(F.conv2d(h, params[n + '.conv2.weight'], padding=1), params, bn, n + '.bn2', train)
        if sc:
            s = _bn(F.conv2d(out, params[n + '.shortcut.0.weight'], stride=stride),
                    params, bn, n + '.shortcut.1', train)
        else:
            s = out
        out = F.relu(h + s)
    out = F.avg_pool2d(out, 4)
    return out.reshape(out.size(0), -1)


def forward(spec, params, bn, x, train):
    """Logits (Reduced_ResNet18, resnet.py:106-109) or the L2-normalised projection
    (SupConResNet.forward, resnet.py:159-165)."""
    feat = features(spec, params, bn, x, train)
    if not spec.is_supcon:
        return F.linear(feat, params['linear.weight'], params['linear.bias'])
    if spec.head == 'mlp':
        feat = F.linear(F.relu(F.linear(feat, params['head.0.weight'], params['head.0.bias'])),
                        params['head.2.weight'], params['head.2.bias'])
    elif spec.head == 'linear':
        feat = F.linear(feat, params['head.weight'], params['head.bias'])
    return F.normalize(feat, dim=1)


def ce_loss_and_grads(spec, params, bn, x, y):
    """One train-mode forward + mean cross-entropy + backward (agents/base.py:113,
    exp_replay.py:40-55).  Returns (loss, logits, grads dict)."""
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
    logits = forward(spec, leaves, bn, x, train=True)
    loss = F.cross_entropy(logits, y)
    loss.backward()
    grads = OrderedDict((k, (v.grad if v.grad is not None else None)) for k, v in leaves.items())
    return loss.detach(), logits.detach(), grads


def sgd_step(params, grads, lr, weight_decay=0.0):
    """torch.optim.SGD without momentum (setup_elements.py:73-75): p -= lr*(g + wd*p);
    tensors without a gradient are skipped."""
    for k, g in grads.items():
        if g is None:
            continue
        if weight_decay != 0.0:
            g = g.add(params[k], alpha=weight_decay)
        params[k].add_(g, alpha=-lr)      # one rounding, like torch.optim.SGD


def mir_scores(spec, params, bn, grads, lr, sub_x, sub_y):
    """MIR interference scores (mir_retrieve.py:15-30): per-sample CE after a
    virtual SGD step minus before, both forwards in TRAIN mode; the real model's
    running statistics move once (logits_pre), the virtual model's copy is dropped."""
    virt = OrderedDict((k, (v - lr * grads[k]) if grads.get(k) is not None else v.clone())
                       for k, v in params.items())
    bn_virt = OrderedDict((k, v.clone()) for k, v in bn.items())     # deepcopy, mir_retrieve.py:41
    with torch.no_grad():
        pre = F.cross_entropy(forward(spec, params, bn, sub_x, True), sub_y, reduction='none')
        post = F.cross_entropy(forward(spec, virt, bn_virt, sub_x, True), sub_y, reduction='none')
    return post - pre
