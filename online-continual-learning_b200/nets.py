"""Model objects of the replay path: Reduced_ResNet18 / SupConResNet with the reference's
constructor signatures (models/resnet.py:112-116,140-168; utils/setup_elements.py:46-68), backed
by the CUDA engine, plus `adopt()` which moves an existing reference nn.Module onto the engine
by aliasing its Parameters and BatchNorm buffers onto the engine's flat arenas (so the
reference's own evaluate(), state_dict() and optimizer objects keep seeing live weights).
"""
import math
import weakref

import torch
import torch.nn as nn

from .engine import Engine
from .memory import input_size_match, n_classes

_ENGINES = weakref.WeakKeyDictionary()


def engine_of(model):
    """The Engine behind a model object (EngineModel or adopted reference module)."""
    if isinstance(model, EngineModel):
        return model.engine
    eng = _ENGINES.get(model)
    if eng is None:
        raise RuntimeError('model is not backed by the b200ocl engine; call b200ocl.nets.adopt(model, in_hw) first')
    return eng


def param_layout(dim_in, num_classes, head=None, feat_dim=128, nf=20):
    """[(state_dict name, shape)] in parameters() order for the networks of
    utils/setup_elements.py:46-68 (dim_in = flattened encoder feature size)."""
    pre = 'encoder.' if head is not None else ''
    out = [(pre + 'conv1.weight', (nf, 3, 3, 3)), (pre + 'bn1.weight', (nf,)), (pre + 'bn1.bias', (nf,))]
    cin = nf
    for li in range(1, 5):
        cout = nf << (li - 1)
        for bi in range(2):
            stride = 2 if (bi == 0 and li > 1) else 1
            b = '%slayer%d.%d.' % (pre, li, bi)
            out += [(b + 'conv1.weight', (cout, cin, 3, 3)), (b + 'bn1.weight', (cout,)), (b + 'bn1.bias', (cout,)),
                    (b + 'conv2.weight', (cout, cout, 3, 3)), (b + 'bn2.weight', (cout,)), (b + 'bn2.bias', (cout,))]
            if stride != 1 or cin != cout:
                out += [(b + 'shortcut.0.weight', (cout, cin, 1, 1)), (b + 'shortcut.1.weight', (cout,)),
                        (b + 'shortcut.1.bias', (cout,))]
            cin = cout
    if head is None:
        out += [('linear.weight', (num_classes, dim_in)), ('linear.bias', (num_classes,))]
    else:
        out += [(pre + 'linear.weight', (100, nf * 8)), (pre + 'linear.bias', (100,))]   # unused encoder classifier
        if head == 'linear':
            out += [('head.weight', (feat_dim, dim_in)), ('head.bias', (feat_dim,))]
        elif head == 'mlp':
            out += [('head.0.weight', (dim_in, dim_in)), ('head.0.bias', (dim_in,)),
                    ('head.2.weight', (feat_dim, dim_in)), ('head.2.bias', (feat_dim,))]
    return out


class EngineModel(nn.Module):
    """nn.Module facade over an Engine: Parameters are views into the parameter arena (their
    .grad views into the gradient arena), features()/forward() run the CUDA kernels.
    forward() follows self.training like the reference modules; it does not build an autograd
    graph -- the learners drive Engine.forward_train / backward / sgd_step directly."""

    def __init__(self, in_hw, num_classes, head=None, feat_dim=128, device='cuda'):
        super().__init__()
        self.engine = Engine(in_hw, num_classes, head=head, feat_dim=feat_dim, device=device)
        self.head_kind = head
        layout = param_layout(self.engine.dim_in, num_classes, head, feat_dim)
        names, shapes = [n for n, _ in layout], [sh for _, sh in layout]
        for (n, sh), (_, numel, _) in zip(layout, self.engine.table):
            assert int(torch.Size(sh).numel()) == numel, 'host/C layout mismatch at ' + n
        self._names = names
        for name, shape, pv, gv in zip(names, shapes, self.engine.param_views(), self.engine.grad_views()):
            p = nn.Parameter(pv.view(shape), requires_grad=True)
            p.grad = gv.view(shape)
            self.register_parameter(name.replace('.', '__'), p)
        self.reset_parameters()

    def reset_parameters(self):
        """Default torch initialisation of the reference layers (kaiming-uniform convs/linears,
        BN weight 1 / bias 0, running mean 0 / var 1)."""
        with torch.no_grad():
            for name, p in zip(self._names, self.parameters()):
                if p.dim() >= 2:
                    fan_in = p[0].numel()
                    bound = 1.0 / math.sqrt(fan_in)          # kaiming_uniform_(a=sqrt(5))
                    p.uniform_(-bound, bound)
                elif name.endswith('.bias') and ('linear' in name or name.startswith('head')):
                    w = self._param(name[:-4] + 'weight')
                    bound = 1.0 / math.sqrt(w.shape[1])
                    p.uniform_(-bound, bound)
                elif name.endswith('.weight'):
                    p.fill_(1.0)
                else:
                    p.zero_()
            for rm, rv in self.engine.bn_views():
                rm.zero_()
                rv.fill_(1.0)
            self.engine.state.bn_tracked.zero_()
        self.engine.pack()

    def _param(self, name):
        return getattr(self, name.replace('.', '__'))

    def named_state(self):
        return dict(zip(self._names, self.parameters()))

    def features(self, x):
        """Encoder features.  Eval mode: running statistics (what ASER and NCM use,
        utils/utils.py:55-58).  Train mode: batch statistics, running stats updated."""
        if self.training:
            raise NotImplementedError('train-mode features() is not on the replay path; use forward()')
        return self.engine.features_eval(x)

    def forward(self, x):
        if self.training:
            out, _ = self.engine.forward_train(x)
            return out
        raise NotImplementedError('eval-mode forward() belongs to evaluate() (SURVEY section 8f); use features()')


def Reduced_ResNet18(nclasses, nf=20, bias=True, in_hw=32):
    """models/resnet.py:112-116 (nf is fixed at 20 there and here)."""
    if nf != 20 or not bias:
        raise NotImplementedError('the engine implements the reference configuration nf=20, bias=True')
    return EngineModel(in_hw, nclasses, head=None)


def SupConResNet(dim_in=160, head='mlp', feat_dim=128, in_hw=None):
    """models/resnet.py:140-157.  dim_in selects the dataset like the reference does
    (160: 32x32 inputs, 640: 84x84 inputs; setup_elements.py:49-51)."""
    if in_hw is None:
        in_hw = {160: 32, 640: 84}[dim_in]
    return EngineModel(in_hw, 100, head=head, feat_dim=feat_dim)


def setup_architecture(params):
    """utils/setup_elements.py:46-68 for the datasets on the replay path."""
    nclass = n_classes[params.data]
    in_hw = input_size_match[params.data][1]
    if params.agent in ['SCR', 'SCP']:
        return SupConResNet(640 if params.data == 'mini_imagenet' else 160, head=params.head)
    if params.data in ('cifar100', 'cifar10', 'mini_imagenet'):
        return Reduced_ResNet18(nclass, in_hw=in_hw)
    raise NotImplementedError('dataset %s is outside the replay-path scope (SURVEY section 8)' % params.data)


def adopt(module, in_hw):
    """Move a reference nn.Module (models.resnet.ResNet with BasicBlocks, or SupConResNet) onto
    the engine.  Its Parameters / BN buffers are re-pointed at the engine arenas, keeping the
    objects (and therefore an optimizer built on module.parameters(), run.py:40) valid."""
    if isinstance(module, EngineModel):
        return module.engine
    if module in _ENGINES:
        return _ENGINES[module]
    is_supcon = hasattr(module, 'encoder')
    enc = module.encoder if is_supcon else module
    head = None
    if is_supcon:
        h = getattr(module, 'head', None)
        if h is None:
            head = 'None'
        elif isinstance(h, nn.Linear):
            head = 'linear'
        else:
            head = 'mlp'
    num_classes = enc.linear.out_features
    feat_dim = 128
    if head == 'linear':
        feat_dim = module.head.out_features
    elif head == 'mlp':
        feat_dim = module.head[2].out_features
    params = list(module.parameters())
    dev = params[0].device
    if dev.type != 'cuda':
        raise RuntimeError('adopt() needs the model on a CUDA device (the reference moves it there, run.py:39)')
    eng = Engine(in_hw, num_classes, head=head, feat_dim=feat_dim, device=dev)
    bns = [m for m in module.modules() if isinstance(m, nn.BatchNorm2d)]
    eng.load(params, [(m.running_mean, m.running_var, int(m.num_batches_tracked)) for m in bns])
    with torch.no_grad():
        for p, pv, gv in zip(params, eng.param_views(), eng.grad_views()):
            p.data = pv.view(p.shape)
            p.grad = gv.view(p.shape)
        for i, (m, (rm, rv)) in enumerate(zip(bns, eng.bn_views())):
            m.running_mean.data = rm
            m.running_var.data = rv
            m.num_batches_tracked.data = eng.state.bn_tracked[i]
    _ENGINES[module] = eng
    return eng
