"""SCR two-view gradient of tests/test_gpu_net.py::test_backward_supcon_two_views with the halo-patch
tensor-core convolutions switched on/off separately for the forward and the backward pass."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import resnet as oresnet, supcon as osup  # noqa: E402
from b200ocl import engine, ops  # noqa: E402


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


def main():
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'resnet.npz'))
    spec = oresnet.Spec(32, 20, 100, head='mlp')
    x1, x2, y = torch.tensor(g['scr_x1']).cuda(), torch.tensor(g['scr_x2']).cuda(), torch.tensor(g['scr_y']).cuda()
    p64, bn64 = oresnet.seeded_state(spec, 13, dtype=torch.float64)
    leaves = {k: v.clone().requires_grad_(True) for k, v in p64.items()}
    o1 = oresnet.forward(spec, leaves, {k: v.clone() for k, v in bn64.items()}, x1.cpu().double(), True)
    o2 = oresnet.forward(spec, leaves, {k: v.clone() for k, v in bn64.items()}, x2.cpu().double(), True)
    _, d64 = osup.supcon_loss_and_grad(torch.stack([o1, o2], dim=1).detach().numpy(), y.cpu().numpy(), 0.07)
    torch.autograd.backward([o1, o2], [torch.from_numpy(d64[:, 0].copy()), torch.from_numpy(d64[:, 1].copy())])
    names = list(leaves.keys())
    for fwd, bwd in (('1', '0'),):
        params, bn = oresnet.seeded_state(spec, 13)
        eng = engine.Engine(32, 100, head='mlp')
        eng.load(list(params.values()), [(bn[n + '.running_mean'], bn[n + '.running_var']) for n in oresnet.bn_names(spec)])
        os.environ['B200OCL_TCP'] = fwd
        f1, ws1 = eng.forward_train(x1, slot=0)
        f2, ws2 = eng.forward_train(x2, slot=1)
        feats = torch.stack([f1, f2], dim=1).contiguous()
        loss, dfeat = ops.supcon(feats, y, 0.07)
        os.environ['B200OCL_TCP'] = bwd
        eng.backward(x1, dfeat[:, 0].contiguous(), ws1)
        eng.backward(x2, dfeat[:, 1].contiguous(), ws2, accumulate=True)
        torch.cuda.synchronize()
        errs = []
        for n_, gv in zip(names, eng.grad_views()):
            ref = leaves[n_].grad
            if ref is None:
                continue
            errs.append((rel_err(gv.cpu().numpy().reshape(ref.shape), ref.numpy()), n_))
        ffeat = rel_err(feats.cpu().numpy(), torch.stack([o1, o2], dim=1).detach().numpy())
        for n_, gv in zip(names, eng.grad_views()):
            if n_ in ('encoder.layer3.1.bn2.bias', 'encoder.layer3.1.bn2.weight', 'encoder.layer3.0.bn2.bias'):
                ref = leaves[n_].grad.numpy()
                d = (gv.cpu().numpy().astype(np.float64) - ref) / np.abs(ref).max()
                print(n_, np.array2string(d, precision=1, max_line_width=250), flush=True)
            if n_ == 'encoder.layer3.1.conv2.weight':
                ref = leaves[n_].grad.numpy()
                d = np.abs(gv.cpu().numpy().astype(np.float64).reshape(ref.shape) - ref).max(axis=(1, 2, 3)) / np.abs(ref).max()
                print(n_, 'by out channel', np.array2string(d, precision=1, max_line_width=250), flush=True)
                d = np.abs(gv.cpu().numpy().astype(np.float64).reshape(ref.shape) - ref).max(axis=(0, 2, 3)) / np.abs(ref).max()
                print(n_, 'by in channel', np.array2string(d, precision=1, max_line_width=250), flush=True)
        print('   all:', ['%s %.1e' % (n_.replace('encoder.', ''), e) for e, n_ in errs], flush=True)
        errs.sort(reverse=True)
        print('fwd TCP=%s bwd TCP=%s  feature err %.2e  worst grads: %s' %
              (fwd, bwd, ffeat, ['%s %.1e' % (n_, e) for e, n_ in errs[:5]]), flush=True)
    os.environ['B200OCL_TCP'] = '1'


if __name__ == '__main__':
    main()
