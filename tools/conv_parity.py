"""Same network pass through the halo-patch tensor-core convolutions (default) and through the other
kernels (B200OCL_TCP=0): element-wise comparison of features, train outputs, BN statistics and gradients.
GPU box.  usage: python tools/conv_parity.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200ocl.engine import Engine  # noqa: E402
from b200ocl import nets  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    torch.manual_seed(0)
    ref_model = nets.Reduced_ResNet18(100)
    eng = nets.engine_of(ref_model)
    g = torch.Generator(device='cuda').manual_seed(1)
    for n in (7, 10, 110):
        x = torch.rand(n, 3, 32, 32, device='cuda', generator=g)
        res = {}
        for tcp in ('1', '0'):
            os.environ['B200OCL_TCP'] = tcp
            bn0 = eng.state.bn_stats.clone()
            f = eng.features_eval(x).clone()
            o, ws = eng.forward_train(x, ws=eng.new_train_workspace(n))
            bn1 = eng.state.bn_stats.clone()
            eng.state.bn_stats.copy_(bn0)
            dout = torch.randn(o.shape, device='cuda', generator=torch.Generator(device='cuda').manual_seed(5)) / n
            eng.backward(x, dout, ws)
            res[tcp] = (f, o.clone(), bn1, eng.state.grads.clone())
            torch.cuda.synchronize()
        names = ('features_eval', 'train_out', 'bn_stats', 'grads')
        print('N=%d' % n, {k: '%.2e' % rel(a, b) for k, a, b in zip(names, res['1'], res['0'])}, flush=True)
        # backward only: same forward state (TCP), data gradients through both paths
        os.environ['B200OCL_TCP'] = '1'
        o, ws = eng.forward_train(x, ws=eng.new_train_workspace(n))
        gg = []
        for tcp in ('1', '0', '1'):
            os.environ['B200OCL_TCP'] = tcp
            eng.backward(x, dout, ws)
            gg.append(eng.state.grads.clone())
        os.environ['B200OCL_TCP'] = '1'
        print('   backward-only toggle: tcp vs other %.2e   tcp vs tcp again %.2e' % (rel(gg[0], gg[1]), rel(gg[0], gg[2])), flush=True)
        # per-tensor gradient differences
        ga, gb = res['1'][3], res['0'][3]
        worst = []
        for i, (off, cnt, has_grad) in enumerate(eng.table):
            if has_grad:
                worst.append((rel(ga[off:off + cnt], gb[off:off + cnt]), 'tensor%d[%d]' % (i, cnt)))
        worst.sort(reverse=True)
        print('   worst grads:', ['%s %.1e' % (nm, e) for e, nm in worst[:6]], flush=True)


if __name__ == '__main__':
    main()
