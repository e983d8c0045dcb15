"""Does any kernel read workspace memory it did not write in the same pass?  Train forward + backward on a
zero-filled and on a NaN-filled workspace must give bit-identical gradients.  GPU box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200ocl import nets  # noqa: E402


def main():
    torch.manual_seed(0)
    g = torch.Generator(device='cuda').manual_seed(1)
    for head in (None, 'mlp'):
        model = nets.Reduced_ResNet18(100) if head is None else nets.SupConResNet(head='mlp')
        eng = nets.engine_of(model)
        for n in (7, 10, 20, 110):
            x = torch.rand(n, 3, 32, 32, device='cuda', generator=g)
            for tcp in ('1', '0'):
                os.environ['B200OCL_TCP'] = tcp
                out = []
                for fill in (0.0, float('nan'), 1e30):
                    ws = eng.new_train_workspace(n)
                    ws.view(torch.float32).fill_(fill)
                    bn0 = eng.state.bn_stats.clone()
                    o, ws = eng.forward_train(x, ws=ws)
                    eng.state.bn_stats.copy_(bn0)
                    dout = torch.randn(o.shape, device='cuda', generator=torch.Generator(device='cuda').manual_seed(5)) / n
                    eng.backward(x, dout, ws)
                    out.append((o.clone(), eng.state.grads.clone()))
                torch.cuda.synchronize()
                same_o = [bool(torch.equal(out[0][0], t[0])) for t in out[1:]]
                same_g = [bool(torch.equal(out[0][1], t[1])) for t in out[1:]]
                nan_g = [int(torch.isnan(t[1]).sum()) for t in out]
                print('head=%s N=%d TCP=%s  outputs identical %s  grads identical %s  NaNs in grads %s' %
                      (head, n, tcp, same_o, same_g, nan_g), flush=True)
                if not all(same_g):
                    gd = (out[0][1] - out[2][1]).abs()
                    bad = torch.nonzero(gd > 0).flatten()
                    offs = [o_ for (o_, c_, h_) in eng.table]
                    import bisect
                    tens = sorted(set(bisect.bisect_right(offs, int(b)) - 1 for b in bad[:100000:997].tolist()))
                    print('    differing tensors (sample):', tens[:30], flush=True)


if __name__ == '__main__':
    main()
