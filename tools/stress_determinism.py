"""Repeat eval pass + train forward + backward many times on fresh workspaces; every repetition must give
bit-identical logits and gradients (sporadic races show up as mismatches).  GPU box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200ocl import nets  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    torch.manual_seed(0)
    g = torch.Generator(device='cuda').manual_seed(1)
    model = nets.Reduced_ResNet18(100)
    eng = nets.engine_of(model)
    bn0 = eng.state.bn_stats.clone()
    for n in (10, 20, 110, 210):
        x = torch.rand(n, 3, 32, 32, device='cuda', generator=g)
        dout = torch.randn(n, 100, device='cuda', generator=g) / n
        for tcp in ('1', '0'):
            os.environ['B200OCL_TCP'] = tcp
            first = None
            bad = {'feat': 0, 'out': 0, 'grads': 0}
            worst = 0.0
            for it in range(reps):
                eng.state.bn_stats.copy_(bn0)
                f = eng.features_eval(x)
                o, ws = eng.forward_train(x, ws=eng.new_train_workspace(n))
                eng.backward(x, dout, ws)
                cur = (f.clone(), o.clone(), eng.state.grads.clone())
                if first is None:
                    first = cur
                else:
                    for k, a, b in zip(('feat', 'out', 'grads'), cur, first):
                        if not torch.equal(a, b):
                            bad[k] += 1
                            worst = max(worst, float((a - b).abs().max() / b.abs().max()))
            torch.cuda.synchronize()
            print('N=%d TCP=%s reps=%d mismatching repetitions %s worst rel %.2e' % (n, tcp, reps, bad, worst), flush=True)


if __name__ == '__main__':
    main()
