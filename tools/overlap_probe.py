"""How much do two independent N=110 train passes (SCR's two views) gain from running on two streams?
Timing only: the shared running statistics / gradient arena race is ignored here."""
import contextlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    torch.cuda.set_device(0)
    with contextlib.redirect_stdout(sys.stderr):
        scr = bench.build_learner('scr', 20)
    eng = scr.engine
    N = int(os.environ.get('PROBE_N', '110'))
    x1 = torch.rand(N, 3, 32, 32, device='cuda')
    x2 = torch.rand(N, 3, 32, 32, device='cuda')
    main_s = torch.cuda.current_stream()
    side = torch.cuda.Stream()

    def seq():
        o1, w1 = eng.forward_train(x1, slot=0)
        o2, w2 = eng.forward_train(x2, slot=1)
        eng.backward(x1, o1, w1)
        eng.backward(x2, o2, w2, accumulate=True)

    def par():
        side.wait_stream(main_s)
        with torch.cuda.stream(side):
            o2, w2 = eng.forward_train(x2, slot=1)
        o1, w1 = eng.forward_train(x1, slot=0)
        main_s.wait_stream(side)
        side.wait_stream(main_s)
        with torch.cuda.stream(side):
            eng.backward(x2, o2, w2)
        eng.backward(x1, o1, w1)
        main_s.wait_stream(side)

    def fwd_only_seq():
        eng.forward_train(x1, slot=0); eng.forward_train(x2, slot=1)

    def fwd_only_par():
        side.wait_stream(main_s)
        with torch.cuda.stream(side):
            eng.forward_train(x2, slot=1)
        eng.forward_train(x1, slot=0)
        main_s.wait_stream(side)

    for name, fn in (('seq', seq), ('par', par), ('fwd_seq', fwd_only_seq), ('fwd_par', fwd_only_par), ('seq', seq), ('par', par)):
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print('%-8s N=%d  median %.3f ms  min %.3f' % (name, N, float(np.median(ts)), min(ts)))


if __name__ == '__main__':
    main()
