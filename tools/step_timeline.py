"""Device-time timeline of one ER+ASER and one SCR replay step at the bench workload, at the granularity of the
engine / plugin calls (CUDA events around every call on the launch stream, CUDA graphs on, median over the timed
steps).  Writes gpurun_out/step_timeline.json.  A diagnostic, not the bench."""
import contextlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

SPANS = []


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        SPANS.append((label, e0, e1))
        return out
    setattr(obj, name, inner)


def main(steps=12, warm=6):
    from b200ocl import ops
    torch.cuda.set_device(0)
    np.random.seed(1000); torch.manual_seed(1000)
    with contextlib.redirect_stdout(sys.stderr):
        aser = bench.build_learner('aser', 10)
        scr = bench.build_learner('scr', 20)
    for L, tag in ((aser, 'aser'), (scr, 'scr')):
        wrap(L.engine, 'forward_train', tag + '.forward_train')
        wrap(L.engine, 'backward', tag + '.backward')
        wrap(L.engine, 'features_eval', tag + '.features_eval')
        wrap(L.buffer, 'retrieve', tag + '.retrieve(total)')
        wrap(L.buffer, 'update', tag + '.update(total)')
        wrap(L, '_optimizer_step', tag + '.optimizer_step')
        wrap(L, 'replay_step', tag + '.STEP')
    wrap(ops, 'knn_sv', 'ops.knn_sv')
    if hasattr(ops, 'supcon'):
        wrap(ops, 'supcon', 'ops.supcon')
    rs = np.random.RandomState(7)
    total = steps + warm
    x = torch.from_numpy(rs.rand(2 * total, 10, 3, 32, 32).astype(np.float32)).cuda()
    yh = rs.randint(0, 100, (2 * total, 10)).astype(np.int64)
    y = torch.from_numpy(yh).cuda()
    per_step = []
    for i in range(total):
        del SPANS[:]
        aser.replay_step(x[2 * i], y[2 * i], yh[2 * i])
        scr.replay_step(x[2 * i + 1], y[2 * i + 1], yh[2 * i + 1])
        torch.cuda.synchronize()
        if i >= warm:
            per_step.append([(lab, e0.elapsed_time(e1)) for lab, e0, e1 in SPANS])
    labels = [lab for lab, _ in per_step[0]]
    assert all([lab for lab, _ in s] == labels for s in per_step), 'call sequence differs between steps'
    med = np.median(np.array([[ms for _, ms in s] for s in per_step]), axis=0)
    rows = [{'call': lab, 'ms': round(float(m), 4)} for lab, m in zip(labels, med)]
    for r in rows:
        print('%-28s %8.3f ms' % (r['call'], r['ms']))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'step_timeline.json'), 'w') as fh:
        json.dump(rows, fh, indent=1)


if __name__ == '__main__':
    main()
