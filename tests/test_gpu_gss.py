"""GPU parity of the GSS-greedy path (utils/buffer/gss_greedy_update.py): the gradient-cosine
kernel, the differentiable eval-mode pass of the engine, and the update plugin replayed on the run recorded from
the reference (tests/golden/gss.npz) -- same replaced slots, same labels, scores within 1e-3."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import gss as ogss
from oracle import resnet as oresnet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def b():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import b200ocl  # noqa: F401
    from b200ocl import engine, memory, nets, ops, registry, update
    return SimpleNamespace(engine=engine, memory=memory, nets=nets, ops=ops, registry=registry, update=update)


@pytest.mark.parametrize('K,n', [(1, 1000), (3, 4097), (10, 1093710), (64, 70001)])
def test_grad_cosine_kernel(b, K, n):
    """b200ocl_grad_cosine against the reference's cosine_similarity (buffer_utils.py:50-55) in fp64."""
    rs = np.random.RandomState(K)
    mem = rs.standard_normal((K, n)).astype(np.float32)
    g = (0.3 * mem[0] + rs.standard_normal(n)).astype(np.float32)
    if K > 2:
        mem[2] = 0.0                                   # a zero gradient: the clamped denominator gives 0, not NaN
    cos, mx = b.ops.grad_cosine(torch.from_numpy(mem).cuda(), torch.from_numpy(g).cuda())
    m64, g64 = mem.astype(np.float64), g.astype(np.float64)
    ref = (m64 @ g64) / np.maximum(np.linalg.norm(m64, axis=1) * np.linalg.norm(g64), 1e-8)
    np.testing.assert_allclose(cos.cpu().numpy(), ref, rtol=0, atol=2e-6)
    assert abs(float(mx) - ref.max()) <= 2e-6
    cos2, mx2 = b.ops.grad_cosine(torch.from_numpy(mem).cuda(), torch.from_numpy(g).cuda())
    assert torch.equal(cos, cos2) and torch.equal(mx, mx2)          # deterministic (fixed-order partials)


def _state(seed, small_head=True):
    spec = oresnet.Spec(32, 20, 10)
    p, bn = oresnet.seeded_state(spec, seed)
    if small_head:
        p['linear.weight'] = p['linear.weight'] * 0.02
        p['linear.bias'] = torch.zeros_like(p['linear.bias'])
    return spec, p, bn


def _load(model, spec, p, bn):
    model.engine.load(list(p.values()), [(bn[n + '.running_mean'], bn[n + '.running_var']) for n in oresnet.bn_names(spec)])


@pytest.mark.parametrize('N', [1, 3, 10])
def test_eval_mode_gradient(b, N):
    """forward_train(eval_stats=True) + backward(eval_stats=True) = the gradient of the network in eval mode
    (gss_greedy_update.py:16,80-83): every tensor within 1e-3 of the oracle, running statistics untouched."""
    spec, p, bn = _state(31 + N, small_head=False)
    params = SimpleNamespace(data='cifar10', agent='ER', head='mlp')
    model = b.nets.setup_architecture(params)
    _load(model, spec, p, bn)
    eng = model.engine
    rs = np.random.RandomState(N)
    x = torch.from_numpy(rs.rand(N, 3, 32, 32).astype(np.float32))
    y = torch.from_numpy(rs.randint(0, 10, N).astype(np.int64))
    stats_before = eng.state.bn_stats.clone()
    tracked_before = eng.state.bn_tracked.clone()
    logits, ws = eng.forward_train(x.cuda(), slot=5, eval_stats=True)
    ce = b.engine.ce_loss(logits, y.cuda(), want_grad=True)
    eng.backward(x.cuda(), ce['dlogits'], ws, eval_stats=True)
    assert torch.equal(eng.state.bn_stats, stats_before) and torch.equal(eng.state.bn_tracked, tracked_before)
    ref_logits = oresnet.forward(spec, p, bn, x, train=False)
    torch.testing.assert_close(logits.cpu(), ref_logits, rtol=1e-4, atol=1e-4)
    flat = ogss.eval_grad_vector(spec, p, bn, x, y)
    got = eng.state.grads.cpu()
    assert float((got - flat).norm() / flat.norm()) < 2e-4
    off = 0
    for name, v in p.items():
        n = v.numel()
        a, r = got[off:off + n], flat[off:off + n]
        off += n
        den = float(r.norm())
        if den > 1e-12:
            assert float((a - r).norm()) / den < 1e-3, name
    # a train-mode pass afterwards is unaffected by the eval-mode workspace (same engine, other slot)
    out_t, _ = eng.forward_train(x.cuda(), slot=0)
    bn2 = {k: v.clone() for k, v in bn.items()}
    ref_t = oresnet.forward(spec, p, bn2, x, train=True)
    if N > 1:
        torch.testing.assert_close(out_t.cpu(), ref_t, rtol=2e-3, atol=2e-3)


def test_gss_update_golden(b, golden_dir):
    """The plugin (registry.update_methods['GSS']) replays the reference run of gss.npz: fill phase with per-sample
    scores, two replacement lotteries and three updates that replace nothing."""
    g = np.load(os.path.join(golden_dir, 'gss.npz'))
    mem, batch = int(g['mem']), int(g['batch'])
    spec, p, bn = _state(int(g['model_seed']))
    params = SimpleNamespace(data='cifar10', agent='ER', head='mlp', cuda=True, mem_size=mem, update='GSS', retrieve='random',
                             gss_mem_strength=10, gss_batch_size=10, buffer_tracker=False, eps_mem_batch=10)
    model = b.nets.setup_architecture(params)
    _load(model, spec, p, bn)
    buf = b.memory.Buffer(model, params)
    upd = buf.update_method
    assert isinstance(upd, b.update.GSSGreedyUpdate)
    rs = np.random.RandomState(int(g['data_seed']))
    src = np.full((mem, 2), -1, dtype=np.int64)
    model.train()
    for u in range(g['y'].shape[0]):
        x = torch.from_numpy(rs.rand(batch, 3, 32, 32).astype(np.float32))
        rs.randint(0, 3, batch)
        y = torch.from_numpy(g['y'][u])
        torch.manual_seed(int(g['torch_seed0']) + u)
        # the replacement lottery runs on the scores' device: replay the reference's CPU draw for it
        before = buf.buffer_img.clone()
        orig = torch.multinomial

        def multinomial(probs, *a, **k):
            if probs.is_cuda:
                return orig(probs.cpu(), *a, **k).to(probs.device)
            return orig(probs, *a, **k)
        torch.multinomial = multinomial
        try:
            buf.update(x.cuda(), y.cuda(), y_host=g['y'][u])
        finally:
            torch.multinomial = orig
        changed = (buf.buffer_img != before).flatten(1).any(1).nonzero().flatten().tolist()
        for sl in changed:
            pos = [i for i in range(batch) if torch.equal(buf.buffer_img[sl].cpu(), x[i])]
            assert len(pos) == 1
            src[sl] = (u, pos[0])
        if not np.isnan(g['batch_sim'][u]):
            assert abs(upd.last_batch_sim - float(g['batch_sim'][u])) <= 1e-3, u
        np.testing.assert_array_equal(buf.buffer_label.cpu().numpy(), g['labels'][u], err_msg='update %d' % u)
        np.testing.assert_array_equal(buf.labels_host, g['labels'][u])
        np.testing.assert_array_equal(src, g['src'][u], err_msg='update %d' % u)
        np.testing.assert_allclose(upd.buffer_score.cpu().numpy(), g['scores'][u], rtol=0, atol=1e-3)
        upd.buffer_score.copy_(torch.from_numpy(g['scores'][u]))      # continue from the recorded scores
        assert model.training                                            # gss_greedy_update.py:64
    assert buf.current_index == mem
