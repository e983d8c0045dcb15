"""GPU parity of the plugin surface (retrieve / update plugins, agents) against the oracle and the
vectors recorded from the reference: bit-exact retrieved / evicted indices for ASER (up to exact
score ties, which the reference's unstable argsort leaves undefined), replay-step trajectories."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import aser as oaser
from oracle import knn_sv as oknn
from oracle import replay_step as ors
from oracle import resnet as oresnet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def b():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import b200ocl
    from b200ocl import engine, learners, memory, nets, ops, registry, retrieve, update
    return SimpleNamespace(engine=engine, learners=learners, memory=memory, nets=nets, ops=ops, registry=registry,
                           retrieve=retrieve, update=update)


def dev(a, dtype=None):
    t = torch.as_tensor(np.asarray(a))
    return (t.to(dtype) if dtype is not None else t).cuda()


def make_params(**kw):
    base = dict(data='cifar10', cuda=True, epoch=1, batch=10, verbose=False, mem_size=40, eps_mem_batch=10, mem_iters=1,
                update='random', retrieve='random', agent='ER', k=3, aser_type='asvm', n_smp_cls=1.5, num_tasks=5,
                buffer_tracker=False, optimizer='SGD', learning_rate=0.01, weight_decay=0, temp=0.07, head='mlp',
                subsample=20, error_analysis=False,
                trick={'labels_trick': False, 'kd_trick': False, 'separated_softmax': False, 'review_trick': False,
                       'ncm_trick': False, 'kd_trick_star': False})
    base.update(kw)
    return SimpleNamespace(**base)


def test_aser_decisions_golden(b, golden_dir):
    """ASER retrieve / update decisions on the reference's recorded sample sets (feature space)."""
    g = np.load(os.path.join(golden_dir, 'aser.npz'))
    n_exact = 0
    for i in range(int(g['n_cases'])):
        t = 'a%d_' % i
        bx, by, cur_x, cur_y = g[t + 'bx'], g[t + 'by'], g[t + 'cur_x'], g[t + 'cur_y']
        k, typ = int(g[t + 'k']), str(g[t + 'type'])
        cand, coop = g[t + 'ret_cand_ind'], g[t + 'ret_coop_ind']
        pos = b.retrieve.aser_retrieve_select(None, dev(cur_x), dev(cur_y), dev(bx[coop]) if len(coop) else None,
                                              dev(by[coop]) if len(coop) else None, dev(bx[cand]), dev(by[cand]), k, typ,
                                              10).cpu().numpy()
        sv_adv, _, _ = oknn.knn_sv_matrix(cur_x, cur_y, bx[cand], by[cand], k)
        sv_coop = oknn.knn_sv_matrix(bx[coop], by[coop], bx[cand], by[cand], k)[0] if typ != 'neg_sv' else None
        score = oaser.retrieve_score(sv_adv, sv_coop, typ)
        if oaser.min_adjacent_gap(score, 10) > 1e-6:
            np.testing.assert_array_equal(pos, g[t + 'ret_pos'])          # bit-exact indices, in rank order
            n_exact += 1
        else:
            assert oaser.rank_equivalent(pos, g[t + 'ret_pos'], score, atol=1e-6), i
        # update
        ev, ci = g[t + 'upd_eval_ind'], g[t + 'upd_cand_ind']
        mpos = oaser.minority_positions(cur_y, g[t + 'upd_counts_before'], int(g[t + 'mem']), float(g[t + 'upd_threshold']))
        eval_f, eval_y = np.concatenate([bx[ev], cur_x[mpos]]), np.concatenate([by[ev], cur_y[mpos]])
        cand_f, cand_y = np.concatenate([bx[ci], cur_x]), np.concatenate([by[ci], cur_y])
        sv_sum = b.ops.knn_sv(dev(eval_f), dev(eval_y), dev(cand_f), dev(cand_y), k)['sum']
        order = b.ops.rank_desc(sv_sum).cpu().numpy()
        ind_cur, ind_buf = b.update.aser_update_partition(order, len(ci), ci)
        srt = np.argsort(ind_buf)
        np.testing.assert_array_equal(ind_buf[srt], g[t + 'upd_changed_slots'])     # bit-exact evicted slots
        np.testing.assert_array_equal(cur_y[ind_cur][srt], g[t + 'upd_new_labels'])
        np.testing.assert_array_equal(cur_x[ind_cur][srt], g[t + 'upd_new_rows'])
    assert n_exact >= 6


def _fake_buffer(b, mem, row_shape):
    return SimpleNamespace(buffer_img=torch.zeros((mem,) + row_shape, device='cuda'),
                           buffer_label=torch.zeros(mem, dtype=torch.long, device='cuda'),
                           labels_host=np.zeros(mem, dtype=np.int64), current_index=0, n_seen_so_far=0,
                           write=None, gather=None)


def test_reservoir_update_golden(b, golden_dir, monkeypatch):
    g = np.load(os.path.join(golden_dir, 'reservoir.npz'))
    mem = int(g['mem'])
    params = make_params(mem_size=mem, cuda=True)
    buf = b.memory.Buffer.__new__(b.memory.Buffer)
    torch.nn.Module.__init__(buf)
    buf.register_buffer('buffer_img', torch.zeros(mem, 5, device='cuda'))
    buf.register_buffer('buffer_label', torch.zeros(mem, dtype=torch.long, device='cuda'))
    buf.labels_host = np.zeros(mem, dtype=np.int64)
    buf.current_index = buf.n_seen_so_far = 0
    upd = b.update.Reservoir_update(params)
    for s in range(g['x'].shape[0]):
        d = g['draws'][s]
        monkeypatch.setattr(b.update, 'reservoir_draws', lambda n, n_seen, device=None, d=d: d[d >= 0][:n])
        ret = upd.update(buf, dev(g['x'][s]), dev(g['y'][s]), y_host=g['y'][s])
        ref = g['rets'][s]
        assert list(ret) == [int(v) for v in ref[ref >= 0]], s
    assert buf.n_seen_so_far == int(g['n_seen'])
    np.testing.assert_array_equal(buf.buffer_img.cpu().numpy(), g['final_img'])
    np.testing.assert_array_equal(buf.buffer_label.cpu().numpy(), g['final_label'])
    np.testing.assert_array_equal(buf.labels_host, g['final_label'])


def test_mir_retrieve_golden(b, golden_dir, monkeypatch):
    g = np.load(os.path.join(golden_dir, 'resnet.npz'))
    spec = oresnet.Spec(32, 20, 100)
    params, bn = oresnet.seeded_state(spec, 11)
    model = b.nets.Reduced_ResNet18(100)
    model.engine.load(list(params.values()), [(bn[n + '.running_mean'], bn[n + '.running_var']) for n in oresnet.bn_names(spec)])
    eng = model.engine
    x, y = dev(g['cifar_x']), dev(g['cifar_y'])
    logits, ws = eng.forward_train(x)
    eng.backward(x, b.engine.ce_loss(logits, y)['dlogits'], ws)            # the stream-batch gradient MIR uses
    perm = g['mir_perm']
    sub_x, sub_y = dev(g['mir_sub_x'][perm]), dev(g['mir_sub_y'][perm])
    monkeypatch.setattr(b.retrieve, 'random_retrieve', lambda buffer, n: (sub_x, sub_y))
    retr = b.retrieve.MIR_retrieve(SimpleNamespace(subsample=12, eps_mem_batch=4, learning_rate=0.1))
    rx, ry = retr.retrieve(SimpleNamespace(model=model))
    pre, post = retr.last_scores
    scores = (post - pre).cpu().numpy()
    np.testing.assert_allclose(scores, g['mir_scores'], rtol=2e-3, atol=2e-5)
    np.testing.assert_array_equal(ry.cpu().numpy(), g['mir_sub_y'][perm][g['mir_top']])
    rm, _ = eng.bn_views()[0]
    np.testing.assert_allclose(rm.cpu().numpy(), g['mir_rm_after__bn1'], rtol=1e-4, atol=1e-6)


def _agent_and_oracle(b, seed, ncls, **kw):
    params = make_params(**kw)
    spec = oresnet.Spec(32, 20, ncls, head='mlp' if params.agent == 'SCR' else None)
    p0, bn0 = oresnet.seeded_state(spec, seed)
    model = b.nets.setup_architecture(params)
    model.engine.load(list(p0.values()), [(bn0[n + '.running_mean'], bn0[n + '.running_var']) for n in oresnet.bn_names(spec)])
    agent = b.registry.agents[params.agent](model, None, params)
    st = ors.ReplayState(spec, p0, bn0, params.mem_size, (3, 32, 32), ncls, lr=params.learning_rate)
    return params, agent, st


def _batches(seed, n_steps, n_lab):
    rs = np.random.RandomState(seed)
    for _ in range(n_steps):
        x = (rs.randint(0, 256, (10, 3, 32, 32)).astype(np.float32) / 255.0)
        yield torch.tensor(x), torch.tensor(rs.randint(0, n_lab, 10))


def _sync_weights(agent, st):
    """Load the oracle's weights and BN statistics into the engine: every step is then compared from an
    identical state (free-running trajectories of this net amplify 1e-7 differences ~3-70x per step)."""
    spec = st.spec
    agent.engine.load(list(st.params.values()),
                      [(st.bn[n + '.running_mean'], st.bn[n + '.running_var']) for n in oresnet.bn_names(spec)])


def _param_err(agent, st):
    flat = torch.cat([v.reshape(-1) for v in st.params.values()])
    return float((agent.engine.state.params.cpu() - flat).abs().max() / flat.abs().max())


def test_er_random_steps(b):
    """ER with uniform retrieval + reservoir update: same numpy / torch-CPU random streams as the
    reference, so retrieval and reservoir decisions coincide without injecting choices."""
    params, agent, st = _agent_and_oracle(b, 201, 10)
    agent.model.train()
    for i, (x, y) in enumerate(_batches(7, 7, 4)):
        np.random.seed(100 + i); torch.manual_seed(100 + i)
        agent.replay_step(x.cuda(), y.cuda(), y.numpy())
        gpu_loss = float(agent.last_loss)
        np.random.seed(100 + i); torch.manual_seed(100 + i)
        cpu_loss = ors.er_step(st, x, y, retrieve='random', update='random')
        assert abs(gpu_loss - cpu_loss) < 2e-4 * abs(cpu_loss), (i, gpu_loss, cpu_loss)
        assert _param_err(agent, st) < 2e-4, i
        np.testing.assert_array_equal(agent.buffer.buffer_label.cpu().numpy(), st.buffer_label.numpy())
        _sync_weights(agent, st)
    np.testing.assert_array_equal(agent.buffer.labels_host, st.buffer_label.numpy())
    torch.testing.assert_close(agent.buffer.buffer_img.cpu(), st.buffer_img, rtol=0, atol=0)
    assert agent.buffer.n_seen_so_far == st.n_seen_so_far


def test_er_mir_steps(b):
    params, agent, st = _agent_and_oracle(b, 202, 10, retrieve='MIR')
    agent.model.train()
    n_exact = 0
    for i, (x, y) in enumerate(_batches(8, 6, 4)):
        np.random.seed(200 + i); torch.manual_seed(200 + i)
        agent.replay_step(x.cuda(), y.cuda(), y.numpy())
        np.random.seed(200 + i); torch.manual_seed(200 + i)
        cpu_loss = ors.er_step(st, x, y, retrieve='MIR', update='random', subsample=20)
        assert abs(float(agent.last_loss) - cpu_loss) < 2e-4 * abs(cpu_loss), i
        if 'mir_scores' in st.log and i > 0:
            pre, post = agent.buffer.retrieve_method.last_scores
            # a score is the difference of two per-sample losses of ~2.0, each good to ~1e-4 absolute
            np.testing.assert_allclose((post - pre).cpu().numpy(), st.log['mir_scores'], rtol=2e-3, atol=5e-4)
            top = agent.buffer.retrieve_method.last_top.cpu().numpy()
            if set(top.tolist()) == set(st.log['mir_top'].tolist()):          # same ten samples replayed
                assert _param_err(agent, st) < 3e-4, i
                n_exact += 1
            else:   # a near-tie at the cut: every sample the two selections disagree on scores within tolerance of the cut
                sc = st.log['mir_scores']
                cut = np.sort(sc)[::-1][min(9, len(sc) - 1)]
                diff = set(top.tolist()) ^ set(st.log['mir_top'].tolist())
                assert all(abs(sc[j] - cut) < 1e-3 for j in diff), (i, diff)
        np.testing.assert_array_equal(agent.buffer.labels_host, st.buffer_label.numpy())
        _sync_weights(agent, st)


def test_er_aser_steps(b):
    """ER + ASER retrieve + ASER update: the sampler's choices are taken from the GPU run and replayed
    in the oracle; retrieved positions and evicted slots must be bit-exact wherever the oracle's scores
    are not within rounding of a tie (exact ties are frequent with four classes and leave the
    reference's own order undefined)."""
    params, agent, st = _agent_and_oracle(b, 203, 10, retrieve='ASER', update='ASER', mem_size=40)
    agent.model.train()
    exact_ret = exact_upd = 0
    for i, (x, y) in enumerate(_batches(9, 10, 4)):
        np.random.seed(300 + i); torch.manual_seed(300 + i)
        was_random = agent.buffer.n_seen_so_far <= params.mem_size
        will_update = agent.buffer.current_index + 10 >= params.mem_size
        agent.replay_step(x.cuda(), y.cuda(), y.numpy())
        ch = {}
        if was_random:
            ch['ret_idx'] = agent.buffer.last_random_idx
        else:
            ch.update(agent.buffer.retrieve_method.last_choices)
        if will_update:
            ch.update(agent.buffer.update_method.last_choices)
        np.random.seed(300 + i); torch.manual_seed(300 + i)
        cpu_loss = ors.er_step(st, x, y, retrieve='ASER', update='ASER', k=3, aser_type='asvm', n_smp_cls=1, choices=ch)
        same_replay = True
        if not was_random:
            pos = agent.buffer.retrieve_method.last_pos.cpu().numpy()
            if oaser.min_adjacent_gap(st.log['ret_score'], len(pos)) > 1e-6:
                np.testing.assert_array_equal(pos, st.log['ret_pos'])        # bit-exact retrieved indices
                exact_ret += 1
            else:
                same_replay = set(pos.tolist()) == set(np.asarray(st.log['ret_pos']).tolist())
        if same_replay:
            assert abs(float(agent.last_loss) - cpu_loss) < 2e-4 * abs(cpu_loss), (i, float(agent.last_loss), cpu_loss)
        if will_update:
            ind_cur, ind_buf = agent.buffer.update_method.last_decision
            sv = np.sort(st.log['upd_sv_sum'])[::-1]
            n_buf = len(st.log['upd_cand_ind'])
            boundary_gap = sv[n_buf - 1] - sv[n_buf] if len(sv) > n_buf else np.inf
            if boundary_gap > 1e-6:       # which samples enter / leave is decided away from a tie
                assert set(ind_cur.tolist()) == set(st.log['upd_ind_cur'].tolist()), i
                assert set(ind_buf.tolist()) == set(st.log['upd_ind_buffer'].tolist()), i
                exact_upd += 1
            if oaser.min_adjacent_gap(st.log['upd_sv_sum'], len(sv)) > 1e-6:
                np.testing.assert_array_equal(ind_cur, st.log['upd_ind_cur'])
                np.testing.assert_array_equal(ind_buf, st.log['upd_ind_buffer'])
            # pairing inside tie groups may differ: continue from the GPU's buffer
            st.buffer_img.copy_(agent.buffer.buffer_img.cpu()); st.buffer_label.copy_(agent.buffer.buffer_label.cpu())
            st.class_index_cache = {int(c): set(np.flatnonzero(agent.buffer.labels_host == c).tolist())
                                    for c in np.unique(agent.buffer.labels_host)}
            st.class_num_cache[:] = np.bincount(agent.buffer.labels_host, minlength=10)
        np.testing.assert_array_equal(agent.buffer.labels_host, agent.buffer.buffer_label.cpu().numpy())
        _sync_weights(agent, st)
    assert exact_ret >= 2 and exact_upd >= 2, (exact_ret, exact_upd)


def test_scr_steps(b):
    params, agent, st = _agent_and_oracle(b, 204, 100, agent='SCR', data='cifar100', eps_mem_batch=20)
    agent.transform = torch.nn.Identity()            # kornia is unpinned: identical second view on both sides
    agent.model.train()
    for i, (x, y) in enumerate(_batches(10, 6, 4)):
        np.random.seed(400 + i); torch.manual_seed(400 + i)
        agent.replay_step(x.cuda(), y.cuda(), y.numpy())
        np.random.seed(400 + i); torch.manual_seed(400 + i)
        cpu_loss = ors.scr_step(st, x, y, eps_mem_batch=20, temperature=0.07)
        if cpu_loss is not None:
            assert abs(float(agent.last_loss) - cpu_loss) < 2e-4 * abs(cpu_loss), (i, float(agent.last_loss), cpu_loss)
            assert _param_err(agent, st) < 5e-4, i
        np.testing.assert_array_equal(agent.buffer.buffer_label.cpu().numpy(), st.buffer_label.numpy())
        _sync_weights(agent, st)


def test_scr_augment_kernel(b):
    from b200ocl.augment import SCRTransform, draw_params
    x = torch.rand(16, 3, 32, 32, device='cuda')
    p = draw_params(16, 32, 32, rng=np.random.RandomState(0))
    p[0] = [0, 0, 32, 32, 0, 0, 0, 1, 1, 0, 0, 0]            # identity parameters
    p[1] = [0, 0, 32, 32, 1, 0, 0, 1, 1, 0, 0, 0]            # pure horizontal flip
    p[2] = [0, 0, 32, 32, 0, 0, 0, 1, 1, 0, 0, 1]            # pure grayscale
    out = SCRTransform((32, 32))(x, params=p)
    torch.testing.assert_close(out[0], x[0], rtol=0, atol=1e-6)
    torch.testing.assert_close(out[1], x[1].flip(-1), rtol=0, atol=1e-6)
    gray = 0.299 * x[2, 0] + 0.587 * x[2, 1] + 0.114 * x[2, 2]
    torch.testing.assert_close(out[2, 1], gray, rtol=0, atol=1e-6)
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0 + 1e-6 and torch.isfinite(out).all()
    out2 = SCRTransform((32, 32))(x)                           # random parameters: shape / range only
    assert out2.shape == x.shape and torch.isfinite(out2).all()


@pytest.mark.parametrize('hw,n,seed', [(32, 64, 0), (84, 12, 1), (32, 220, 2)])
def test_scr_augment_kernel_vs_oracle(b, hw, n, seed):
    """csrc/augment.cu with drawn parameters (crop boxes, flips, the four colour operations in every order, grayscale)
    against oracle/augment.py (float64; itself checked against grid_sample and colorsys on the CPU).  kornia is absent:
    the pipeline's definition is parity-unpinned, the kernel's arithmetic is not."""
    from b200ocl.augment import SCRTransform, draw_params
    from oracle import augment as oaug
    rs = np.random.RandomState(seed)
    x = rs.rand(n, 3, hw, hw).astype(np.float32)
    x[0, :, :4, :4] = 0.5                                     # grey and black patches (hue undefined / value 0)
    x[1, :, :4, :4] = 0.0
    p = draw_params(n, hw, hw, rng=rs)
    p[:, 5] = 1                                               # colour jitter on for every sample
    for i in range(n):                                        # every sample its own operation order
        order = rs.permutation(4)
        p[i, 10] = float(sum(int(op) << (2 * k) for k, op in enumerate(order)))
    p[::3, 5] = rs.rand(len(p[::3])) < 0.5
    out = SCRTransform((hw, hw))(torch.from_numpy(x).cuda(), params=p).cpu().numpy()
    ref = oaug.scr_view(x, p)
    err = np.abs(out - ref)
    assert err.max() <= 5e-5, (err.max(), np.unravel_index(err.argmax(), err.shape))
    assert err.mean() <= (2e-7 if hw <= 32 else 3e-6)     # fp32 sampling coordinates: ulp(83) = 7.6e-6 pixels on noise images
    assert (p[:, 2] < hw).any() and (p[:, 4] > 0.5).any() and (p[:, 11] > 0.5).any()     # the draws exercised crop / flip / gray


@pytest.mark.parametrize('kind', ['aser', 'scr'])
def test_concurrent_step_equals_sequential_step(b, kind):
    """learners._CONCURRENT (two streams, deferred running statistics, second gradient arena) against the plain sequential step: same memory, same random draws, weights bit
    for bit, running statistics to rounding."""
    from b200ocl import learners
    res = {}
    for mode in (False, True):
        learners.set_concurrent(mode)
        try:
            kw = dict(retrieve='ASER', update='ASER', n_smp_cls=1) if kind == 'aser' else \
                dict(agent='SCR', eps_mem_batch=20, data='cifar100', mem_size=60)
            b.memory.ClassBalancedRandomSampling.reset()
            params, agent, st = _agent_and_oracle(b, 260, 10 if kind == 'aser' else 100, **kw)
            if kind == 'scr':
                from b200ocl.augment import Identity
                agent.transform = Identity()
            agent.model.train()
            np.random.seed(77); torch.manual_seed(77)
            for i, (x, y) in enumerate(_batches(12, 9, 4)):
                agent.replay_step(x.cuda(), y.cuda(), y.numpy())
            torch.cuda.synchronize()
            res[mode] = (agent.engine.state.params.cpu(), agent.engine.state.bn_stats.cpu(), agent.buffer.buffer_label.cpu(),
                         agent.buffer.buffer_img.cpu(), agent.buffer.labels_host.copy(), np.random.get_state()[1].copy(),
                         torch.get_rng_state().clone(), agent.engine.state.bn_tracked.cpu())
        finally:
            learners.set_concurrent(True)
    s, c = res[False], res[True]
    assert torch.equal(s[2], c[2]) and torch.equal(s[3], c[3]) and np.array_equal(s[4], c[4])
    assert np.array_equal(s[5], c[5]) and torch.equal(s[6], c[6]) and torch.equal(s[7], c[7])
    torch.testing.assert_close(c[1], s[1], rtol=1e-6, atol=1e-9)
    if kind == 'scr':
        assert torch.equal(s[0], c[0])                       # train-mode arithmetic does not read the running statistics
    else:
        # ASER's eval features read them: a one-rounding difference may reach the weights through a different retrieval
        # only at an exact score tie; the memory is identical above, so the weights are too
        assert torch.equal(s[0], c[0])
