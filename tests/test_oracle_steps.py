"""CPU: oracle/replay_step.py replays the reference agents' recorded minibatches decision by
decision (tests/golden/steps.npz, written by running agents/exp_replay.py and agents/scr.py)."""
import os

import numpy as np
import torch

from oracle import replay_step as ors
from oracle import resnet as oresnet


def _strip(a):
    a = np.asarray(a)
    return a[a >= 0]


def _replay(g, tag, agent, retrieve, update, ncls, eps, seed):
    spec = oresnet.Spec(32, 20, ncls, head='mlp' if agent == 'SCR' else None)
    params, bn = oresnet.seeded_state(spec, 100 + seed)
    st = ors.ReplayState(spec, params, bn, 40, (3, 32, 32), ncls, lr=0.01)
    bx, by = g[tag + '_bx'], g[tag + '_by']
    i_cbrs = i_upd = 0
    for i in range(bx.shape[0]):
        x = torch.tensor(bx[i]).float().div(255)
        y = torch.tensor(by[i])
        ch = {}
        if tag + '_draws' in g.files:
            ch['reservoir_draws'] = _strip(g[tag + '_draws'][i])
        if agent == 'SCR':
            ch['ret_idx'] = _strip(g[tag + '_ret_idx'][i])
            ors.scr_step(st, x, y, eps_mem_batch=eps, temperature=0.07, choices=ch)
            continue
        if retrieve == 'random':
            ch['ret_idx'] = _strip(g[tag + '_ret_idx'][i])
        elif retrieve == 'MIR':
            ch['mir_idx'] = _strip(g[tag + '_mir_idx'][i])
        else:
            if st.n_seen_so_far <= st.mem_size:
                ch['ret_idx'] = _strip(g[tag + '_ret_idx'][i])
            else:
                ch['ret_cand_ind'] = g['%s_cbrs%d' % (tag, i_cbrs)]
                ch['ret_coop_ind'] = g['%s_cbrs%d' % (tag, i_cbrs + 1)]
                i_cbrs += 2
        if update == 'ASER' and st.current_index + 10 >= st.mem_size:   # an exactly-filling batch still runs the SV update (aser_update.py:38-41)
            ch['upd_eval_ind'] = g['%s_cbrs%d' % (tag, i_cbrs)]
            i_cbrs += 1
            ch['upd_cand_ind'] = g['%s_updcand%d' % (tag, i_upd)]
            ch['upd_threshold'] = float(g[tag + '_thr'][i_upd])
            i_upd += 1
        ors.er_step(st, x, y, retrieve=retrieve, update=update, eps_mem_batch=eps, k=3, aser_type='asvm',
                    n_smp_cls=1, subsample=20, choices=ch)
    if update == 'ASER':
        assert i_cbrs == int(g[tag + '_n_cbrs']) and i_upd == int(g[tag + '_n_updcand'])
    return st


def _check(g, tag, st):
    assert st.n_seen_so_far == int(g[tag + '_n_seen'])
    np.testing.assert_array_equal(st.buffer_label.numpy(), g[tag + '_final_labels'])
    np.testing.assert_allclose(st.buffer_img.numpy().reshape(40, -1).sum(1), g[tag + '_final_img_sum'], rtol=1e-5)
    names = [str(n) for n in g[tag + '_param_names']]
    for n, ref in zip(names, g[tag + '_param_norms']):
        got = float(st.params[n].double().norm())
        assert abs(got - ref) <= 2e-3 * max(ref, 1e-6), (tag, n, got, ref)
    bn1 = [k for k in st.bn if k.endswith('bn1.running_mean')][0]
    np.testing.assert_allclose(st.bn[bn1].numpy(), g[tag + '_rm_bn1'], rtol=1e-3, atol=1e-5)


def test_er_random_replays_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'steps.npz'))
    _check(g, 'er', _replay(g, 'er', 'ER', 'random', 'random', 10, 10, 1))


def test_er_mir_replays_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'steps.npz'))
    _check(g, 'mir', _replay(g, 'mir', 'ER', 'MIR', 'random', 10, 10, 2))


def test_er_aser_replays_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'steps.npz'))
    _check(g, 'aser', _replay(g, 'aser', 'ER', 'ASER', 'ASER', 10, 10, 3))


def test_scr_replays_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'steps.npz'))
    _check(g, 'scr', _replay(g, 'scr', 'SCR', 'random', 'random', 100, 20, 4))
