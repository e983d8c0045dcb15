"""CPU: the oracle restatement agrees with the vectors recorded from the reference
itself (tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import aser as oaser
from oracle import knn_sv as oknn
from oracle import resnet as oresnet
from oracle import supcon as osup


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_knn_sv_matches_reference(golden_dir):
    g = _load(golden_dir, 'knn_sv.npz')
    for i in range(int(g['n_cases'])):
        ef, cf, ey, cy, k = (g['c%d_%s' % (i, n)] for n in ('ef', 'cf', 'ey', 'cy', 'k'))
        sv64, order64, dist64 = oknn.knn_sv_matrix(ef, ey, cf, cy, int(k), dtype=np.float64)
        # ordering: identical to the reference's fp32 argsort except fp32 near-ties
        n_diff, n_bad = oknn.order_mismatch_explained(order64, g['c%d_order' % i], dist64)
        assert n_bad == 0, (i, n_diff, n_bad)
        if n_diff == 0:
            np.testing.assert_allclose(sv64, g['c%d_sv' % i], rtol=0, atol=2e-6)
        sv32, _, _ = oknn.knn_sv_matrix(ef, ey, cf, cy, int(k), dtype=np.float32)
        if n_diff == 0:
            np.testing.assert_allclose(sv32, g['c%d_sv' % i], rtol=0, atol=2e-6)


def test_knn_sv_row_loop_cross_check():
    rs = np.random.RandomState(0)
    for (E, C, d, k) in [(4, 1, 5, 3), (4, 2, 5, 3), (5, 9, 7, 3), (3, 17, 4, 1), (3, 6, 4, 10)]:
        ef, cf = rs.rand(E, d), rs.rand(C, d)
        ey, cy = rs.randint(0, 3, E), rs.randint(0, 3, C)
        sv, order, dist = oknn.knn_sv_matrix(ef, ey, cf, cy, k)
        for r in range(E):
            np.testing.assert_allclose(sv[r], oknn.knn_sv_row_loop(dist[r], ey[r], cy, k), atol=1e-12)
        if C >= k:   # Shapley efficiency: row sum == kNN utility of the full candidate set
            for r in range(E):
                util = (cy[order[r, :k]] == ey[r]).sum() / k
                assert abs(sv[r].sum() - util) < 1e-12


def test_supcon_matches_reference(golden_dir):
    g = _load(golden_dir, 'supcon.npz')
    for i in range(int(g['n_cases'])):
        f, y, T = g['c%d_f' % i], g['c%d_y' % i], float(g['c%d_T' % i])
        loss, grad = osup.supcon_loss_and_grad(f, y, T)
        ref_loss, ref_grad = g['c%d_loss' % i], g['c%d_grad' % i]
        if np.isnan(ref_loss):
            assert np.isnan(loss)
            continue
        np.testing.assert_allclose(loss, ref_loss, rtol=2e-5)
        np.testing.assert_allclose(grad, ref_grad, rtol=1e-3, atol=2e-6)


def test_supcon_errors():
    with pytest.raises(ValueError):
        osup.supcon_loss_and_grad(np.zeros((4, 8)), np.zeros(4, dtype=int))
    with pytest.raises(ValueError):
        osup.supcon_loss_and_grad(np.zeros((4, 2, 8)), np.zeros(5, dtype=int))


def _grad_check(g, tag, grads, prefix_names):
    names = [str(n) for n in g[tag + '_grad_names']]
    norms = g[tag + '_grad_norms']
    for n, ref in zip(names, norms):
        got = 0.0 if grads[n] is None else float(grads[n].double().norm())
        assert abs(got - ref) <= 1e-3 * max(ref, 1e-6) + 1e-7, (n, got, ref)
    for key in g.files:
        if key.startswith(tag + '_grad__'):
            n = key[len(tag + '_grad__'):]
            got = grads[n].numpy()
            ref = g[key]
            if got.size > 30000:
                got = got.reshape(got.shape[0], -1)[:8]
            np.testing.assert_allclose(got, ref, rtol=2e-3, atol=1e-5 * np.abs(ref).max())


def test_resnet_cifar_matches_reference(golden_dir):
    g = _load(golden_dir, 'resnet.npz')
    spec = oresnet.Spec(32, 20, 100)
    params, bn = oresnet.seeded_state(spec, 11)
    x, y = torch.tensor(g['cifar_x']), torch.tensor(g['cifar_y'])
    with torch.no_grad():
        feat = oresnet.features(spec, params, bn, x, train=False)
        logits = oresnet.forward(spec, params, bn, x, train=False)
    np.testing.assert_allclose(feat.numpy(), g['cifar_feat_eval'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(logits.numpy(), g['cifar_logits_eval'], rtol=1e-4, atol=1e-5)
    loss, logits_t, grads = oresnet.ce_loss_and_grads(spec, params, bn, x, y)
    np.testing.assert_allclose(logits_t.numpy(), g['cifar_logits_train'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(float(loss), float(g['cifar_loss']), rtol=1e-5)
    _grad_check(g, 'cifar', grads, None)
    for k in ['bn1', 'layer1.0.bn2', 'layer2.0.shortcut.1', 'layer4.1.bn2']:
        np.testing.assert_allclose(bn[k + '.running_mean'].numpy(), g['cifar_rm__' + k], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(bn[k + '.running_var'].numpy(), g['cifar_rv__' + k], rtol=1e-4, atol=1e-6)
    # MIR scores on the same state
    sub_x, sub_y = torch.tensor(g['mir_sub_x'])[g['mir_perm']], torch.tensor(g['mir_sub_y'])[g['mir_perm']]
    scores = oresnet.mir_scores(spec, params, bn, grads, 0.1, sub_x, sub_y)
    np.testing.assert_allclose(scores.numpy(), g['mir_scores'], rtol=2e-3, atol=2e-5)
    assert np.array_equal(np.argsort(-scores.numpy(), kind='stable')[:4], g['mir_top'])
    np.testing.assert_allclose(bn['bn1.running_mean'].numpy(), g['mir_rm_after__bn1'], rtol=1e-4, atol=1e-6)


def test_resnet_mini_matches_reference(golden_dir):
    g = _load(golden_dir, 'resnet.npz')
    spec = oresnet.Spec(84, 20, 100)
    assert spec.dim_in == 640 and spec.final_hw == 11 and spec.pooled_hw == 2
    params, bn = oresnet.seeded_state(spec, 12)
    x = torch.tensor(g['mini_x'])
    with torch.no_grad():
        feat = oresnet.features(spec, params, bn, x, train=False)
    np.testing.assert_allclose(feat.numpy(), g['mini_feat_eval'], rtol=1e-4, atol=1e-5)
    loss, logits_t, grads = oresnet.ce_loss_and_grads(spec, params, bn, x, torch.tensor(g['mini_y']))
    np.testing.assert_allclose(logits_t.numpy(), g['mini_logits_train'], rtol=1e-4, atol=1e-5)
    _grad_check(g, 'mini', grads, None)


def test_supcon_resnet_matches_reference(golden_dir):
    g = _load(golden_dir, 'resnet.npz')
    spec = oresnet.Spec(32, 20, 100, head='mlp')
    params, bn = oresnet.seeded_state(spec, 13)
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    x1, x2, y = torch.tensor(g['scr_x1']), torch.tensor(g['scr_x2']), g['scr_y']
    f1 = oresnet.forward(spec, leaves, bn, x1, train=True)
    f2 = oresnet.forward(spec, leaves, bn, x2, train=True)
    feats = torch.stack([f1, f2], dim=1)
    np.testing.assert_allclose(feats.detach().numpy(), g['scr_feats'], rtol=1e-4, atol=1e-6)
    loss, dfeat = osup.supcon_loss_and_grad(feats.detach().numpy(), y, 0.07)
    np.testing.assert_allclose(loss, float(g['scr_loss']), rtol=1e-5)
    feats.backward(torch.tensor(dfeat, dtype=torch.float32))
    grads = {k: v.grad for k, v in leaves.items()}
    _grad_check(g, 'scr', grads, None)
    np.testing.assert_allclose(bn['encoder.bn1.running_mean'].numpy(), g['scr_rm__encoder.bn1'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(bn['encoder.bn1.running_var'].numpy(), g['scr_rv__encoder.bn1'], rtol=1e-4, atol=1e-6)


def test_aser_decisions_match_reference(golden_dir):
    g = _load(golden_dir, 'aser.npz')
    for i in range(int(g['n_cases'])):
        t = 'a%d_' % i
        bx, by, cur_x, cur_y = g[t + 'bx'], g[t + 'by'], g[t + 'cur_x'], g[t + 'cur_y']
        k, typ, mem, ncls = int(g[t + 'k']), str(g[t + 'type']), int(g[t + 'mem']), int(g[t + 'ncls'])
        # retrieve
        cand = g[t + 'ret_cand_ind']
        sv_adv, _, _ = oknn.knn_sv_matrix(cur_x, cur_y, bx[cand], by[cand], k)
        sv_coop = None
        if typ != 'neg_sv':
            coop = g[t + 'ret_coop_ind']
            assert len(set(coop.tolist()) & set(cand.tolist())) == 0
            sv_coop, _, _ = oknn.knn_sv_matrix(bx[coop], by[coop], bx[cand], by[cand], k)
        pos = oaser.retrieve_indices(sv_adv, sv_coop, typ, 10)
        score = oaser.retrieve_score(sv_adv, sv_coop, typ)
        if oaser.min_adjacent_gap(score, 10) > 1e-6:      # no tie: bit-exact indices and order
            assert np.array_equal(pos, g[t + 'ret_pos']), (i, pos, g[t + 'ret_pos'])
            np.testing.assert_array_equal(by[cand][pos], g[t + 'ret_y'])
        else:                                              # exact ties: equal up to tie order
            assert oaser.rank_equivalent(pos, g[t + 'ret_pos'], score, atol=1e-6), (i, pos, g[t + 'ret_pos'])
        # update
        ev, ci = g[t + 'upd_eval_ind'], g[t + 'upd_cand_ind']
        mpos = oaser.minority_positions(cur_y, g[t + 'upd_counts_before'], mem, float(g[t + 'upd_threshold']))
        assert len(mpos) == int(g[t + 'upd_n_minority'])
        eval_f = np.concatenate([bx[ev], cur_x[mpos]])
        eval_y = np.concatenate([by[ev], cur_y[mpos]])
        cand_f = np.concatenate([bx[ci], cur_x])
        cand_y = np.concatenate([by[ci], cur_y])
        sv, _, _ = oknn.knn_sv_matrix(eval_f, eval_y, cand_f, cand_y, k)
        ind_cur, ind_buf = oaser.update_partition(sv.sum(0), len(ci), ci)
        assert len(ind_cur) == len(ind_buf)
        order = np.argsort(ind_buf)
        assert np.array_equal(ind_buf[order], g[t + 'upd_changed_slots']), i
        np.testing.assert_array_equal(cur_y[ind_cur][order], g[t + 'upd_new_labels'])
        np.testing.assert_array_equal(cur_x[ind_cur][order], g[t + 'upd_new_rows'])


def test_reservoir_matches_reference(golden_dir):
    g = _load(golden_dir, 'reservoir.npz')
    mem = int(g['mem'])
    img = np.zeros((mem, 5), np.float32)
    lab = np.zeros(mem, np.int64)
    cur = seen = 0
    for s in range(g['x'].shape[0]):
        x, y = g['x'][s], g['y'][s]
        place = max(0, mem - cur)
        off = min(place, 10)
        img[cur:cur + off], lab[cur:cur + off] = x[:off], y[:off]
        written = list(range(cur, cur + off))
        cur += off
        seen += off
        if off < 10:
            x, y = x[place:], y[place:]
            draws = g['draws'][s][:len(x)]
            slots, src = oaser.reservoir_slots(draws, mem)
            seen += len(x)
            img[slots], lab[slots] = x[src], y[src]
            written = slots if off == 0 else slots
        ref = g['rets'][s]
        assert written == [int(v) for v in ref[ref >= 0]], s
    assert seen == int(g['n_seen'])
    np.testing.assert_array_equal(img, g['final_img'])
    np.testing.assert_array_equal(lab, g['final_label'])


def test_gss_greedy_matches_reference(golden_dir):
    """oracle/gss.py replays the reference GSSGreedyUpdate run recorded in gss.npz (fill phase, two replacement
    lotteries, three updates that replace nothing): same labels, same slot sources, scores within 1e-4."""
    from oracle import gss as ogss
    g = _load(golden_dir, 'gss.npz')
    mem, batch = int(g['mem']), int(g['batch'])
    spec = oresnet.Spec(32, 20, 10)
    p, bn = oresnet.seeded_state(spec, int(g['model_seed']))
    p['linear.weight'] = p['linear.weight'] * 0.02          # the two lines make_golden.gen_gss applies to the reference model
    p['linear.bias'] = torch.zeros_like(p['linear.bias'])
    st = ogss.GSSState(spec, p, bn, mem, (3, 32, 32))
    rs = np.random.RandomState(int(g['data_seed']))
    src = np.full((mem, 2), -1, dtype=np.int64)
    n_neg = 0
    for u in range(g['y'].shape[0]):
        x = torch.from_numpy(rs.rand(batch, 3, 32, 32).astype(np.float32))
        rs.randint(0, 3, batch)                            # the generator drew the labels from the same stream
        y = torch.from_numpy(g['y'][u])
        torch.manual_seed(int(g['torch_seed0']) + u)
        for pos, slot in ogss.update(st, x, y):
            src[slot] = (u, pos)
        np.testing.assert_array_equal(st.buffer_label.numpy(), g['labels'][u], err_msg='update %d' % u)
        np.testing.assert_array_equal(src, g['src'][u], err_msg='update %d' % u)
        np.testing.assert_allclose(st.buffer_score.numpy(), g['scores'][u], rtol=0, atol=1e-4)
        if not np.isnan(g['batch_sim'][u]):
            assert abs(st.last_batch_sim - float(g['batch_sim'][u])) <= 1e-4
            n_neg += st.last_batch_sim < 0
    assert n_neg == 2 and st.current_index == mem
