"""CPU: the class-balanced sampler's slot table stays consistent with the reference-style caches under random
relabelling, and its draws have the semantics of ClassBalancedRandomSampling.sample (buffer_utils.py:81-121)."""
import numpy as np


def check_consistency(CB, n_classes):
    lab, member = CB.labels_host, CB._member
    for c in range(CB._tab.shape[0]):
        k = int(CB._cnt[c])
        slots = CB._tab[c, :k]
        expect = set(np.flatnonzero(member & (lab == c)).tolist())
        assert set(slots.tolist()) == expect
        assert len(set(slots.tolist())) == k
        assert (CB._pos[slots] == np.arange(k)).all()
        if c in CB.class_index_cache:
            assert CB.class_index_cache[c] == expect
        else:
            assert not expect


def test_table_follows_incremental_updates():
    from b200ocl.memory import ClassBalancedRandomSampling as CB
    rs = np.random.RandomState(3)
    mem, ncls = 300, 12
    CB.reset()
    lab = rs.randint(0, ncls, mem)
    CB.update_cache(np.zeros(mem), ncls, new_y=lab[:200], ind=np.arange(200))      # fill phase
    check_consistency(CB, ncls)
    assert int(CB.class_num_cache.sum()) == 200
    for _ in range(60):                                                             # replacement phase
        k = rs.randint(1, 12)
        ind = rs.choice(mem, k, replace=False)
        CB.update_cache(np.zeros(mem), ncls, new_y=rs.randint(0, ncls, k), ind=ind)
        check_consistency(CB, ncls)
    CB.reset()


def test_draw_semantics():
    from b200ocl.memory import ClassBalancedRandomSampling as CB
    rs = np.random.RandomState(4)
    mem, ncls = 500, 10
    CB.reset()
    lab = rs.randint(0, ncls, mem)
    lab[lab == 7] = 6                       # class 7 absent
    lab[:3] = 9
    CB.update_cache(np.zeros(mem), ncls, labels_host=lab)
    check_consistency(CB, ncls)
    for n in (1, 2, 5, 1000):
        idx = CB.sample_indices(n)
        assert len(set(idx.tolist())) == len(idx)
        counts = np.bincount(lab[idx], minlength=ncls)
        full = np.bincount(lab, minlength=ncls)
        assert (counts == np.minimum(full, n)).all()
        assert (np.diff(lab[idx]) >= 0).all()                      # class-major output
        excl = idx[::2]
        idx2 = CB.sample_indices(n, excl_indices=excl)
        assert not set(idx2.tolist()) & set(excl.tolist())
        left = full - np.bincount(lab[excl], minlength=ncls)
        assert (np.bincount(lab[idx2], minlength=ncls) == np.minimum(left, n)).all()
    # every member of a class is reachable, roughly uniformly
    hits = np.zeros(mem)
    for _ in range(400):
        hits[CB.sample_indices(1)] += 1
    cls0 = np.flatnonzero(lab == 0)
    assert hits[cls0].min() > 0 and hits[cls0].max() < 6 * hits[cls0].mean()
    CB.reset()
