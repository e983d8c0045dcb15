"""B200OCL_MODE=parity: the host-side random decisions consume the reference's generators call for call
(VERDICT r01 item 9).  CPU tests against the unmodified reference classes (baseline/_ref or /root/reference)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'baseline'))
import ref_harness  # noqa: E402

pytestmark = pytest.mark.skipif(ref_harness.locate() is None, reason='no reference tree (baseline/_ref)')


@pytest.fixture()
def ref():
    ref_harness.import_reference()
    from utils.buffer import buffer_utils, reservoir_update
    from b200ocl import memory
    memory.set_mode(True, 'cpu')
    buffer_utils.ClassBalancedRandomSampling.class_index_cache = None
    buffer_utils.ClassBalancedRandomSampling.class_num_cache = None
    memory.ClassBalancedRandomSampling.reset()
    yield SimpleNamespace(bu=buffer_utils, ru=reservoir_update)
    memory.set_mode(False)
    memory.ClassBalancedRandomSampling.reset()
    buffer_utils.ClassBalancedRandomSampling.class_index_cache = None
    buffer_utils.ClassBalancedRandomSampling.class_num_cache = None


def _both_sample(ref, mine, buf_x, buf_y, n, excl, seed):
    torch.manual_seed(seed)
    _, _, r = ref.sample(buf_x, buf_y, n, excl_indices=None if excl is None else set(excl), device='cpu')
    torch.manual_seed(seed)
    m = mine.sample_indices(n, excl_indices=excl)
    return r.numpy(), m


@pytest.mark.parametrize('n_cls,mem', [(10, 200), (100, 1000)])
def test_sampler_incremental_stream_identical(ref, n_cls, mem):
    from b200ocl.memory import ClassBalancedRandomSampling as Mine
    Ref = ref.bu.ClassBalancedRandomSampling
    rs = np.random.RandomState(3)
    labels = rs.randint(0, n_cls, mem).astype(np.int64)
    buf_y = torch.zeros(mem, dtype=torch.int64)
    buf_x = torch.zeros(mem, 1)
    # fill phase in batches of 10, as ASER_update does (aser_update.py:28-35)
    for s in range(0, mem, 10):
        ind = torch.arange(s, s + 10)
        ny = torch.from_numpy(labels[s:s + 10])
        Ref.update_cache(buf_y, n_cls, new_y=ny, ind=ind, device='cpu')
        Mine.update_cache(buf_y, n_cls, new_y=labels[s:s + 10], ind=np.arange(s, s + 10))
        buf_y[s:s + 10] = ny
    assert list(Ref.class_index_cache.keys()) == list(Mine.class_index_cache.keys())
    for step in range(25):
        r1, m1 = _both_sample(Ref, Mine, buf_x, buf_y, 1, None, 100 + step)
        assert np.array_equal(r1, m1)
        r2, m2 = _both_sample(Ref, Mine, buf_x, buf_y, 2, r1.tolist(), 200 + step)
        assert np.array_equal(r2, m2)
        # replacement of a few slots (aser_update.py:108-112)
        slots = rs.choice(mem, 4, replace=False)
        ny = rs.randint(0, n_cls, 4).astype(np.int64)
        Ref.update_cache(buf_y, n_cls, new_y=torch.from_numpy(ny), ind=torch.from_numpy(slots), device='cpu')
        Mine.update_cache(buf_y, n_cls, new_y=ny, ind=slots)
        buf_y[torch.from_numpy(slots)] = torch.from_numpy(ny)
        assert np.array_equal(Ref.class_num_cache.numpy(), Mine.class_num_cache)


def test_sampler_rebuild_stream_identical(ref):
    from b200ocl.memory import ClassBalancedRandomSampling as Mine
    Ref = ref.bu.ClassBalancedRandomSampling
    rs = np.random.RandomState(4)
    labels = rs.randint(0, 20, 300).astype(np.int64)
    buf_y = torch.from_numpy(labels)
    buf_x = torch.zeros(300, 1)
    Ref.update_cache(buf_y, 20)
    Mine.update_cache(buf_y, 20, labels_host=labels)
    assert list(Ref.class_index_cache.keys()) == list(Mine.class_index_cache.keys())
    for step in range(10):
        r, m = _both_sample(Ref, Mine, buf_x, buf_y, 3, None, step)
        assert np.array_equal(r, m)
        r2, m2 = _both_sample(Ref, Mine, buf_x, buf_y, 3, r.tolist(), 50 + step)
        assert np.array_equal(r2, m2)


def test_reservoir_stream_identical(ref):
    from b200ocl.memory import Buffer
    from b200ocl import registry
    params = SimpleNamespace(data='cifar10', cuda=False, mem_size=50, update='random', retrieve='random',
                             eps_mem_batch=10, buffer_tracker=False)
    mine = Buffer(None, params)
    theirs = SimpleNamespace(buffer_img=torch.zeros(50, 3, 32, 32), buffer_label=torch.zeros(50, dtype=torch.int64),
                             current_index=0, n_seen_so_far=0, params=params)
    ref_upd = ref.ru.Reservoir_update(params)
    rs = np.random.RandomState(0)
    for step in range(30):
        x = torch.from_numpy(rs.rand(10, 3, 32, 32).astype(np.float32))
        y = torch.from_numpy(rs.randint(0, 10, 10).astype(np.int64))
        torch.manual_seed(step)
        a = ref_upd.update(theirs, x, y)
        torch.manual_seed(step)
        b = mine.update(x, y)
        assert list(a) == list(b)
        assert torch.equal(theirs.buffer_label, mine.buffer_label)
        assert torch.equal(theirs.buffer_img, mine.buffer_img)
        assert theirs.n_seen_so_far == mine.n_seen_so_far


def test_stream_order_identical(ref):
    from continuum.data_utils import dataset_transform
    from utils.setup_elements import transforms_match
    from torch.utils import data
    from b200ocl.learners import StreamFeeder
    rs = np.random.RandomState(1)
    x = rs.randint(0, 256, (57, 32, 32, 3)).astype(np.uint8)
    y = rs.randint(0, 10, 57).astype(np.int64)
    torch.manual_seed(11)
    loader = data.DataLoader(dataset_transform(x, y, transform=transforms_match['cifar10']), batch_size=10,
                             shuffle=True, num_workers=0, drop_last=True)
    ref_batches = [(bx.clone(), by.clone()) for bx, by in loader]
    after_ref = torch.rand(1)
    torch.manual_seed(11)
    mine = list(StreamFeeder(x, y, 10, 'cpu'))
    after_mine = torch.rand(1)
    assert len(mine) == len(ref_batches) == 5
    for (rx, ry), (mx, my, myh) in zip(ref_batches, mine):
        assert torch.equal(ry, my) and np.array_equal(ry.numpy(), myh)
        assert torch.allclose(rx, mx, atol=0, rtol=0)
    assert torch.equal(after_ref, after_mine)          # the default generator ends in the same state
