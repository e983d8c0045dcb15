"""Oracle: ASER retrieve / update index selection, minority-class rule, reservoir slots.

Restates the host-side decision logic of the reference plugins on *Shapley
value matrices* (the matrices themselves come from oracle/knn_sv.py).  Test
infrastructure only -- see oracle/__init__.py.

Tie policy: the reference ranks with ``argsort(descending=True)`` (unstable,
aser_retrieve.py:88, aser_update.py:88); here equal scores rank lowest index
first, which is the policy of csrc/topk.cu.
"""
import numpy as np


def argsort_desc_stable(v):
    v = np.asarray(v)
    return np.argsort(-v, kind='stable')


def retrieve_score(sv_adv, sv_coop, aser_type):
    """Per-candidate score (aser_retrieve.py:77-86).
    asv: coop.max(0) - adv.min(0); asvm (and anything else): coop.mean(0) - adv.mean(0);
    neg_sv: -adv.sum(0) (no cooperative term)."""
    if aser_type == 'neg_sv':
        return -sv_adv.sum(0)
    if aser_type == 'asv':
        return sv_coop.max(0) - sv_adv.min(0)
    return sv_coop.mean(0) - sv_adv.mean(0)


def retrieve_indices(sv_adv, sv_coop, aser_type, num_retrieve):
    """Positions (into the candidate set) of the retrieved samples, in rank order
    (aser_retrieve.py:88-91)."""
    return argsort_desc_stable(retrieve_score(sv_adv, sv_coop, aser_type))[:num_retrieve]


def update_partition(sv_sum, n_cand_buf, cand_ind):
    """Replacement sets of ASER update (aser_update.py:82-102).

    sv_sum     [n_cand]  column sums of the SV matrix, candidates = n_cand_buf buffered
                         samples followed by the current batch;
    cand_ind   [n_cand_buf]  buffer slot of each buffered candidate.
    Returns (ind_cur, ind_buffer): current-batch positions to insert (in rank
    order) and the buffer slots they overwrite (in rank order)."""
    order = argsort_desc_stable(sv_sum)
    large, small = order[:n_cand_buf], order[n_cand_buf:]
    ind_cur = large[large >= n_cand_buf] - n_cand_buf
    ind_buffer = np.asarray(cand_ind)[small[small < n_cand_buf]]
    return ind_cur, ind_buffer


def minority_positions(cur_y, class_num_cache, mem_size, threshold):
    """Positions of current-batch samples whose class share of the memory is below
    the threshold (aser_utils.py:148-157; threshold ~ U(0, 1/num_class) drawn by the
    caller)."""
    share = np.asarray(class_num_cache, dtype=np.float32) / np.float32(mem_size)
    return np.nonzero(share[np.asarray(cur_y)] < np.float32(threshold))[0]


def reservoir_slots(draws, mem_size):
    """Reservoir overwrite map (reservoir_update.py:35-60): draws[i] is the integer
    slot drawn for new sample i (already floor()ed); kept when < mem_size; a slot
    drawn twice keeps the LAST sample, and slots are reported in first-seen order
    (python dict semantics, reservoir_update.py:53)."""
    idx_map = {}
    for i, s in enumerate(np.asarray(draws).tolist()):
        if s < mem_size:
            idx_map[int(s)] = i
    return list(idx_map.keys()), list(idx_map.values())


def rank_equivalent(pos_a, pos_b, score, atol=0.0):
    """True when two top-N selections differ only inside groups of candidates whose
    scores are equal to within atol (the reference's unstable argsort leaves the
    order inside such groups implementation-defined)."""
    pos_a, pos_b, score = np.asarray(pos_a), np.asarray(pos_b), np.asarray(score, dtype=np.float64)
    if pos_a.shape != pos_b.shape:
        return False
    if pos_a.size == 0:
        return True
    if np.any(np.abs(score[pos_a] - score[pos_b]) > atol):
        return False
    cutoff = score[pos_a[-1]]
    sure_a = set(pos_a[score[pos_a] > cutoff + atol].tolist())
    sure_b = set(pos_b[score[pos_b] > cutoff + atol].tolist())
    return sure_a == sure_b


def min_adjacent_gap(score, n_top):
    """Smallest gap between consecutive sorted scores among the first n_top+1 ranks
    (how close a selection is to a tie)."""
    s = np.sort(np.asarray(score, dtype=np.float64))[::-1][:n_top + 1]
    return np.inf if s.size < 2 else float(np.min(s[:-1] - s[1:]))
