"""Oracle: exact kNN Shapley values of candidates w.r.t. evaluation points.

Restates ``compute_knn_sv`` (reference utils/buffer/aser_utils.py:7-61) on
*feature matrices* (the network forward of aser_utils.py:27 is covered by
oracle/resnet.py).  Test infrastructure only -- see oracle/__init__.py.

Conventions shared with the CUDA kernel (csrc/knn_sv.cu):
  * distance  = squared L2, direct-difference form sum((u-v)^2)
                (utils/utils.py:93-95), never |u|^2+|v|^2-2uv;
  * ordering  = ascending distance, ties broken lowest candidate index first
                (the reference's argsort at aser_utils.py:115 is unstable, so the
                tie order is ours to define -- SURVEY.md section 7.3-2);
  * recurrence (aser_utils.py:38-52), ranks i = 1..N over the sorted list,
                m_i = 1[label(cand_i) == label(eval)]:
                    s_N = m_N / N
                    s_i = s_{i+1} + (m_i - m_{i+1}) * min(i, k) / (i * k)
"""
import numpy as np


def sq_dist_matrix(eval_f, cand_f, dtype=np.float64, chunk=256):
    """[E,d],[C,d] -> [E,C] squared L2, direct-difference (utils/utils.py:93-95)."""
    eval_f = np.asarray(eval_f, dtype=dtype)
    cand_f = np.asarray(cand_f, dtype=dtype)
    E, C = eval_f.shape[0], cand_f.shape[0]
    out = np.empty((E, C), dtype=dtype)
    for s in range(0, E, chunk):
        diff = eval_f[s:s + chunk, None, :] - cand_f[None, :, :]
        out[s:s + chunk] = np.einsum('ecd,ecd->ec', diff, diff)
    return out


def sorted_cand_ind(eval_f, cand_f, dtype=np.float64):
    """Candidate indices by ascending distance per eval row, stable
    (aser_utils.py:94-116)."""
    dist = sq_dist_matrix(eval_f, cand_f, dtype=dtype)
    return np.argsort(dist, axis=1, kind='stable'), dist


def sv_factor(n_cand, k, dtype=np.float64):
    """factor[j] for 0-based sorted position j (rank i=j+1), aser_utils.py:43-49:
    min(i,k)/(i*k) for i < N, and 1/N for i == N."""
    i = np.arange(1, n_cand + 1, dtype=dtype)
    numer = i.copy()
    denom = i.copy()
    denom[:n_cand - 1] *= k
    numer[k:n_cand - 1] = k
    numer[n_cand - 1] = 1
    return numer / denom


def knn_sv_from_sorted(sorted_ind, eval_y, cand_y, k, dtype=np.float64):
    """SV matrix [E,C] in *candidate order* from per-row sorted candidate indices
    (aser_utils.py:32-59)."""
    sorted_ind = np.asarray(sorted_ind)
    eval_y = np.asarray(eval_y)
    cand_y = np.asarray(cand_y)
    E, C = sorted_ind.shape
    match = (cand_y[sorted_ind] == eval_y[:, None]).astype(dtype)          # :35-38
    nxt = np.zeros_like(match)
    nxt[:, :C - 1] = match[:, 1:]                                          # :39-40
    term = (match - nxt) * sv_factor(C, k, dtype)[None, :]                 # :41-51
    sv_sorted = np.cumsum(term[:, ::-1], axis=1, dtype=dtype)[:, ::-1]     # :52
    sv = np.zeros((E, C), dtype=dtype)
    np.put_along_axis(sv, sorted_ind, sv_sorted, axis=1)                   # :55-59
    return sv


def knn_sv_matrix(eval_f, eval_y, cand_f, cand_y, k, dtype=np.float64):
    """Full restatement on features.  Returns (sv [E,C], sorted_ind [E,C], dist [E,C])."""
    sorted_ind, dist = sorted_cand_ind(eval_f, cand_f, dtype=dtype)
    return knn_sv_from_sorted(sorted_ind, eval_y, cand_y, k, dtype=dtype), sorted_ind, dist


def knn_sv_row_loop(dist_row, eval_label, cand_y, k):
    """Pure-python per-row recurrence (Jia et al. 2019), fp64 -- small cases only.
    Independent of the vectorised form above; used to cross-check it."""
    C = len(dist_row)
    order = sorted(range(C), key=lambda j: (dist_row[j], j))
    m = [1.0 if cand_y[j] == eval_label else 0.0 for j in order]
    s = [0.0] * C
    s[C - 1] = m[C - 1] / C
    for pos in range(C - 2, -1, -1):
        i = pos + 1
        s[pos] = s[pos + 1] + (m[pos] - m[pos + 1]) * min(i, k) / (i * k)
    out = [0.0] * C
    for pos, j in enumerate(order):
        out[j] = s[pos]
    return out


def column_reductions(sv):
    """The three row-reductions the ASER plugins consume
    (aser_retrieve.py:79,82,86; aser_update.py:80)."""
    return sv.sum(0), sv.max(0), sv.min(0)


def order_mismatch_explained(sorted_a, sorted_b, dist64, rel_tol=4e-6):
    """Compare two per-row orderings.  A difference is *explained* when every
    position where they differ lies inside a run of candidates whose fp64
    distances agree to rel_tol (an fp32 summation-order near-tie).  Returns
    (n_rows_different, n_rows_unexplained)."""
    n_diff = n_bad = 0
    for r in range(sorted_a.shape[0]):
        a, b = sorted_a[r], sorted_b[r]
        if np.array_equal(a, b):
            continue
        n_diff += 1
        da, db = dist64[r][a], dist64[r][b]
        scale = np.maximum(np.abs(da), np.abs(db)) + 1e-30
        pos = np.nonzero(a != b)[0]
        if np.any(np.abs(da[pos] - db[pos]) > rel_tol * scale[pos]):
            n_bad += 1
    return n_diff, n_bad
