"""Oracle: supervised-contrastive loss and its gradient, all-views-anchor mode.

Restates ``SupConLoss.forward`` (reference utils/loss.py:19-96) in numpy, with
the analytic gradient the fused CUDA kernel produces (csrc/supcon.cu).  Test
infrastructure only -- see oracle/__init__.py.

With A = V*B anchors in *view-major* order (loss.py:56: cat(unbind(features,1))):
    l_ij   = c_i . c_j / T                                   (loss.py:67-69)
    mx_i   = max_j l_ij   (diagonal included, detached)      (loss.py:71-72)
    Z_i    = sum_{j != i} exp(l_ij - mx_i)                   (loss.py:77-86)
    P(i)   = { j != i : label_j == label_i }                 (loss.py:51,75,83)
    loss   = -(1/A) sum_i [ sum_{j in P(i)} (l_ij - mx_i - log Z_i) ] / |P(i)|   (loss.py:87-94)
No base_temperature factor (loss.py:93 differs from upstream SupContrast).
Gradient:  G_ij = (exp(l_ij - mx_i)/Z_i - 1[j in P(i)]/|P(i)|) / A  for j != i, 0 on the diagonal;
           dL/dc = (G + G^T) c / T.
"""
import numpy as np


def supcon_loss_and_grad(features, labels, temperature=0.07, dtype=np.float64):
    """features [B,V,...] float, labels [B] int -> (loss scalar, dfeatures [B,V,d])."""
    f = np.asarray(features, dtype=dtype)
    if f.ndim < 3:
        raise ValueError('`features` needs to be [bsz, n_views, ...]')   # loss.py:36-38
    B, V = f.shape[0], f.shape[1]
    f = f.reshape(B, V, -1)
    labels = np.asarray(labels).reshape(-1)
    if labels.shape[0] != B:
        raise ValueError('Num of labels does not match num of features')  # loss.py:49-50
    d = f.shape[2]
    A = V * B
    c = f.transpose(1, 0, 2).reshape(A, d)            # view-major, loss.py:56
    lab = np.tile(labels, V)
    logits = (c @ c.T) / dtype(temperature)
    mx = logits.max(axis=1, keepdims=True)
    sh = logits - mx
    off = 1.0 - np.eye(A, dtype=dtype)
    pos = (lab[:, None] == lab[None, :]).astype(dtype) * off
    e = np.exp(sh) * off
    Z = e.sum(1, keepdims=True)
    log_prob = sh - np.log(Z)
    npos = pos.sum(1)
    with np.errstate(invalid='ignore', divide='ignore'):
        mean_lp = (pos * log_prob).sum(1) / npos       # 0/0 -> nan, as loss.py:90
        loss = -mean_lp.mean()
        G = (e / Z - pos / npos[:, None]) / A
    G = G * off
    dc = ((G + G.T) @ c) / dtype(temperature)
    dfeat = dc.reshape(V, B, d).transpose(1, 0, 2)
    return dtype(loss), dfeat


def supcon_loss_only(features, labels, temperature=0.07, dtype=np.float64):
    return supcon_loss_and_grad(features, labels, temperature, dtype)[0]
