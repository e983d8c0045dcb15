"""Oracle: GSS-greedy memory update (utils/buffer/gss_greedy_update.py:7-124) on the CPU.

Restates the update rule on top of oracle/resnet.py: every gradient is an EVAL-mode forward (the reference
switches the model to eval() first, gss_greedy_update.py:16) + mean cross-entropy + backward, flattened in
parameters() order (get_grad_vector, buffer_utils.py:58-73); scores are cosine similarities with clamped
denominators (buffer_utils.py:50-55).  The random decisions are the reference's own torch calls
(randperm :98, multinomial :30 and :38) on the default CPU generator, so a seeded run reproduces the
reference's draws.  Test infrastructure only -- see oracle/__init__.py.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import resnet as oresnet


def eval_grad_vector(spec, params, bn, x, y):
    """model.eval(); zero_grad(); F.cross_entropy(model.forward(x), y).backward(); get_grad_vector()
    (gss_greedy_update.py:80-83 with buffer_utils.py:58-73).  Returns the flat fp32 gradient."""
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
    loss = F.cross_entropy(oresnet.forward(spec, leaves, bn, x, train=False), y)
    loss.backward()
    return torch.cat([(v.grad if v.grad is not None else torch.zeros_like(v)).reshape(-1) for v in leaves.values()])


def cosine_scores(mem_grads, g, eps=1e-8):
    """cosine_similarity(mem_grads, g[None]) -> [K] (buffer_utils.py:50-55), and its maximum."""
    w1 = mem_grads.norm(p=2, dim=1, keepdim=True)
    w2 = g.reshape(1, -1).norm(p=2, dim=1, keepdim=True)
    sim = torch.mm(mem_grads, g.reshape(-1, 1)) / (w1 * w2.t()).clamp(min=eps)
    return sim.reshape(-1), sim.max()


class GSSState(object):
    """The pieces of Buffer + GSSGreedyUpdate the rule reads and writes."""

    def __init__(self, spec, params, bn, mem_size, in_shape, mem_strength=10, gss_batch_size=10):
        self.spec, self.params, self.bn = spec, params, bn
        self.buffer_img = torch.zeros((mem_size,) + tuple(in_shape), dtype=torch.float32)
        self.buffer_label = torch.zeros(mem_size, dtype=torch.int64)
        self.buffer_score = torch.zeros(mem_size, dtype=torch.float32)
        self.current_index = 0
        self.mem_strength, self.gss_batch_size = mem_strength, gss_batch_size
        self.last_batch_sim = None


def rand_mem_grads(st):
    """gss_greedy_update.py:87-107."""
    gss_batch_size = min(st.gss_batch_size, st.current_index)
    num_mem_subs = min(st.mem_strength, st.current_index // gss_batch_size)
    shuffled = torch.randperm(st.current_index)
    rows = []
    for i in range(num_mem_subs):
        ind = shuffled[i * gss_batch_size:i * gss_batch_size + gss_batch_size]
        rows.append(eval_grad_vector(st.spec, st.params, st.bn, st.buffer_img[ind], st.buffer_label[ind]))
    return torch.stack(rows)


def each_sample_sim(st, mem_grads, x, y):
    """gss_greedy_update.py:109-124."""
    out = torch.zeros(x.size(0))
    for i in range(x.size(0)):
        g = eval_grad_vector(st.spec, st.params, st.bn, x[i:i + 1], y[i:i + 1])
        out[i] = cosine_scores(mem_grads, g)[1]
    return out


def update(st, x, y):
    """gss_greedy_update.py:15-64.  Returns the list of (batch position, slot) pairs written."""
    written = []
    place_left = st.buffer_img.size(0) - st.current_index
    if place_left <= 0:
        mem_grads = rand_mem_grads(st)
        batch_grad = eval_grad_vector(st.spec, st.params, st.bn, x, y)
        batch_sim = cosine_scores(mem_grads, batch_grad)[1]
        st.last_batch_sim = float(batch_sim)
        if batch_sim < 0:
            score = st.buffer_score[:st.current_index]
            buffer_sim = (score - torch.min(score)) / ((torch.max(score) - torch.min(score)) + 0.01)
            index = torch.multinomial(buffer_sim, x.size(0), replacement=False)
            batch_item_sim = each_sample_sim(st, mem_grads, x, y)
            scaled = ((batch_item_sim + 1) / 2).unsqueeze(1)
            repl = ((st.buffer_score[index] + 1) / 2).unsqueeze(1)
            outcome = torch.multinomial(torch.cat((scaled, repl), dim=1), 1, replacement=False)
            sub = outcome.squeeze(1).bool()
            added = torch.arange(end=batch_item_sim.size(0))
            st.buffer_img[index[sub]] = x[added[sub]].clone()
            st.buffer_label[index[sub]] = y[added[sub]].clone()
            st.buffer_score[index[sub]] = batch_item_sim[added[sub]].clone()
            written = list(zip(added[sub].tolist(), index[sub].tolist()))
    else:
        offset = min(place_left, x.size(0))
        x, y = x[:offset], y[:offset]
        if st.current_index == 0:
            cos = torch.zeros(x.size(0)) + 0.1
        else:
            mem_grads = rand_mem_grads(st)
            cos = each_sample_sim(st, mem_grads, x, y)
        s = st.current_index
        st.buffer_img[s:s + offset] = x
        st.buffer_label[s:s + offset] = y
        st.buffer_score[s:s + offset] = cos
        st.current_index += offset
        written = [(i, s + i) for i in range(offset)]
    return written
