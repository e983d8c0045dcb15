"""compute_knn_sv with the reference signature (utils/buffer/aser_utils.py:7-61): deep features
from the CUDA engine (eval mode, like mini_batch_deep_features, utils/utils.py:45-90) and the
fused kNN-SV kernel.  Returns the [n_eval, n_cand] Shapley matrix the reference returns."""
import torch

from . import ops
from .memory import input_size_match
from .nets import EngineModel, adopt, engine_of


def deep_features(model, eval_x, n_eval, cand_x, n_cand):
    """aser_utils.py:64-91: one eval-mode pass over eval and candidate images."""
    if isinstance(model, EngineModel):
        eng = model.engine
    else:
        try:
            eng = engine_of(model)
        except RuntimeError:
            eng = adopt(model, eval_x.shape[-1])
    total = eval_x if cand_x is None else torch.cat((eval_x, cand_x), 0)
    feats = eng.features_eval(total)
    return feats[:n_eval], feats[n_eval:]


def compute_knn_sv(model, eval_x, eval_y, cand_x, cand_y, k, device='cpu'):
    n_eval, n_cand = eval_x.size(0), cand_x.size(0)
    if eval_x.dim() == 2:                     # already feature matrices
        eval_f, cand_f = eval_x, cand_x
    else:
        eval_f, cand_f = deep_features(model, eval_x, n_eval, cand_x, n_cand)
    return ops.knn_sv(eval_f, eval_y, cand_f, cand_y, k, want_matrix=True, want_sum=False)['sv']
