"""SupConLoss with the reference's signature (utils/loss.py:14-96), computed by the fused
CUDA kernel pair of csrc/supcon.cu through the C ABI."""
import torch
import torch.nn as nn

from . import ops


class _SupConFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, labels, temperature):
        loss, grad = ops.supcon(features, labels, temperature, need_grad=features.requires_grad)
        ctx.save_for_backward(grad if grad is not None else torch.empty(0, device=features.device))
        ctx.shape = features.shape
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return (grad * grad_out).reshape(ctx.shape), None, None


class SupConLoss(nn.Module):
    """Supervised contrastive loss, all-views-anchor mode, labels given (the only mode the
    replay path uses: agents/base.py:109-111).  Same constructor and call signature as the
    reference; `mask=` and contrast_mode='one' are outside the replay path and raise."""

    def __init__(self, temperature=0.07, contrast_mode='all'):
        super().__init__()
        self.temperature = temperature
        self.contrast_mode = contrast_mode

    def forward(self, features, labels=None, mask=None):
        if len(features.shape) < 3:
            raise ValueError('`features` needs to be [bsz, n_views, ...],'
                             'at least 3 dimensions are required')
        if labels is not None and mask is not None:
            raise ValueError('Cannot define both `labels` and `mask`')
        if self.contrast_mode != 'all':
            raise ValueError('Unknown mode: {}'.format(self.contrast_mode)) if self.contrast_mode != 'one' else \
                NotImplementedError("contrast_mode='one' is not on the replay path")
        if mask is not None:
            raise NotImplementedError('explicit `mask` is not on the replay path (agents/base.py:110 passes labels)')
        if labels is None:
            # SimCLR degenerate case (loss.py:45-46): every sample is its own class
            labels = torch.arange(features.shape[0], device=features.device)
        labels = labels.contiguous().view(-1)
        if labels.shape[0] != features.shape[0]:
            raise ValueError('Num of labels does not match num of features')
        return _SupConFn.apply(features, labels, float(self.temperature))

    def loss_and_grad(self, features, labels):
        """Engine path: (loss[1], dL/dfeatures) in one call, no autograd graph."""
        return ops.supcon(features, labels, float(self.temperature), need_grad=True)
