"""'train'-phase priors with autograd (VNETDetector / META_VNETDetector forward(y,'train')).

Forward runs the CUDA priors kernel; backward and double backward run the kernels of
csrc/train_kernels.cu, so both the first-order loops (trainer.py:492-505) and the second-order
MAML step (trainer.py:437 create_graph=True, :441-449) work through torch.autograd unchanged.
"""
import torch
from torch.autograd.function import once_differentiable

from . import ops


class _PriorsBwdFn(torch.autograd.Function):
    """(y, grad_priors, weights) -> grads of the six weights; differentiable once more."""

    @staticmethod
    def forward(ctx, y, grad_priors, *weights):
        from . import train
        ctx.save_for_backward(y, grad_priors, *weights)
        return train.priors_backward(y, weights, grad_priors)

    @staticmethod
    @once_differentiable
    def backward(ctx, *upstream):
        from . import train
        y, grad_priors, *weights = ctx.saved_tensors
        ggp, gw = train.priors_backward2(y, weights, grad_priors, upstream)
        return (None, ggp) + gw


class _PriorsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, *weights):
        ctx.save_for_backward(y, *weights)
        return ops.vnet_priors(y, list(weights))

    @staticmethod
    def backward(ctx, grad_priors):
        y, *weights = ctx.saved_tensors
        return (None,) + _PriorsBwdFn.apply(y, grad_priors.contiguous(), *weights)


def priors_function(y, w1, b1, w2, b2, w3, b3):
    return _PriorsFn.apply(ops.dev_f32(y), w1, b1, w2, b2, w3, b3)
