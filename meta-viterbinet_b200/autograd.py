"""'train'-phase priors with autograd (VNETDetector / META_VNETDetector forward(y,'train')).

Forward runs the CUDA priors kernel.  The backward (and the double backward MAML needs,
trainer.py:437 create_graph=True) is provided by train.py's kernels once loaded.
"""
import torch

from . import ops


class _PriorsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, w1, b1, w2, b2, w3, b3):
        ctx.save_for_backward(y, w1, b1, w2, b2, w3, b3)
        with torch.no_grad():
            return ops.vnet_priors(y, [w1, b1, w2, b2, w3, b3])

    @staticmethod
    def backward(ctx, grad_priors):
        from . import train
        return train.priors_backward(ctx.saved_tensors, grad_priors)


def priors_function(y, w1, b1, w2, b2, w3, b3):
    return _PriorsFn.apply(ops.dev_f32(y), w1, b1, w2, b2, w3, b3)
