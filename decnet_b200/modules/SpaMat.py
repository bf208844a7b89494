"""SpaMat -- sparse matching module + autograd function backed by libdecnet_b200.so.

Drop-in for the reference's SpaMat (modules/SparseMatching/modules/SpaMat.py:12-28) and
SpaMatFunction (modules/SparseMatching/functions/SpaMat.py:8-50): same constructor (no
arguments, no parameters/buffers, so state_dicts are unaffected), same forward signature
`(ref_feas, tar_feas, ref_mask, tar_mask, max_disp) -> Tensor[B,H,W]`, same preconditions
(contiguous inputs -> AssertionError otherwise), autograd-capable.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.nn import Module

from .. import ops


class SpaMatFunction(Function):
    @staticmethod
    def forward(ctx, ref_feas, tar_feas, ref_mask, tar_mask, max_disp):
        # reference: functions/SpaMat.py:21-22
        assert ref_feas.is_contiguous() and tar_feas.is_contiguous()
        assert ref_mask.is_contiguous() and tar_mask.is_contiguous()
        max_disp = int(max_disp)  # arrives as numpy.int64 (SparseDenseNetRefinementMask.py:124)
        output, sum_similarities, max_cost = ops.spamat_forward(ref_feas, tar_feas, ref_mask, tar_mask, max_disp)
        ctx.save_for_backward(ref_feas, tar_feas, ref_mask, tar_mask, output, sum_similarities, max_cost)
        ctx.max_disp = max_disp
        return output

    @staticmethod
    def backward(ctx, grad_output):
        ref_feas, tar_feas, ref_mask, tar_mask, output, sum_similarities, max_cost = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        grad_ref, grad_tar = ops.spamat_backward(ref_feas, tar_feas, ref_mask, tar_mask, output,
                                                 sum_similarities, max_cost, grad_output, ctx.max_disp)
        # the reference returns dummy Tensor([0]) for the masks (functions/SpaMat.py:50); masks are
        # produced under no_grad, so None is the equivalent autograd answer.
        return grad_ref, grad_tar, None, None, None


class SpaMat(Module):
    def __init__(self):
        super().__init__()

    def forward(self, ref_feas, tar_feas, ref_mask, tar_mask, max_disp):
        """ref_feas/tar_feas: [B,C,H,W]; ref_mask/tar_mask: [B,H,W]; returns disparity [B,H,W]."""
        return SpaMatFunction.apply(ref_feas, tar_feas, ref_mask, tar_mask, max_disp)
