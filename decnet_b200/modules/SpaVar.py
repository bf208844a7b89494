"""SpaVar -- variance of the sparse matching distribution, backed by libdecnet_b200.so.

Drop-in for the reference's SpaVar (modules/SparseVar/modules/SpaVar.py:12-28) and
SpaVarFunction (modules/SparseVar/functions/SpaVar.py:8-52):
`(ref_feas, tar_feas, ref_mask, tar_mask, disparity, max_disp) -> Tensor[B,H,W]`.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.nn import Module

from .. import ops


class SpaVarFunction(Function):
    @staticmethod
    def forward(ctx, ref_feas, tar_feas, ref_mask, tar_mask, disparity, max_disp):
        # reference: functions/SpaVar.py:21-22 (disparity contiguity is NOT asserted there;
        # the raw-pointer ABI needs it, so it is made contiguous here)
        assert ref_feas.is_contiguous() and tar_feas.is_contiguous()
        assert ref_mask.is_contiguous() and tar_mask.is_contiguous()
        max_disp = int(max_disp)
        disparity = disparity.contiguous()
        output, sum_similarities, max_cost = ops.spavar_forward(ref_feas, tar_feas, ref_mask, tar_mask,
                                                                disparity, max_disp)
        ctx.save_for_backward(ref_feas, tar_feas, ref_mask, tar_mask, disparity, output,
                              sum_similarities, max_cost)
        ctx.max_disp = max_disp
        return output

    @staticmethod
    def backward(ctx, grad_output):
        (ref_feas, tar_feas, ref_mask, tar_mask, disparity, output,
         sum_similarities, max_cost) = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        grad_ref, grad_tar, grad_disp = ops.spavar_backward(
            ref_feas, tar_feas, ref_mask, tar_mask, disparity, output, sum_similarities, max_cost,
            grad_output, ctx.max_disp)
        return grad_ref, grad_tar, None, None, grad_disp, None


class SpaVar(Module):
    def __init__(self):
        super().__init__()

    def forward(self, ref_feas, tar_feas, ref_mask, tar_mask, disparity, max_disp):
        return SpaVarFunction.apply(ref_feas, tar_feas, ref_mask, tar_mask, disparity, max_disp)
