"""Extension-level shims: objects with the exact call surface of the reference's pybind
modules `SpaMat` / `SpaVar` (modules/SparseMatching/src/SM_cuda.cpp:7-35,
modules/SparseVar/src/SV_cuda.cpp:7-38), so the reference's OWN functions/SpaMat.py and
functions/SpaVar.py can run unmodified on top of libdecnet_b200.so:

    sys.modules["modules.SparseMatching.build.lib"].SpaMat = decnet_b200.ext.SpaMat

Contract kept: caller allocates (zero-filled) outputs, callee writes in place and returns 1.
Added: dtype/device/contiguity checks and a raised DecnetError instead of silent failure.
"""
from __future__ import annotations

from . import ops


class _SpaMatExt:
    __name__ = "SpaMat"

    @staticmethod
    def sparse_matching_cuda_forward(ref_feas, tar_feas, ref_mask, tar_mask, output,
                                     sum_similarities, max_cost, max_disp):
        ops.spamat_forward(ref_feas, tar_feas, ref_mask, tar_mask, int(max_disp),
                           output=output, sum_sim=sum_similarities, max_cost=max_cost)
        return 1

    @staticmethod
    def sparse_matching_cuda_backward(ref_feas, tar_feas, ref_mask, tar_mask, output, sum_similarities,
                                      max_cost, grad_output, grad_ref_feas, grad_tar_feas, max_disp):
        ops.spamat_backward(ref_feas, tar_feas, ref_mask, tar_mask, output, sum_similarities, max_cost,
                            grad_output, int(max_disp), grad_ref=grad_ref_feas, grad_tar=grad_tar_feas)
        return 1


class _SpaVarExt:
    __name__ = "SpaVar"

    @staticmethod
    def sparse_var_cuda_forward(ref_feas, tar_feas, ref_mask, tar_mask, disparity, output,
                                sum_similarities, max_cost, max_disp):
        ops.spavar_forward(ref_feas, tar_feas, ref_mask, tar_mask, disparity.contiguous(), int(max_disp),
                           output=output, sum_sim=sum_similarities, max_cost=max_cost)
        return 1

    @staticmethod
    def sparse_var_cuda_backward(ref_feas, tar_feas, ref_mask, tar_mask, disparity, output,
                                 sum_similarities, max_cost, grad_output, grad_ref_feas, grad_tar_feas,
                                 grad_disparity, max_disp):
        ops.spavar_backward(ref_feas, tar_feas, ref_mask, tar_mask, disparity.contiguous(), output,
                            sum_similarities, max_cost, grad_output, int(max_disp),
                            grad_ref=grad_ref_feas, grad_tar=grad_tar_feas, grad_disp=grad_disparity)
        return 1


SpaMat = _SpaMatExt()
SpaVar = _SpaVarExt()
