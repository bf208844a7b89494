"""Torch-facing wrappers over the C ABI: validate tensors, pass raw device pointers and
the current CUDA stream.  PyTorch is plumbing here (memory + streams), not the product.
"""
from __future__ import annotations

import torch

from . import _lib


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _chk(name: str, t: torch.Tensor, ref: torch.Tensor | None = None, shape=None) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise _lib.DecnetError(f"{name} must be a CUDA tensor (decnet_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if ref is not None and t.device != ref.device:
        raise ValueError(f"{name} is on {t.device}, expected {ref.device}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")


def _feat_args(ref_feas, tar_feas, ref_mask, tar_mask):
    _chk("ref_feas", ref_feas)
    if ref_feas.dim() != 4:
        raise ValueError("ref_feas must be [B,C,H,W]")
    B, Cc, H, W = ref_feas.shape
    _chk("tar_feas", tar_feas, ref_feas, (B, Cc, H, W))
    _chk("ref_mask", ref_mask, ref_feas, (B, H, W))
    _chk("tar_mask", tar_mask, ref_feas, (B, H, W))
    return B, Cc, H, W


def spamat_forward(ref_feas, tar_feas, ref_mask, tar_mask, max_disp, output=None, sum_sim=None, max_cost=None):
    """SpaMat forward -> (output, sum_similarities, max_cost), each [B,H,W]."""
    B, Cc, H, W = _feat_args(ref_feas, tar_feas, ref_mask, tar_mask)
    output = torch.empty_like(ref_mask) if output is None else output
    sum_sim = torch.empty_like(ref_mask) if sum_sim is None else sum_sim
    max_cost = torch.empty_like(ref_mask) if max_cost is None else max_cost
    for n, t in (("output", output), ("sum_similarities", sum_sim), ("max_cost", max_cost)):
        _chk(n, t, ref_feas, (B, H, W))
    with torch.cuda.device_of(ref_feas):
        st = _lib.lib().decnet_spamat_fwd(
            ref_feas.data_ptr(), tar_feas.data_ptr(), ref_mask.data_ptr(), tar_mask.data_ptr(),
            output.data_ptr(), sum_sim.data_ptr(), max_cost.data_ptr(),
            B, Cc, H, W, int(max_disp), _stream(ref_feas))
    _lib.check(st, "decnet_spamat_fwd")
    return output, sum_sim, max_cost


def spavar_forward(ref_feas, tar_feas, ref_mask, tar_mask, disparity, max_disp,
                   output=None, sum_sim=None, max_cost=None):
    """SpaVar forward -> (variance, sum_similarities, max_cost)."""
    B, Cc, H, W = _feat_args(ref_feas, tar_feas, ref_mask, tar_mask)
    _chk("disparity", disparity, ref_feas, (B, H, W))
    output = torch.empty_like(ref_mask) if output is None else output
    sum_sim = torch.empty_like(ref_mask) if sum_sim is None else sum_sim
    max_cost = torch.empty_like(ref_mask) if max_cost is None else max_cost
    for n, t in (("output", output), ("sum_similarities", sum_sim), ("max_cost", max_cost)):
        _chk(n, t, ref_feas, (B, H, W))
    with torch.cuda.device_of(ref_feas):
        st = _lib.lib().decnet_spavar_fwd(
            ref_feas.data_ptr(), tar_feas.data_ptr(), ref_mask.data_ptr(), tar_mask.data_ptr(),
            disparity.data_ptr(), output.data_ptr(), sum_sim.data_ptr(), max_cost.data_ptr(),
            B, Cc, H, W, int(max_disp), _stream(ref_feas))
    _lib.check(st, "decnet_spavar_fwd")
    return output, sum_sim, max_cost


def spamat_spavar_forward(ref_feas, tar_feas, ref_mask, tar_mask, max_disp):
    """Fused SpaMat + SpaVar(disparity = SpaMat output) -> (disp, var, sum_sim, max_cost)."""
    B, Cc, H, W = _feat_args(ref_feas, tar_feas, ref_mask, tar_mask)
    outs = [torch.empty_like(ref_mask) for _ in range(4)]
    with torch.cuda.device_of(ref_feas):
        st = _lib.lib().decnet_spamat_spavar_fwd(
            ref_feas.data_ptr(), tar_feas.data_ptr(), ref_mask.data_ptr(), tar_mask.data_ptr(),
            outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(),
            B, Cc, H, W, int(max_disp), _stream(ref_feas))
    _lib.check(st, "decnet_spamat_spavar_fwd")
    return tuple(outs)


def spamat_backward(ref_feas, tar_feas, ref_mask, tar_mask, output, sum_sim, max_cost, grad_output,
                    max_disp, grad_ref=None, grad_tar=None):
    B, Cc, H, W = _feat_args(ref_feas, tar_feas, ref_mask, tar_mask)
    for n, t in (("output", output), ("sum_similarities", sum_sim), ("max_cost", max_cost),
                 ("grad_output", grad_output)):
        _chk(n, t, ref_feas, (B, H, W))
    grad_ref = torch.zeros_like(ref_feas) if grad_ref is None else grad_ref
    grad_tar = torch.zeros_like(tar_feas) if grad_tar is None else grad_tar
    _chk("grad_ref_feas", grad_ref, ref_feas, (B, Cc, H, W))
    _chk("grad_tar_feas", grad_tar, ref_feas, (B, Cc, H, W))
    with torch.cuda.device_of(ref_feas):
        st = _lib.lib().decnet_spamat_bwd(
            ref_feas.data_ptr(), tar_feas.data_ptr(), ref_mask.data_ptr(), tar_mask.data_ptr(),
            output.data_ptr(), sum_sim.data_ptr(), max_cost.data_ptr(), grad_output.data_ptr(),
            grad_ref.data_ptr(), grad_tar.data_ptr(), B, Cc, H, W, int(max_disp), _stream(ref_feas))
    _lib.check(st, "decnet_spamat_bwd")
    return grad_ref, grad_tar


def spavar_backward(ref_feas, tar_feas, ref_mask, tar_mask, disparity, output, sum_sim, max_cost,
                    grad_output, max_disp, grad_ref=None, grad_tar=None, grad_disp=None):
    B, Cc, H, W = _feat_args(ref_feas, tar_feas, ref_mask, tar_mask)
    for n, t in (("disparity", disparity), ("output", output), ("sum_similarities", sum_sim),
                 ("max_cost", max_cost), ("grad_output", grad_output)):
        _chk(n, t, ref_feas, (B, H, W))
    grad_ref = torch.zeros_like(ref_feas) if grad_ref is None else grad_ref
    grad_tar = torch.zeros_like(tar_feas) if grad_tar is None else grad_tar
    grad_disp = torch.zeros_like(disparity) if grad_disp is None else grad_disp
    _chk("grad_ref_feas", grad_ref, ref_feas, (B, Cc, H, W))
    _chk("grad_tar_feas", grad_tar, ref_feas, (B, Cc, H, W))
    _chk("grad_disparity", grad_disp, ref_feas, (B, H, W))
    with torch.cuda.device_of(ref_feas):
        st = _lib.lib().decnet_spavar_bwd(
            ref_feas.data_ptr(), tar_feas.data_ptr(), ref_mask.data_ptr(), tar_mask.data_ptr(),
            disparity.data_ptr(), output.data_ptr(), sum_sim.data_ptr(), max_cost.data_ptr(),
            grad_output.data_ptr(), grad_ref.data_ptr(), grad_tar.data_ptr(), grad_disp.data_ptr(),
            B, Cc, H, W, int(max_disp), _stream(ref_feas))
    _lib.check(st, "decnet_spavar_bwd")
    return grad_ref, grad_tar, grad_disp


def candidate_signature(ref_mask, tar_mask, max_disp):
    """(count int32 [B,H,W], hash int64 [B,H,W]) of every pixel's candidate set."""
    _chk("ref_mask", ref_mask)
    _chk("tar_mask", tar_mask, ref_mask, ref_mask.shape)
    B, H, W = ref_mask.shape
    count = torch.empty((B, H, W), dtype=torch.int32, device=ref_mask.device)
    hsh = torch.empty((B, H, W), dtype=torch.int64, device=ref_mask.device)
    with torch.cuda.device_of(ref_mask):
        st = _lib.lib().decnet_candidate_signature(
            ref_mask.data_ptr(), tar_mask.data_ptr(), count.data_ptr(), hsh.data_ptr(),
            B, H, W, int(max_disp), _stream(ref_mask))
    _lib.check(st, "decnet_candidate_signature")
    return count, hsh
