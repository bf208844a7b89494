"""oracle/sparse.py -- TEST INFRASTRUCTURE: Python access to the C restatement
(oracle/sparse_oracle.c) of SpaMat / SpaVar, plus an independent vectorised torch
restatement used to cross-check the C code on CPU.

Reference followed: modules/SparseMatching/src/SM_kernel.cu:22-125,143-195,300-355;
modules/SparseVar/src/SV_kernel.cu:76-124,142-325; zero-fill contract
functions/SpaMat.py:25-27.  Parity pinning: see the header of sparse_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import torch

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "liboracle_sparse.so"
_lib = None


def build() -> Path:
    subprocess.run(["make", "-C", str(_HERE), "oracle"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        _lib = C.CDLL(str(_SO))
    return _lib


def _p(t: torch.Tensor):
    assert t.device.type == "cpu" and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def _f32(t):
    return t.detach().to("cpu", torch.float32).contiguous()


def spamat_forward(L, R, ml, mr, D):
    L, R, ml, mr = map(_f32, (L, R, ml, mr))
    B, Cc, H, W = L.shape
    out, ssim, mx = (torch.zeros(B, H, W) for _ in range(3))
    lib().oracle_spamat_forward(_p(L), _p(R), _p(ml), _p(mr), _p(out), _p(ssim), _p(mx),
                                B, Cc, H, W, int(D))
    return out, ssim, mx


def spavar_forward(L, R, ml, mr, disp, D):
    L, R, ml, mr, disp = map(_f32, (L, R, ml, mr, disp))
    B, Cc, H, W = L.shape
    var, ssim, mx = (torch.zeros(B, H, W) for _ in range(3))
    lib().oracle_spavar_forward(_p(L), _p(R), _p(ml), _p(mr), _p(disp), _p(var), _p(ssim), _p(mx),
                                B, Cc, H, W, int(D))
    return var, ssim, mx


def spamat_backward(L, R, ml, mr, out, ssim, mx, g, D):
    L, R, ml, mr, out, ssim, mx, g = map(_f32, (L, R, ml, mr, out, ssim, mx, g))
    B, Cc, H, W = L.shape
    dL, dR = torch.zeros_like(L), torch.zeros_like(R)
    lib().oracle_spamat_backward(_p(L), _p(R), _p(ml), _p(mr), _p(out), _p(ssim), _p(mx), _p(g),
                                 _p(dL), _p(dR), B, Cc, H, W, int(D))
    return dL, dR


def spavar_backward(L, R, ml, mr, disp, var, ssim, mx, g, D):
    L, R, ml, mr, disp, var, ssim, mx, g = map(_f32, (L, R, ml, mr, disp, var, ssim, mx, g))
    B, Cc, H, W = L.shape
    dL, dR, dd = torch.zeros_like(L), torch.zeros_like(R), torch.zeros_like(disp)
    lib().oracle_spavar_backward(_p(L), _p(R), _p(ml), _p(mr), _p(disp), _p(var), _p(ssim), _p(mx),
                                 _p(g), _p(dL), _p(dR), _p(dd), B, Cc, H, W, int(D))
    return dL, dR, dd


def candidate_signature(ml, mr, D):
    ml, mr = map(_f32, (ml, mr))
    B, H, W = ml.shape
    count = torch.zeros(B, H, W, dtype=torch.int32)
    hsh = torch.zeros(B, H, W, dtype=torch.int64)
    lib().oracle_candidate_signature(_p(ml), _p(mr), C.c_void_p(count.data_ptr()),
                                     C.c_void_p(hsh.data_ptr()), B, H, W, int(D))
    return count, hsh


# --------------------------------------------------------------------------------------
# Independent vectorised torch restatement (second opinion for the C code; follows the
# closed form in SURVEY.md appendix A, summation order differs -> compare at 1e-5).
# --------------------------------------------------------------------------------------
def torch_forward(L, R, ml, mr, D, disp=None):
    """Returns dict(out, var (around disp or out), sum_sim, max_cost) on L's device."""
    B, Cc, H, W = L.shape
    neg = torch.full((B, H, W), float("-inf"), dtype=L.dtype, device=L.device)
    costs = []
    for d in range(min(D, W)):
        c = neg.clone()
        prod = (L[:, :, :, d:] * R[:, :, :, : W - d]).sum(1)
        valid = (ml[:, :, d:] != 0) & (mr[:, :, : W - d] != 0)
        c[:, :, d:] = torch.where(valid, prod, neg[:, :, d:])
        costs.append(c)
    cost = torch.stack(costs, 0) if costs else neg[None][:0]
    if cost.shape[0] == 0:
        mx = torch.full((B, H, W), 1e-6, dtype=L.dtype, device=L.device)
        e = cost
    else:
        mx = cost.max(0).values.clamp(min=1e-6)
        e = torch.exp(cost - mx)
    dvals = torch.arange(cost.shape[0], dtype=L.dtype, device=L.device).view(-1, 1, 1, 1)
    s0 = e.sum(0)
    ssim = 1e-6 + s0
    out = (1e-6 + (e * dvals).sum(0)) / ssim
    mu = out if disp is None else disp
    var = (1e-6 + (e * (dvals - mu) ** 2).sum(0)) / ssim
    m = (ml != 0).to(L.dtype)
    return {"out": out * m, "var": var * m, "sum_sim": ssim * m, "max_cost": mx * m}
