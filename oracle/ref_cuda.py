"""oracle/ref_cuda.py -- TEST INFRASTRUCTURE: ctypes access to the UNMODIFIED reference
CUDA kernels compiled into oracle/_ref/ by oracle/Makefile (`make ref`).  GPU only; used
by the -m gpu tests to pin the CPU restatement and our kernels against the reference
itself, and by bench.py to time the reference kernels beside ours.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
_libs = {}


def available() -> bool:
    return (_HERE / "_ref" / "libref_spamat.so").exists() and (_HERE / "_ref" / "libref_spavar.so").exists()


def _lib(name):
    if name not in _libs:
        _libs[name] = C.CDLL(str(_HERE / "_ref" / f"libref_{name}.so"))
    return _libs[name]


def _p(t):
    assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32
    return C.c_void_p(t.data_ptr())


def _sync_check(rc, what):
    torch.cuda.synchronize()
    if rc != 0:
        raise RuntimeError(f"{what}: cudaError {rc}")


def spamat_forward(L, R, ml, mr, D, sync=True):
    """Reference kernels launch on the legacy default stream; callers must not rely on
    torch's current stream ordering unless sync=True."""
    B, Cc, H, W = L.shape
    out, ssim, mx = (torch.zeros(B, H, W, device=L.device) for _ in range(3))
    if sync:
        torch.cuda.synchronize()
    rc = _lib("spamat").ref_spamat_forward(_p(L), _p(R), _p(ml), _p(mr), _p(out), _p(ssim), _p(mx),
                                           B, Cc, H, W, int(D))
    if sync:
        _sync_check(rc, "ref_spamat_forward")
    return out, ssim, mx


def spavar_forward(L, R, ml, mr, disp, D, sync=True):
    B, Cc, H, W = L.shape
    var, ssim, mx = (torch.zeros(B, H, W, device=L.device) for _ in range(3))
    if sync:
        torch.cuda.synchronize()
    rc = _lib("spavar").ref_spavar_forward(_p(L), _p(R), _p(ml), _p(mr), _p(disp), _p(var), _p(ssim),
                                           _p(mx), B, Cc, H, W, int(D))
    if sync:
        _sync_check(rc, "ref_spavar_forward")
    return var, ssim, mx


def spamat_backward(L, R, ml, mr, out, ssim, mx, g, D):
    B, Cc, H, W = L.shape
    dL, dR = torch.zeros_like(L), torch.zeros_like(R)
    torch.cuda.synchronize()
    rc = _lib("spamat").ref_spamat_backward(_p(L), _p(R), _p(ml), _p(mr), _p(out), _p(ssim), _p(mx),
                                            _p(g), _p(dL), _p(dR), B, Cc, H, W, int(D))
    _sync_check(rc, "ref_spamat_backward")
    return dL, dR


def spavar_backward(L, R, ml, mr, disp, var, ssim, mx, g, D):
    B, Cc, H, W = L.shape
    dL, dR, dd = torch.zeros_like(L), torch.zeros_like(R), torch.zeros_like(disp)
    torch.cuda.synchronize()
    rc = _lib("spavar").ref_spavar_backward(_p(L), _p(R), _p(ml), _p(mr), _p(disp), _p(var), _p(ssim),
                                            _p(mx), _p(g), _p(dL), _p(dR), _p(dd), B, Cc, H, W, int(D))
    _sync_check(rc, "ref_spavar_backward")
    return dL, dR, dd
