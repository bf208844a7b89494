"""CPU: the C restatement (oracle/sparse_oracle.c) against the independent torch
restatement, the edge semantics of SURVEY.md section 8a, and its own backward against
autograd through the closed form."""
import pytest
import torch

from oracle import sparse as osp
from helpers import LEVEL_SHAPES, make_feats, make_masks


@pytest.mark.parametrize("name,B,C,H,W,D", LEVEL_SHAPES[:2] + LEVEL_SHAPES[5:])
def test_c_oracle_matches_torch_restatement(name, B, C, H, W, D):
    L, R = make_feats(B, C, H, W)
    ml, mr = make_masks(B, H, W, 0.3, 0.3)
    out, ssim, mx = osp.spamat_forward(L, R, ml, mr, D)
    t = osp.torch_forward(L, R, ml, mr, D)
    assert torch.allclose(out, t["out"], atol=1e-4, rtol=1e-5)
    assert torch.allclose(ssim, t["sum_sim"], atol=1e-5, rtol=1e-5)
    assert torch.allclose(mx, t["max_cost"], atol=1e-6, rtol=1e-6)
    var, ssim2, mx2 = osp.spavar_forward(L, R, ml, mr, out, D)
    assert torch.allclose(var, t["var"], atol=1e-3, rtol=1e-5)
    assert torch.equal(ssim2, ssim) and torch.equal(mx2, mx)


def test_edge_semantics():
    B, C, H, W, D = 1, 4, 2, 16, 6
    L, R = make_feats(B, C, H, W)
    ml = torch.zeros(B, H, W); mr = torch.zeros(B, H, W)
    ml[0, 0, 5] = 1.0                       # masked pixel with NO valid candidate
    ml[0, 1, 7] = 2.5                       # any non-zero counts as masked
    mr[0, 1, 7] = 1.0; mr[0, 1, 3] = -1.0   # d = 0 and d = 4
    mr[0, 1, 0] = 1.0                       # d = 7 is outside D = 6
    out, ssim, mx = osp.spamat_forward(L, R, ml, mr, D)
    assert out[0, 0, 5] == 1.0 and ssim[0, 0, 5] == pytest.approx(1e-6) and mx[0, 0, 5] == pytest.approx(1e-6)
    var, _, _ = osp.spavar_forward(L, R, ml, mr, out, D)
    assert var[0, 0, 5] == 1.0
    # unmasked -> exactly zero everywhere
    keep = ml == 0
    assert out[keep].abs().max() == 0 and ssim[keep].abs().max() == 0 and mx[keep].abs().max() == 0
    cnt, _ = osp.candidate_signature(ml, mr, D)
    assert cnt[0, 1, 7] == 2 and cnt[0, 0, 5] == 0
    # all-negative costs -> max_cost stays at the 1e-6 floor
    Ln = L.abs(); Rn = -R.abs()
    _, _, mxn = osp.spamat_forward(Ln, Rn, ml, mr, D)
    assert mxn[0, 1, 7] == pytest.approx(1e-6)


def test_backward_matches_autograd():
    B, C, H, W, D = 2, 6, 4, 40, 12
    L, R = make_feats(B, C, H, W)
    ml, mr = make_masks(B, H, W, 0.4, 0.4)
    g = torch.randn(B, H, W, generator=torch.Generator().manual_seed(3))
    out, ssim, mx = osp.spamat_forward(L, R, ml, mr, D)
    L2, R2 = L.clone().requires_grad_(), R.clone().requires_grad_()
    (osp.torch_forward(L2, R2, ml, mr, D)["out"] * g).sum().backward()
    dL, dR = osp.spamat_backward(L, R, ml, mr, out, ssim, mx, g, D)
    assert torch.allclose(dL, L2.grad, atol=2e-5, rtol=1e-4)
    assert torch.allclose(dR, R2.grad, atol=2e-5, rtol=1e-4)
    # SpaVar: grads wrt L, R and disp
    disp = (out + 0.37).detach()
    var, ssim_v, mx_v = osp.spavar_forward(L, R, ml, mr, disp, D)
    L3, R3, d3 = L.clone().requires_grad_(), R.clone().requires_grad_(), disp.clone().requires_grad_()
    (osp.torch_forward(L3, R3, ml, mr, D, disp=d3)["var"] * g).sum().backward()
    dLv, dRv, dd = osp.spavar_backward(L, R, ml, mr, disp, var, ssim_v, mx_v, g, D)
    assert torch.allclose(dLv, L3.grad, atol=2e-4, rtol=1e-4)
    assert torch.allclose(dRv, R3.grad, atol=2e-4, rtol=1e-4)
    assert torch.allclose(dd, d3.grad, atol=2e-4, rtol=1e-4)
