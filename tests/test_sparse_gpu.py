"""GPU parity: our CUDA SpaMat/SpaVar (through the C ABI) against the CPU oracle and, when
oracle/_ref is present, the unmodified reference CUDA kernels.  Tolerances (north star):
masks / candidate sets bit-exact; costs and disparities <= 1e-3 abs; var rel 1e-5 + abs 1e-3."""
import pytest
import torch

from helpers import LEVEL_SHAPES, make_feats, make_masks

pytestmark = pytest.mark.gpu

ATOL = 1e-3


def _cmp_forward(got, want, what):
    for k in want:
        assert torch.allclose(got[k].cpu(), want[k].cpu(), atol=ATOL, rtol=1e-5), \
            f"{what}: {k} max abs diff {(got[k].cpu() - want[k].cpu()).abs().max().item()}"


# variants: 1 one row per CTA, 2 persistent with next-row prefetch (both staged: cp.async or TMA path),
#           3 sector-gather kernel (the default), 4 its software-pipelined persistent form (no staging: path 3)
@pytest.mark.parametrize("variant", [1, 2, 3, 4])
@pytest.mark.parametrize("path", [1, 2])
@pytest.mark.parametrize("name,B,C,H,W,D", LEVEL_SHAPES)
@pytest.mark.parametrize("rho", [0.03, 0.3, 1.0])
def test_forward_vs_cpu_oracle(name, B, C, H, W, D, rho, path, variant):
    from decnet_b200 import ops, _lib
    from oracle import sparse as osp
    L, R = make_feats(B, C, H, W, device="cuda")
    ml, mr = make_masks(B, H, W, rho, rho, device="cuda")
    if variant >= 3:
        if path == 1:
            pytest.skip("the sector-gather kernel has one load path")
        path = 0
    _lib.lib().decnet_set_sparse_path(path)
    _lib.lib().decnet_set_sparse_variant(variant)
    try:
        try:
            out, ssim, mx = ops.spamat_forward(L, R, ml, mr, D)
        except _lib.DecnetError:
            if path == 2:
                pytest.skip("shape not eligible for the TMA path")
            raise
        var, ssim_v, mx_v = ops.spavar_forward(L, R, ml, mr, out, D)
        f_out, f_var, f_ssim, f_mx = ops.spamat_spavar_forward(L, R, ml, mr, D)
        assert _lib.lib().decnet_last_sparse_path() == (path if variant < 3 else 3)
        assert _lib.lib().decnet_last_sparse_variant() == variant
    finally:
        _lib.lib().decnet_set_sparse_path(0)
        _lib.lib().decnet_set_sparse_variant(0)
    o_out, o_ssim, o_mx = osp.spamat_forward(L, R, ml, mr, D)
    o_var, _, _ = osp.spavar_forward(L, R, ml, mr, o_out, D)
    want = {"out": o_out, "sum_sim": o_ssim, "max_cost": o_mx}
    _cmp_forward({"out": out, "sum_sim": ssim, "max_cost": mx}, want, "spamat")
    _cmp_forward({"out": f_out, "sum_sim": f_ssim, "max_cost": f_mx}, want, "fused")
    _cmp_forward({"var": var, "sum_sim": ssim_v, "max_cost": mx_v},
                 {"var": o_var, "sum_sim": o_ssim, "max_cost": o_mx}, "spavar")
    assert torch.allclose(f_var.cpu(), o_var, atol=ATOL, rtol=1e-4)
    # max_cost is an exact quantity (same FMA chain): must be bit-identical
    assert torch.equal(mx.cpu(), o_mx)
    # unmasked pixels are exactly zero in every output
    keep = (ml == 0)
    for t in (out, ssim, mx, var, f_out, f_var):
        assert t[keep].abs().max().item() == 0 if keep.any() else True


@pytest.mark.parametrize("name,B,C,H,W,D", LEVEL_SHAPES)
def test_candidate_sets_bit_exact(name, B, C, H, W, D):
    from decnet_b200 import ops
    from oracle import sparse as osp
    for rho in (0.05, 0.5):
        ml, mr = make_masks(B, H, W, rho, rho, device="cuda", seed=5)
        cnt, hsh = ops.candidate_signature(ml, mr, D)
        o_cnt, o_hsh = osp.candidate_signature(ml, mr, D)
        assert torch.equal(cnt.cpu(), o_cnt)
        assert torch.equal(hsh.cpu(), o_hsh)


def test_edge_semantics_gpu():
    from decnet_b200 import ops
    B, C, H, W, D = 1, 4, 2, 16, 6
    L, R = make_feats(B, C, H, W, device="cuda")
    ml = torch.zeros(B, H, W, device="cuda"); mr = torch.zeros(B, H, W, device="cuda")
    ml[0, 0, 5] = 1.0
    ml[0, 1, 7] = float("nan")              # NaN != 0 -> masked, like the reference's `== 0` test
    mr[0, 1, 7] = 1.0; mr[0, 1, 3] = -1.0; mr[0, 1, 0] = 1.0
    mr[0, 0, 2] = -0.0                      # negative zero is unmasked
    out, var, ssim, mx = ops.spamat_spavar_forward(L, R, ml, mr, D)
    assert out[0, 0, 5].item() == 1.0 and var[0, 0, 5].item() == 1.0
    assert ssim[0, 0, 5].item() == pytest.approx(1e-6) and mx[0, 0, 5].item() == pytest.approx(1e-6)
    cnt, _ = ops.candidate_signature(ml, mr, D)
    assert cnt[0, 1, 7].item() == 2 and cnt[0, 0, 5].item() == 0
    # max_disp <= 0 and numpy integer max_disp
    import numpy as np
    o0, _, _ = ops.spamat_forward(L, R, ml, mr, 0)
    assert o0[0, 1, 7].item() == 1.0
    from decnet_b200 import SpaMat
    o1 = SpaMat()(L, R, ml, mr, np.int64(D))
    assert torch.equal(o1, out)


@pytest.mark.parametrize("name,B,C,H,W,D", LEVEL_SHAPES[:2] + LEVEL_SHAPES[3:4] + LEVEL_SHAPES[5:])
def test_backward_vs_cpu_oracle(name, B, C, H, W, D):
    from decnet_b200 import SpaMat, SpaVar
    from oracle import sparse as osp
    L, R = make_feats(B, C, H, W, device="cuda")
    ml, mr = make_masks(B, H, W, 0.3, 0.3, device="cuda")
    g = torch.randn(B, H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    L1, R1 = L.clone().requires_grad_(), R.clone().requires_grad_()
    out = SpaMat()(L1, R1, ml, mr, D)
    (out * g).sum().backward()
    o_out, o_ssim, o_mx = osp.spamat_forward(L, R, ml, mr, D)
    dL, dR = osp.spamat_backward(L, R, ml, mr, o_out, o_ssim, o_mx, g, D)
    assert torch.allclose(L1.grad.cpu(), dL, atol=1e-4, rtol=1e-4)
    assert torch.allclose(R1.grad.cpu(), dR, atol=1e-4, rtol=1e-4)
    disp = (o_out + 0.37).cuda()
    L2, R2, d2 = L.clone().requires_grad_(), R.clone().requires_grad_(), disp.clone().requires_grad_()
    var = SpaVar()(L2, R2, ml, mr, d2, D)
    (var * g).sum().backward()
    o_var, o_ssim_v, o_mx_v = osp.spavar_forward(L, R, ml, mr, disp, D)
    dLv, dRv, dd = osp.spavar_backward(L, R, ml, mr, disp, o_var, o_ssim_v, o_mx_v, g, D)
    scale = max(1.0, dLv.abs().max().item())
    assert torch.allclose(L2.grad.cpu(), dLv, atol=1e-4 * scale, rtol=1e-4)
    assert torch.allclose(R2.grad.cpu(), dRv, atol=1e-4 * scale, rtol=1e-4)
    assert torch.allclose(d2.grad.cpu(), dd, atol=1e-4 * scale, rtol=1e-4)


@pytest.mark.parametrize("name,B,C,H,W,D", LEVEL_SHAPES[:4] + LEVEL_SHAPES[5:6])
def test_reference_cuda_pins_oracle_and_ours(name, B, C, H, W, D):
    """The UNMODIFIED reference kernels (oracle/_ref) on the same inputs: pins the CPU
    restatement and our kernels to the reference itself."""
    from oracle import ref_cuda, sparse as osp
    from decnet_b200 import ops
    if not ref_cuda.available():
        pytest.skip("oracle/_ref not built")
    L, R = make_feats(B, C, H, W, device="cuda")
    ml, mr = make_masks(B, H, W, 0.2, 0.2, device="cuda")
    g = torch.randn(B, H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    r_out, r_ssim, r_mx = ref_cuda.spamat_forward(L, R, ml, mr, D)
    r_var, r_ssim_v, r_mx_v = ref_cuda.spavar_forward(L, R, ml, mr, r_out, D)
    o_out, o_ssim, o_mx = osp.spamat_forward(L, R, ml, mr, D)
    o_var, _, _ = osp.spavar_forward(L, R, ml, mr, r_out, D)
    out, var, ssim, mx = ops.spamat_spavar_forward(L, R, ml, mr, D)
    torch.cuda.synchronize()
    for got, name_ in ((o_out, "oracle"), (out.cpu(), "ours")):
        assert torch.allclose(got, r_out.cpu(), atol=ATOL, rtol=1e-5), name_
    assert torch.equal(o_mx, r_mx.cpu()) and torch.equal(mx.cpu(), r_mx.cpu())
    assert torch.allclose(o_ssim, r_ssim.cpu(), atol=1e-5, rtol=1e-5)
    assert torch.allclose(ssim.cpu(), r_ssim.cpu(), atol=1e-5, rtol=1e-5)
    assert torch.allclose(o_var, r_var.cpu(), atol=ATOL, rtol=1e-5)
    assert torch.allclose(var.cpu(), r_var.cpu(), atol=ATOL, rtol=1e-4)
    # backward
    r_dL, r_dR = ref_cuda.spamat_backward(L, R, ml, mr, r_out, r_ssim, r_mx, g, D)
    o_dL, o_dR = osp.spamat_backward(L, R, ml, mr, r_out, r_ssim, r_mx, g, D)
    m_dL, m_dR = ops.spamat_backward(L, R, ml, mr, r_out, r_ssim, r_mx, g, D)
    torch.cuda.synchronize()
    for a, b in ((o_dL, r_dL), (o_dR, r_dR), (m_dL.cpu(), r_dL), (m_dR.cpu(), r_dR)):
        assert torch.allclose(a, b.cpu(), atol=1e-4, rtol=1e-4)
    r_g = ref_cuda.spavar_backward(L, R, ml, mr, r_out, r_var, r_ssim_v, r_mx_v, g, D)
    o_g = osp.spavar_backward(L, R, ml, mr, r_out, r_var, r_ssim_v, r_mx_v, g, D)
    m_g = ops.spavar_backward(L, R, ml, mr, r_out, r_var, r_ssim_v, r_mx_v, g, D)
    torch.cuda.synchronize()
    for a, b, c in zip(o_g, r_g, m_g):
        scale = max(1.0, b.abs().max().item())
        assert torch.allclose(a, b.cpu(), atol=1e-4 * scale, rtol=1e-4)
        assert torch.allclose(c.cpu(), b.cpu(), atol=1e-4 * scale, rtol=1e-4)


def test_multi_level_launch_equals_per_level_launches():
    """decnet_spamat_spavar_fwd_levels: the rows of several levels in one launch -- same row code, so the same bits as one
    launch per level (the model's three SceneFlow levels, a KITTI-style odd width, and a single level)."""
    from decnet_b200 import ops
    shapes = [(2, 8, 54, 972, 216), (2, 24, 18, 324, 72), (2, 72, 6, 108, 24), (1, 8, 9, 141, 24)]
    levels = []
    for i, (B, C, H, W, D) in enumerate(shapes):
        L, R = make_feats(B, C, H, W, seed=30 + i, device="cuda")
        ml, mr = make_masks(B, H, W, 0.15, 0.2, seed=40 + i, device="cuda")
        levels.append((L, R, ml, mr, D))
    for sel in (levels, levels[:3], levels[2:3], levels[::-1]):
        got = ops.spamat_spavar_forward_levels(sel)
        for lv, out in zip(sel, got):
            want = ops.spamat_spavar_forward(*lv)
            for a, b in zip(out, want):
                assert torch.equal(a, b)
