"""GPU parity of the tcgen05 implicit-GEMM Conv3d (through the C ABI).
Tolerance (north star): the bf16 aggregation may move the disparity by <= 0.05 px EPE; single layers are
checked against an fp32 torch convolution of the SAME bf16-rounded operands (so only accumulation order
and the bf16 output rounding differ)."""
import pytest
import torch
import torch.nn.functional as F

from golden_util import CASES, gold_list, load_case

pytestmark = pytest.mark.gpu


def _ref_layer(x_ndhwc, w_packed, bias, relu, residual=None):
    """fp32 reference on the bf16-rounded operands: x [B,D,H,W,CP], w [27][NP][CP]."""
    torch.backends.cudnn.allow_tf32 = False
    x = x_ndhwc.float().permute(0, 4, 1, 2, 3).contiguous()
    NP, CP = w_packed.shape[1:]
    w = w_packed.float().view(3, 3, 3, NP, CP).permute(3, 4, 0, 1, 2).contiguous()
    y = F.conv3d(x, w, bias, padding=1)
    if relu:
        y = F.relu(y)
    y = y.permute(0, 2, 3, 4, 1)
    if residual is not None:
        y = y + residual.float()
    return y


@pytest.mark.parametrize("B,D,H,W,C,Cout", [(1, 8, 20, 36, 216, 216), (2, 8, 14, 47, 216, 216), (1, 3, 5, 7, 32, 48),
                                              (1, 8, 4, 6, 216, 1), (1, 29, 9, 11, 64, 64), (1, 1, 1, 1, 16, 16)])
@pytest.mark.parametrize("relu,use_res", [(True, False), (True, True), (False, False)])
def test_single_layer_vs_fp32_conv(B, D, H, W, C, Cout, relu, use_res):
    from decnet_b200 import conv3d as c3
    g = torch.Generator(device="cuda").manual_seed(11)
    cp, np_ = c3._pad16(C), c3._pad16(Cout)
    x = torch.zeros(B, D, H, W, cp, device="cuda", dtype=torch.bfloat16)
    x[..., :C] = (torch.randn(B, D, H, W, C, device="cuda", generator=g)).to(torch.bfloat16)
    w = torch.zeros(27, np_, cp, device="cuda", dtype=torch.bfloat16)
    w[:, :Cout, :C] = (torch.randn(27, Cout, C, device="cuda", generator=g) * (2.0 / (27 * C)) ** 0.5).to(torch.bfloat16)
    bias = torch.zeros(np_, device="cuda"); bias[:Cout] = torch.randn(Cout, device="cuda", generator=g) * 0.1
    res = None
    if use_res:
        if Cout == 1:
            pytest.skip("no residual on the single-channel layer")
        res = torch.zeros(B, D, H, W, np_, device="cuda", dtype=torch.bfloat16)
        res[..., :Cout] = torch.randn(B, D, H, W, Cout, device="cuda", generator=g).to(torch.bfloat16)
    want = _ref_layer(x, w, bias, relu, res)
    if Cout == 1:
        got = c3.conv3d_layer(x, w, bias, np_, relu, out_f32=True)
        assert torch.allclose(got, want[..., 0], atol=2e-3, rtol=1e-3), (got - want[..., 0]).abs().max()
    else:
        got = c3.conv3d_layer(x, w, bias, np_, relu, residual=res).float()
        err = (got - want).abs().max().item()
        scale = want.abs().max().item()
        assert err <= 2 ** -7 * scale + 1e-3, (err, scale)          # bf16 output rounding (2^-8 rel) + slack
        if np_ > Cout:
            assert got[..., Cout:].abs().max().item() == 0          # padded channels stay exactly zero


@pytest.mark.parametrize("B,D,H,W,C", [(8, 8, 20, 36, 216), (2, 8, 14, 47, 216), (1, 5, 7, 9, 64), (1, 8, 1, 3, 32),
                                        (1, 29, 9, 11, 64), (1, 16, 6, 10, 48), (1, 1, 4, 4, 16)])
def test_last_layer_with_fused_softargmin(B, D, H, W, C):
    """The 216->1 layer + soft-argmin in one launch (in the epilogue when one tile box spans D, else the second kernel):
    cost identical to the plain mode-1 layer and pred identical, bit for bit, to decnet_softargmin of that cost."""
    from decnet_b200 import _lib, conv3d as c3, ops
    g = torch.Generator(device="cuda").manual_seed(5)
    cp = c3._pad16(C)
    x = torch.zeros(B, D, H, W, cp, device="cuda", dtype=torch.bfloat16)
    x[..., :C] = torch.randn(B, D, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.zeros(27, 16, cp, device="cuda", dtype=torch.bfloat16)
    w[:, :1, :C] = (torch.randn(27, 1, C, device="cuda", generator=g) * (8.0 / (27 * C)) ** 0.5).to(torch.bfloat16)
    bias = torch.zeros(16, device="cuda"); bias[0] = 0.3
    cost_ref = c3.conv3d_layer(x, w, bias, 16, False, out_f32=True)
    pred_ref = ops.softargmin(cost_ref)
    cost = torch.full((B, D, H, W), float("nan"), device="cuda")
    pred = torch.full((B, H, W), float("nan"), device="cuda")
    st = _lib.lib().decnet_conv3d_bf16_softargmin(x.data_ptr(), w.data_ptr(), bias.data_ptr(), cost.data_ptr(), pred.data_ptr(),
                                                  B, D, H, W, cp, 16, 0, torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "decnet_conv3d_bf16_softargmin")
    assert torch.equal(cost, cost_ref)
    assert torch.equal(pred, pred_ref)
    assert 0 <= pred.min().item() and pred.max().item() <= D - 1


@pytest.mark.parametrize("name", CASES)
def test_stack_vs_reference_golden(name):
    """Full a2+a3+a4 with the bf16 tcgen05 stack against the reference's fp32 golden coarse disparity."""
    from decnet_b200.model import DecompMatching
    z, P, left, right, lmasks, rmasks, cfg = load_case(name, device="cuda")
    m = DecompMatching(max_disp=cfg["max_disp"], skip_stage_id=cfg["skip_stage_id"], use_detail=cfg["use_detail"],
                       thold=cfg["thold"])
    m.load_state_dict(P)
    m = m.cuda()
    pred0, cost = m.dense_stage(left["stage0"], right["stage0"], cfg["max_disp"] // 27)
    gpred = gold_list(z, "pred", "cuda")[0]
    gcost = torch.from_numpy(z["cost"]).cuda()
    epe = (pred0 - gpred).abs().mean().item()
    assert epe <= 0.05, f"coarse EPE delta {epe}"
    rel = (cost - gcost).abs().max().item() / gcost.abs().max().item()
    assert rel <= 0.05, f"cost relative error {rel}"
    # drop-in forward([B,C,D,H,W]) route gives the same numbers as the fused route
    cost2 = m.cost_regularizer(torch.from_numpy(z["vol"]).cuda())
    assert (cost2 - cost).abs().max().item() <= 0.02 * gcost.abs().max().item()


@pytest.mark.parametrize("B,D,H,W,C,Cout", [(2, 8, 20, 36, 216, 216), (1, 8, 14, 47, 216, 216), (1, 8, 4, 6, 216, 1), (3, 3, 5, 7, 32, 48)])
def test_cta_pair_variant_matches_single_cta(B, D, H, W, C, Cout):
    """The cta_group::2 kernel (opt-in) must give the same numbers as the default single-CTA kernel."""
    from decnet_b200 import _lib, conv3d as c3
    g = torch.Generator(device="cuda").manual_seed(12)
    cp, np_ = c3._pad16(C), c3._pad16(Cout)
    x = torch.zeros(B, D, H, W, cp, device="cuda", dtype=torch.bfloat16)
    x[..., :C] = torch.randn(B, D, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.zeros(27, np_, cp, device="cuda", dtype=torch.bfloat16)
    w[:, :Cout, :C] = (torch.randn(27, Cout, C, device="cuda", generator=g) * (2.0 / (27 * C)) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(np_, device="cuda", generator=g) * 0.1
    f32 = Cout == 1
    a = c3.conv3d_layer(x, w, bias, np_, True, out_f32=f32)
    _lib.lib().decnet_conv3d_set_variant(2)
    try:
        b_ = c3.conv3d_layer(x, w, bias, np_, True, out_f32=f32)
    finally:
        _lib.lib().decnet_conv3d_set_variant(0)
    assert torch.equal(a, b_), (a.float() - b_.float()).abs().max()


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 60, 108, 217, 81), (1, 20, 36, 649, 81), (2, 45, 50, 81, 81), (1, 180, 324, 73, 81),
                                            (1, 7, 5, 16, 16)])
@pytest.mark.parametrize("relu", [True, False])
def test_conv2d_tf32_tcgen05_vs_fp32(B, H, W, Cin, Cout, relu):
    """kind::tf32 implicit GEMM (DynamicUpsampling's 81-channel convs) against cuDNN fp32: TF32 rounds the
    operands to 10 mantissa bits, so the gate is the TF32 class (2e-3 of the output scale), like cuDNN's default."""
    import torch.nn.functional as F
    from decnet_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(21)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    want = F.conv2d(x, w, b, padding=1)
    if relu:
        want = F.relu(want)
    cp = (Cin + 7) // 8 * 8
    xn = torch.zeros(B, H, W, cp, device="cuda")
    xn[..., :Cin] = x.permute(0, 2, 3, 1)
    wp, bp, np_ = ops.pack_conv2d_tf32_weights(w, b, cp)
    xn = ((xn.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)      # operands pre-rounded to TF32
    got = ops.conv2d_tf32_nhwc(xn.contiguous(), wp, bp, relu)
    assert got.shape == (B, H, W, np_)
    err = (got[..., :Cout].permute(0, 3, 1, 2) - want).abs().max().item()
    assert err <= 2e-3 * want.abs().max().item() + 1e-4, (err, want.abs().max().item())
    if np_ > Cout:
        assert got[..., Cout:].abs().max().item() == 0


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 60, 108, 217, 81), (1, 20, 36, 649, 81), (2, 45, 50, 81, 81), (1, 180, 324, 73, 81),
                                            (1, 60, 108, 145, 72), (1, 21, 47, 72, 36), (1, 7, 5, 16, 16)])
@pytest.mark.parametrize("relu", [True, False])
def test_conv2d_3xtf32_nhwc_halo_is_fp32_class(B, H, W, Cin, Cout, relu):
    """decnet_conv2d_tc_nhwc_halo, split = 1 (the product default): unrounded fp32 activations, hi/lo weights, three MMAs per
    tap -> an fp64 convolution reproduced at fp32 level; the zero border is kept."""
    import torch.nn.functional as F
    from decnet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(22)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    want = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    want = (F.relu(want) if relu else want).float()
    cp = (Cin + 7) // 8 * 8
    xn = ops.nchw_cat_to_nhwc_pad([x], cp, round_tf32=False)
    wp, bp, np_ = ops.pack_conv2d_tf32_weights(w, b, cp, split=True)
    assert wp.shape == (18, np_, cp)
    got = ops.conv2d_tf32_nhwc_halo(xn, wp, bp, relu, split=True)
    assert got.shape == (B, H + 2, W + 2, np_)
    assert float(got[:, 0].abs().max()) == 0 and float(got[:, -1].abs().max()) == 0
    assert float(got[:, :, 0].abs().max()) == 0 and float(got[:, :, -1].abs().max()) == 0
    err = (got[:, 1:-1, 1:-1, :Cout].permute(0, 3, 1, 2) - want).abs().max().item()
    assert err <= 1e-5 * max(1.0, want.abs().max().item()), (err, want.abs().max().item())
    if np_ > Cout:
        assert got[..., Cout:].abs().max().item() == 0


@pytest.mark.parametrize("kind", [1, 2])
@pytest.mark.parametrize("xs,ws", [(1.0, 1.0), (300.0, 1e-3), (1e-3, 30.0), (3e3, 1e-4)])
def test_halo_split_kinds_and_operand_ranges(kind, xs, ws, monkeypatch):
    """Both fp32-class forms of the channels-last kernel (1: 3xTF32; 2, the default: TF32 hi*hi + fp16 K-concatenated corrections)
    across operand magnitudes, conv and GEMM mode, a ragged last channel chunk (cp = 88 -> chunks of 32, 32, 24)."""
    import torch.nn.functional as F
    from decnet_b200 import ops
    monkeypatch.setattr(ops, "SPLIT_KIND", kind)
    g = torch.Generator(device="cuda").manual_seed(91)
    B, Cin, Cout, H, W = 2, 81, 81, 40, 52
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g) * xs
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5 * ws
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1 * xs * ws
    want = F.conv2d(x.double(), w.double(), b.double(), padding=1).float()
    cp = (Cin + 7) // 8 * 8
    xn = ops.nchw_cat_to_nhwc_pad([x], cp, round_tf32=False)
    wp, bp, np_ = ops.pack_conv2d_tf32_weights(w, b, cp, split=True)
    assert bp.numel() == np_ + 4
    got = ops.conv2d_tf32_nhwc_halo(xn, wp, bp, False, split=True)
    err = (got[:, 1:-1, 1:-1, :Cout].permute(0, 3, 1, 2) - want).abs().max().item()
    assert err <= 1e-5 * want.abs().max().item(), (kind, xs, ws, err, want.abs().max().item())
    # GEMM mode (taps = 1) on the same operands: the centre tap as a 1x1 conv
    w1 = w[:, :, 1, 1].contiguous()
    wg, bg, npg = ops.pack_gemm_weights(w1, b, cp, split=True)
    rows = xn[:, 1:-1, 1:-1].reshape(-1, cp).contiguous()
    out = torch.empty(rows.shape[0], npg, device="cuda")
    ops.gemm_tc(rows, wg, bg, out, False, split=True)
    want1 = (rows[:, :Cin].double() @ w1.double().t() + b.double()).float()
    err1 = (out[:, :Cout] - want1).abs().max().item()
    assert err1 <= 1e-5 * want1.abs().max().item(), (kind, xs, ws, err1)


@pytest.mark.parametrize("B,H,W,Cin,Cout,relu", [(1, 180, 324, 73, 81, True), (3, 120, 107, 81, 81, False), (2, 200, 150, 145, 72, True),
                                                  (6, 61, 109, 72, 36, True)])
def test_pair_kernel_matches_one_tile_per_cta(B, H, W, Cin, Cout, relu):
    """conv2d_nhwc_pair_kernel (opt-in variant 2: two tiles per CTA share each weight stage) against the default
    one-tile-per-CTA kernel: same MMAs per accumulator, other drain points -> equal to fp32 rounding; odd tile counts and the
    zero border included."""
    from decnet_b200 import _lib, ops
    g = torch.Generator(device="cuda").manual_seed(31)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    cp = (Cin + 7) // 8 * 8
    xn = ops.nchw_cat_to_nhwc_pad([x], cp, round_tf32=False)
    wp, bp, np_ = ops.pack_conv2d_tf32_weights(w, b, cp, split=True)
    assert B * (H + 2) * (W + 2) >= 2 * 148 * 128
    want = ops.conv2d_tf32_nhwc_halo(xn, wp, bp, relu, split=True)
    _lib.lib().decnet_conv2d_nhwc_set_variant(2)
    try:
        got = ops.conv2d_tf32_nhwc_halo(xn, wp, bp, relu, split=True)
    finally:
        _lib.lib().decnet_conv2d_nhwc_set_variant(0)
    scale = max(1.0, float(want.abs().max()))
    assert float((got - want).abs().max()) <= 2e-6 * scale
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1)
    ref = (torch.relu(ref) if relu else ref).float()
    assert float((got[:, 1:-1, 1:-1, :Cout].permute(0, 3, 1, 2) - ref).abs().max()) <= 1e-5 * scale
    assert float(got[:, 0].abs().max()) == 0 and float(got[:, :, -1].abs().max()) == 0
