"""GPU: the kernels behind the coarse half of the feature extractor (csrc/featext.cu and the GEMM mode of
conv2d_nhwc_halo_kernel) against torch: im2col for the three source layouts, GEMM with column slices / bordered rows /
flat-to-bordered destination, the GEMM + shuffle form of ConvTranspose2d(k3, s3), channels-last -> NCHW, the direct stride-3
conv."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _g(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("stride,dil", [(1, 1), (3, 1), (1, 4), (1, 12)])
def test_im2col3x3_all_layouts(stride, dil):
    from decnet_b200 import ops
    g = _g(1)
    B, C, H, W = 2, 24, 27, 36
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    cols = F.unfold(x, 3, dilation=dil, padding=dil, stride=stride)                     # [B, C*9, L], row c*9 + tap
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    want = cols.view(B, C, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, 9 * C)    # column tap*C + c
    a, ho, wo = ops.im2col3x3(x, "nchw", C, stride, dil)
    assert (ho, wo) == (Ho, Wo) and torch.equal(a[:, : 9 * C], want) and float(a[:, 9 * C:].abs().sum()) == 0
    xl = x.permute(0, 2, 3, 1).contiguous()
    b_, _, _ = ops.im2col3x3(xl, "nhwc", C, stride, dil)
    assert torch.equal(b_, a)
    xp = F.pad(F.pad(xl, (0, 8)), (0, 0, 1, 1, 1, 1)).contiguous()                       # bordered, rows of C + 8 floats
    c_, _, _ = ops.im2col3x3(xp, "nhwc_pad", C, stride, dil)
    assert torch.equal(c_, a)


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("P,K,N", [(720, 216, 216), (1000, 72, 72), (333, 1944, 108), (129, 16, 16)])
def test_gemm_mode_with_slices(P, K, N, split):
    from decnet_b200 import ops
    g = _g(2)
    cp = (K + 7) // 8 * 8
    x = torch.zeros(P, cp, device="cuda")
    x[:, :K] = torch.randn(P, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * (1.0 / K) ** 0.5
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    want = F.relu(x[:, :K].double() @ w.double().t() + b.double()).float()
    ld = (N + 15) // 16 * 16 + 16
    out = torch.zeros(P, ld, device="cuda")
    step = 108 if N > 128 else N
    for a in range(0, N, step):
        e = min(a + step, N)
        wp, bp, np_ = ops.pack_gemm_weights(w, b, cp, split, a, e)
        ops.gemm_tc(x if split else ops.rna_tf32(x), wp, bp, out, True, split=split, col=a)
    err = float((out[:, :N] - want).abs().max())
    scale = max(1.0, float(want.abs().max()))
    assert err <= (1e-5 if split else 4 * 2 ** -11) * scale, (err, scale)
    assert float(out[:, N:].abs().max()) == 0                                           # slice padding columns are zeros


def test_gemm_bordered_rows_and_flat_to_bordered():
    from decnet_b200 import ops
    g = _g(3)
    B, h, w, K, N = 2, 7, 9, 72, 72
    xin = torch.randn(B, h, w, 80, device="cuda", generator=g)
    xin[..., K:] = 0
    xp = F.pad(xin, (0, 0, 1, 1, 1, 1)).contiguous()
    wt = torch.randn(N, K, device="cuda", generator=g) * 0.1
    bs = torch.randn(N, device="cuda", generator=g)
    wp, bp, np_ = ops.pack_gemm_weights(wt, bs, 80, True)
    want = F.relu(xin[..., :K].double() @ wt.double().t() + bs.double()).float()
    out = torch.full((B, h + 2, w + 2, 152), 7.0, device="cuda")
    ops.gemm_tc(xp, wp, bp, out, True, split=True, col=72, border=(B, h, w))
    assert float((out[:, 1:-1, 1:-1, 72:72 + N] - want).abs().max()) <= 1e-5 * float(want.abs().max())
    assert float(out[:, 0, :, 72:152].abs().max()) == 0 and float(out[:, :, -1, 72:152].abs().max()) == 0   # border rows -> zeros
    assert float((out[..., :72] - 7.0).abs().max()) == 0                                # the other channel slice is untouched
    out2 = torch.zeros(B, h + 2, w + 2, 80, device="cuda")
    ops.gemm_tc(xin.reshape(-1, 80), wp, bp, out2, True, split=True, dst_hw=(h, w))
    assert float((out2[:, 1:-1, 1:-1, :N] - want).abs().max()) <= 1e-5 * float(want.abs().max())
    assert float(out2[:, 0].abs().max()) == 0 and float(out2[:, :, 0].abs().max()) == 0


def test_deconv_as_gemm_plus_shuffle_and_nhwc_to_nchw():
    from decnet_b200 import ops
    g = _g(4)
    B, h, w, Cin, Cout = 2, 5, 6, 216, 72
    x = torch.randn(B, Cin, h, w, device="cuda", generator=g)
    wt = torch.randn(Cin, Cout, 3, 3, device="cuda", generator=g) * 0.05
    bs = torch.randn(Cout, device="cuda", generator=g) * 0.1
    want = F.relu(F.conv_transpose2d(x.double(), wt.double(), bs.double(), stride=3)).float()
    xl = torch.zeros(B * h * w, 224, device="cuda")
    xl[:, :Cin] = x.permute(0, 2, 3, 1).reshape(-1, Cin)
    wd = wt.permute(2, 3, 1, 0).reshape(9 * Cout, Cin).contiguous()
    up = torch.zeros(B * h * w, 9 * Cout + 16, device="cuda")
    for a in range(0, 9 * Cout, 108):
        wp, bp, _ = ops.pack_gemm_weights(wd, bs.repeat(9), 224, True, a, a + 108)
        ops.gemm_tc(xl, wp, bp, up, True, split=True, col=a)
    out = torch.zeros(B, 3 * h + 2, 3 * w + 2, 152, device="cuda")
    ops.deconv3x3s3_shuffle(up, out, B, h, w, Cout, col=0)
    got = ops.nhwc_to_nchw(out, B, Cout, 3 * h, 3 * w, pad=True)
    assert float((got - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))
    flat = torch.randn(B, h, w, 224, device="cuda", generator=g)
    assert torch.equal(ops.nhwc_to_nchw(flat, B, 216, h, w), flat[..., :216].permute(0, 3, 1, 2))


def test_conv3x3s3_nchw_direct():
    from decnet_b200 import ops
    g = _g(5)
    x = torch.randn(2, 8, 54, 81, device="cuda", generator=g)
    w = torch.randn(24, 8, 3, 3, device="cuda", generator=g) * 0.1
    b = torch.randn(24, device="cuda", generator=g) * 0.1
    want = F.relu(F.conv2d(x.double(), w.double(), b.double(), stride=3, padding=1)).float()
    got = ops.conv3x3s3_nchw(x, w, b, True)
    assert got.shape == want.shape and float((got - want).abs().max()) <= 2e-5


def test_feature_extractor_has_no_library_layer():
    from decnet_b200.features import FeatExtNetChannelPlus
    from decnet_b200.model import Conv2dUnit, Deconv2dUnit
    fe = FeatExtNetChannelPlus(8)
    assert not any(hasattr(m, "library_ok") for m in fe.modules() if isinstance(m, (Conv2dUnit, Deconv2dUnit)))   # no library branch
