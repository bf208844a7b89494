"""GPU parity of the NCHW tcgen05 Conv2d (conv2d_tcgen05.cu) through the C ABI, in both arithmetic modes: plain TF32 and the
error-compensated 3xTF32 ("fp32" precision, the product default).

Checks per shape:
  * 3xTF32: raw fp32 operands against an fp64 convolution at fp32 level (1e-5 of the output scale);
  * exactness of the data path: with operands that are already TF32 values the kernel must agree with an
    fp64 convolution up to fp32 accumulation error (any wrong tap, swizzle, halo or padding shows as O(1));
  * precision class: with arbitrary fp32 operands the deviation from the fp32 result must stay at TF32
    rounding level (operands rounded to nearest, 2^-11 relative each) -- the tolerance is written below.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _tf32(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


SHAPES = [  # B, Cin, Cout, H, W, dilation
    (1, 8, 8, 64, 96, 1),
    (2, 8, 8, 37, 100, 1),        # ragged rows / columns
    (1, 17, 8, 50, 120, 3),       # refinement stage 3: 2C+1 channels, dilation 3
    (1, 8, 4, 41, 64, 6),
    (1, 4, 4, 45, 88, 9),
    (1, 4, 1, 19, 32, 1),         # single output channel
    (1, 8, 3, 33, 60, 1),         # detail-detection head
    (1, 12, 8, 40, 76, 1),        # attention input C+4
    (2, 24, 8, 30, 36, 1),
    (1, 49, 24, 60, 108, 2),      # refinement stage 2
    (1, 24, 12, 20, 108, 4),
    (1, 72, 8, 20, 36, 1),
    (1, 3, 3, 8, 8, 1),           # tile larger than the image
    (8, 8, 8, 135, 240, 1),       # many tiles per CTA
]


@pytest.mark.parametrize("B,Cin,Cout,H,W,dil", SHAPES)
@pytest.mark.parametrize("relu", [True, False])
def test_conv2d_tf32_nchw(B, Cin, Cout, H, W, dil, relu):
    from decnet_b200 import ops
    assert ops.conv2d_tf32_supported(Cin, Cout, H, W, dil)
    g = torch.Generator(device="cuda").manual_seed(5 + Cin + 7 * Cout + dil)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b)

    def ref(xx, ww):
        y = F.conv2d(xx.double(), ww.double(), b.double(), padding=dil, dilation=dil)
        return (F.relu(y) if relu else y).float()

    # (1) exact data path
    xr = _tf32(x)
    got = ops.conv2d_tf32_nchw(xr, wp, bp, Cout, dil, relu)
    want = ref(xr, _tf32(w))
    err = (got - want).abs().max().item()
    assert err <= 2e-5 * max(1.0, want.abs().max().item()), ("exact path", err)
    # (2) TF32 precision class on raw fp32 operands: |err| <= ~ sqrt(K) * 2^-11 * |x||w| ; measured against
    # the output scale with K = 9*Cin terms of unit variance -> 4 * 2^-11 * max|y| is a loose bound
    got2 = ops.conv2d_tf32_nchw(x, wp, bp, Cout, dil, relu)
    want2 = ref(x, w)
    err2 = (got2 - want2).abs().max().item()
    assert err2 <= 4 * 2 ** -11 * max(1.0, want2.abs().max().item()), ("tf32 class", err2)
    # rounding in the kernel == rounding beforehand (bit-exact: same operands reach the MMA)
    assert torch.equal(got, got2) or (got - got2).abs().max().item() <= 4 * 2 ** -11 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("B,Cin,Cout,H,W,dil", SHAPES + [(1, 36, 36, 60, 108, 1), (1, 36, 1, 60, 108, 1), (2, 28, 8, 180, 324, 1)])
@pytest.mark.parametrize("relu", [True, False])
def test_conv2d_3xtf32_nchw_is_fp32_class(B, Cin, Cout, H, W, dil, relu):
    """split mode: hi/lo operand split in the kernel (activations) and on the host (weights), three MMAs per tap.  Raw fp32
    operands must reproduce the fp64 result at fp32 level -- two orders of magnitude below the TF32 class."""
    from decnet_b200 import ops
    assert ops.conv2d_tf32_supported(Cin, Cout, H, W, dil, split=True)
    g = torch.Generator(device="cuda").manual_seed(5 + Cin + 7 * Cout + dil)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b, split=True)
    got = ops.conv2d_tf32_nchw_cat([x], wp, bp, Cout, dil, relu, split=True)
    y = F.conv2d(x.double(), w.double(), b.double(), padding=dil, dilation=dil)
    want = (F.relu(y) if relu else y).float()
    err = (got - want).abs().max().item()
    assert err <= 1e-5 * max(1.0, want.abs().max().item()), ("3xTF32", err, want.abs().max().item())


@pytest.mark.parametrize("kind", [1, 2])
@pytest.mark.parametrize("xs,ws", [(1.0, 1.0), (300.0, 1e-3), (1e-3, 30.0), (3e3, 1e-4), (1e-5, 1.0)])
def test_conv2d_split_kinds_and_operand_ranges(kind, xs, ws, monkeypatch):
    """Both fp32-class forms of the thin kernel (1: three TF32 MMAs per tap; 2, the default: TF32 hi*hi + one fp16 correction MMA)
    on operands of very different magnitudes, and weights whose channels differ by 2^12: the fp16 correction operand is scaled by a
    per-layer power of two, so large / small weights and activations keep the fp32-class error (relative to the output scale)."""
    from decnet_b200 import ops
    monkeypatch.setattr(ops, "SPLIT_KIND", kind)
    g = torch.Generator(device="cuda").manual_seed(77)
    B, srcs, Cout, H, W, dil = 2, [8, 8, 1], 8, 64, 96, 2
    x = [torch.randn(B, c, H, W, device="cuda", generator=g) * xs for c in srcs]
    w = torch.randn(Cout, sum(srcs), 3, 3, device="cuda", generator=g) * (2.0 / (9 * sum(srcs))) ** 0.5 * ws
    w[::2] *= 2.0 ** -12                                              # wide dynamic range between output channels
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1 * xs * ws
    b[::2] *= 2.0 ** -12
    wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b, srcs, split=True)
    assert bp.numel() == 8 + 4
    got = ops.conv2d_tf32_nchw_cat(x, wp, bp, Cout, dil, False, split=True)
    want = F.conv2d(torch.cat(x, 1).double(), w.double(), b.double(), padding=dil, dilation=dil).float()
    for sl in (slice(0, None, 2), slice(1, None, 2)):                 # the small and the large channels, each against its own scale
        err = (got[:, sl] - want[:, sl]).abs().max().item()
        assert err <= 1e-5 * want[:, sl].abs().max().item(), (kind, xs, ws, sl, err, want[:, sl].abs().max().item())


def test_conv2d_split16_beyond_fp16_range_degrades_to_tf32_class():
    """split kind 2 keeps fp32-class accuracy for |x| < 32752 (2^11 * lo(x) must fit fp16); beyond that the correction operand
    saturates (cvt.rn.satfinite) and the affected products fall back to the TF32 class -- finite, never inf / NaN."""
    from decnet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(78)
    x = torch.randn(1, 8, 32, 64, device="cuda", generator=g) * 2e5
    w = torch.randn(8, 8, 3, 3, device="cuda", generator=g) * 0.1
    b = torch.zeros(8, device="cuda")
    wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b, split=True)
    got = ops.conv2d_tf32_nchw_cat([x], wp, bp, 8, 1, False, split=True)
    want = F.conv2d(x.double(), w.double(), None, padding=1).float()
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() <= 4 * 2 ** -11 * want.abs().max().item()


def test_conv2d_tf32_unsupported_shapes_are_refused():
    from decnet_b200 import ops
    assert not ops.conv2d_tf32_supported(8, 8, 64, 97, 1)         # W*4 not a multiple of 16 (TMA stride rule)
    assert not ops.conv2d_tf32_supported(145, 72, 60, 108, 1)     # weights do not fit in shared memory
    assert not ops.conv2d_tf32_supported(8, 8, 64, 96, 13)


def test_conv2d_tf32_matches_direct_kernel_on_layer_shapes():
    """Same layer through the fp32 direct kernel and the tensor-core kernel (SceneFlow stage-3 size, B=1)."""
    from decnet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(1, 8, 540, 972, device="cuda", generator=g)
    w = torch.randn(8, 8, 3, 3, device="cuda", generator=g) * 0.17
    b = torch.randn(8, device="cuda", generator=g) * 0.1
    wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b)
    got = ops.conv2d_tf32_nchw(x, wp, bp, 8, 1, True)
    want = ops.conv2d_small(x, ops.pack_conv2d_weights(w), b.contiguous(), 8, 3, 1, True)
    assert (got - want).abs().max().item() <= 4 * 2 ** -11 * want.abs().max().item()


@pytest.mark.parametrize("chans,Cout,H,W,dil", [((8, 4), 8, 45, 64, 1), ((24, 4), 24, 30, 36, 1), ((8, 8, 1), 8, 50, 120, 3),
                                                ((24, 24, 1), 24, 36, 108, 2), ((5, 1, 3), 4, 20, 32, 1)])
def test_conv2d_tf32_cat_equals_conv_of_concatenation(chans, Cout, H, W, dil):
    """decnet_conv2d_tf32_nchw_cat: each source occupies whole 8-channel chunks; the result must equal the
    single-source kernel run on the materialised torch.cat (same TF32 operands -> same accumulation)."""
    from decnet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(17)
    srcs = [torch.randn(2, c, H, W, device="cuda", generator=g) if c > 1 else torch.randn(2, H, W, device="cuda", generator=g)
            for c in chans]
    cin = sum(chans)
    w = torch.randn(Cout, cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b, chans)
    got = ops.conv2d_tf32_nchw_cat(srcs, wp, bp, Cout, dil, True)
    x = torch.cat([t.unsqueeze(1) if t.dim() == 3 else t for t in srcs], 1).contiguous()
    want = F.relu(F.conv2d(_tf32(x).double(), _tf32(w).double(), b.double(), padding=dil, dilation=dil)).float()
    err = (got - want).abs().max().item()
    assert err <= 2e-5 * max(1.0, want.abs().max().item()), err


def test_model_units_use_the_cat_path_and_match_packed_path():
    """SoftAttention.logits_cat / Refinement (three-source first conv) against the packed-tensor route."""
    from decnet_b200 import model as dm, ops
    torch.manual_seed(0)
    B, C, H, W = 2, 8, 48, 64
    att = dm.SoftAttention(C + 4, C).cuda().eval()
    ref = dm.Refinement(C, stage_id=3).cuda().eval()
    L, R = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
    dense, sparse = torch.rand(B, H, W, device="cuda") * 20, torch.rand(B, H, W, device="cuda") * 20
    mask, var = (torch.rand(B, H, W, device="cuda") > 0.5).float(), torch.rand(B, H, W, device="cuda")
    with torch.no_grad():
        a = att.logits_cat(L, ops.attn_pack(None, dense, sparse, mask, var))
        b = att.logits(ops.attn_pack(L, dense, sparse, mask, var))
        assert (a - b).abs().max().item() <= 1e-4 * max(1.0, b.abs().max().item())
        p1, r1 = ref(L, R, dense)
        p0, r0 = ref.forward_packed(ops.refine_pack(L, R, dense), dense)      # materialised cat, single-source first conv
        assert (r1 - r0).abs().max().item() <= 2e-5 * max(1.0, r0.abs().max().item())     # same 3xTF32 arithmetic, other chunking
        from oracle import glue as og
        P = {f"rf.{k}": v.detach().cpu().double() for k, v in ref.state_dict().items()}
        _, r64 = og.refinement(L.cpu().double(), R.cpu().double(), dense.cpu().double(), P, "rf", 3)
        assert (r1.cpu() - r64.float()).abs().max().item() <= 1e-4 * max(1.0, float(r64.abs().max()))   # fp32 class through 7 layers


@pytest.mark.parametrize("chans,Cout,H,W,dil", [((8,), 8, 64, 96, 1), ((8,), 8, 37, 100, 1), ((8, 8, 1), 8, 50, 120, 3),
                                                ((8, 4), 8, 45, 200, 1), ((5,), 3, 33, 60, 2), ((24,), 8, 60, 108, 1)])
def test_conv2d_tf32_rows_formulation(chans, Cout, H, W, dil):
    """conv2d_rows_tcgen05.cu (pixels on N, block-Toeplitz weights, column taps as accumulator column offsets):
    opt-in second formulation, kept with its measurements in DESIGN.md section 3.3; same exactness bar."""
    from decnet_b200 import ops
    assert ops.conv2d_tf32_rows_supported(ops.padded_cat_channels(chans), Cout, H, W, dil)
    g = torch.Generator(device="cuda").manual_seed(23)
    srcs = [_tf32(torch.randn(2, c, H, W, device="cuda", generator=g)) for c in chans]
    cin = sum(chans)
    w = _tf32(torch.randn(Cout, cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    wc, b8 = ops.pack_conv2d_tf32_rows_weights(w, b, chans)
    got = ops.conv2d_tf32_rows_nchw_cat(srcs, wc, b8, Cout, dil, False)
    want = F.conv2d(torch.cat(srcs, 1).double(), w.double(), b.double(), padding=dil, dilation=dil).float()
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())
    assert not ops.conv2d_tf32_rows_supported(24, 8, 60, 108, 4)          # band matrices + two stages exceed shared memory


def test_widths_not_multiple_of_4_run_on_pitch_padded_copies():
    """KITTI-style widths (1269, 423, 141): the unit pads the row pitch to 16 bytes, keeps the padding at zero through
    a stack (w_valid) and crops; result = an fp64 evaluation of the reference arithmetic on the unpadded tensor, at fp32 level."""
    from decnet_b200 import model as dm, ops
    from oracle import glue as og
    torch.manual_seed(4)
    B, C, H, W = 2, 8, 42, 141
    L, R = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
    disp = torch.rand(B, H, W, device="cuda") * 20
    ref = dm.Refinement(C, stage_id=3).cuda().eval()
    att = dm.SoftAttention(C + 4, C).cuda().eval()
    xp, wv = dm.pad_pitch(L)
    assert wv == W and xp.shape[-1] == 144 and float(xp[..., W:].abs().max()) == 0
    with torch.no_grad():
        y = ref.conv[1](ref.conv[1](xp, w_valid=wv), w_valid=wv)            # two layers deep: padding must stay zero
        assert float(y[..., W:].abs().max()) == 0 and float(y[..., :W].abs().max()) > 0
        p1, r1 = ref(L, R, disp)
        msk = (disp > 10).float()
        a1 = att.logits_cat(L, ops.attn_pack(None, disp, disp, msk, disp))
        a2 = att.logits(ops.attn_pack(L, disp, disp, msk, disp))
        d64 = lambda t: t.cpu().double()
        Pr = {f"rf.{k}": d64(v.detach()) for k, v in ref.state_dict().items()}
        _, r0 = og.refinement(d64(L), d64(R), d64(disp), Pr, "rf", 3)
        Pa = {f"sa.{k}": d64(v.detach()) for k, v in att.state_dict().items()}
        x = torch.cat((d64(L), d64(disp)[:, None], d64(disp)[:, None], d64(msk)[:, None], -d64(disp)[:, None]), 1)
        for i in range(3):
            x = og.conv_bn(x, Pa, f"sa.conv.{i}", relu=i < 2)
        r0, a0 = r0.float().cuda(), x.float().cuda()
    assert r1.shape == r0.shape == (B, H, W) and a1.shape == a0.shape == a2.shape
    assert (r1 - r0).abs().max().item() <= 1e-4 * max(1.0, r0.abs().max().item())
    assert (a1 - a0).abs().max().item() <= 1e-4 * max(1.0, a0.abs().max().item())
    assert (a2 - a0).abs().max().item() <= 1e-4 * max(1.0, a0.abs().max().item())
