"""CPU: host-side logic of the drop-in units that needs no GPU -- BN folding, the weight caches keyed on the identity of the
parameters (reload / in-place edit / move must invalidate them whichever parent module was used), the 3xTF32 operand split,
the inference-only guard and the single-route rule (no CPU / library fallback for hot-path layers)."""
import pytest
import torch


def _randomise(unit, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in unit.parameters():
            p.copy_(torch.randn(p.shape, generator=g))
        for n, b in unit.named_buffers():
            if n.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)
            elif n.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)


def _ref_fold(unit):
    w, b = unit.conv.weight.detach(), torch.zeros(unit.conv.out_channels)
    s = unit.bn.weight.detach() / torch.sqrt(unit.bn.running_var + 1e-5)
    return w * s.view(-1, 1, 1, 1), (b - unit.bn.running_mean) * s + unit.bn.bias.detach()


def test_folded_weights_follow_reloads_through_a_foreign_parent():
    """our units swapped into a parent that is NOT DecompMatching; the parent's
    load_state_dict / in-place edits / .to() never reach any override of ours, yet the folded weights must follow."""
    from decnet_b200.model import Conv2dUnit, Refinement
    parent = torch.nn.Module()
    parent.refinement = torch.nn.ModuleList([Refinement(8, stage_id=3)])
    parent.eval()
    u = parent.refinement[0].conv[1]
    _randomise(parent, 1)
    w1, b1 = u.folded()
    assert u.folded()[0] is w1                                     # cached while nothing changes
    rw, rb = _ref_fold(u)
    assert torch.allclose(w1, rw) and torch.allclose(b1, rb)
    other = torch.nn.Module()
    other.refinement = torch.nn.ModuleList([Refinement(8, stage_id=3)])
    _randomise(other, 2)
    parent.load_state_dict(other.state_dict())                     # the parent's own recursive loader
    w2, b2 = u.folded()
    rw, rb = _ref_fold(u)
    assert not torch.equal(w1, w2) and torch.allclose(w2, rw) and torch.allclose(b2, rb)
    with torch.no_grad():
        u.bn.running_var.mul_(4.0)                                 # in-place edit of a BN statistic
    w3, _ = u.folded()
    assert torch.allclose(w3, _ref_fold(u)[0]) and not torch.equal(w3, w2)
    parent.double().float()                                        # new storages (what .to(device) does)
    assert torch.allclose(u.folded()[0], w3) and u.folded()[0] is not w3


def test_units_are_inference_only_and_have_no_cpu_route():
    from decnet_b200 import _lib
    from decnet_b200.model import Conv2dUnit, DecompMatching, set_precision
    u = Conv2dUnit(8, 8, 3, padding=1)
    u.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        u.folded()
    u.eval()
    with pytest.raises(_lib.DecnetError):
        u(torch.zeros(1, 8, 16, 16))                               # CPU tensor: no fallback
    m = DecompMatching()
    assert not m.training and m.precision == "fp32"
    assert all(c.precision == "fp32" for c in m.modules() if isinstance(c, Conv2dUnit))
    m.set_precision("tf32")
    assert all(c.precision == "tf32" for c in m.modules() if isinstance(c, Conv2dUnit))
    with pytest.raises(ValueError):
        set_precision(m, "bf16")
    assert not any(hasattr(c, "library_ok") for c in m.modules() if isinstance(c, Conv2dUnit))   # no library branch exists


def test_3xtf32_operand_split_is_fp32_class():
    """hi + lo carries 22 of the 24 significand bits, both parts are TF32 values, and hi*hi' + hi*lo' + lo*hi' reproduces
    the fp32 product to ~2^-21 -- the arithmetic of the kernels' split mode restated in fp64."""
    from decnet_b200.ops import rna_tf32, split_tf32
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1 << 16, generator=g) * torch.exp(torch.randn(1 << 16, generator=g) * 3)
    w = torch.randn(1 << 16, generator=g)
    xh, xl = split_tf32(x)
    wh, wl = split_tf32(w)
    for t in (xh, xl, wh, wl):
        assert torch.equal(t, rna_tf32(t)) or torch.equal(t.view(torch.int32) & 0x1FFF, torch.zeros_like(t, dtype=torch.int32))
    assert float(((xh.double() + xl.double() - x.double()).abs() / x.double().abs()).max()) <= 2.0 ** -21
    exact = x.double() * w.double()
    three = xh.double() * wh.double() + xh.double() * wl.double() + xl.double() * wh.double()
    one = xh.double() * wh.double()
    assert float(((three - exact).abs() / exact.abs()).max()) <= 2.0 ** -20
    assert float(((one - exact).abs() / exact.abs()).max()) >= 2.0 ** -13           # plain TF32 for comparison
