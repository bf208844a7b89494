"""Seeded random-init parameters for the hot-path modules, under the reference's state_dict
key names.  The pretrained checkpoints are not available offline, so
parity runs use these on both sides.  Init rule restated from
modules/SparseDenseNetRefinementMask.py:239-257 (conv weights N(0, sqrt(2 / (k*k[*k]*C_out))),
BN gamma 1 / beta 0); BN running statistics (and optionally gamma/beta) are RANDOMISED here
because the defaults (mean 0, var 1) would hide BN-folding bugs.
"""
from __future__ import annotations

import math

import torch

LEVEL_CHANNELS = (216, 72, 24, 8)   # stage0..3 for base_channels = 8 (submodule.py:261-309)


def _conv(g, shape):
    cout = shape[0]
    k = 1
    for s in shape[2:]:
        k *= s
    return torch.randn(shape, generator=g) * math.sqrt(2.0 / (k * cout))


def _bn(sd, prefix, n, g, randomize):
    if randomize:
        sd[prefix + ".weight"] = 1.0 + 0.2 * torch.randn(n, generator=g)
        sd[prefix + ".bias"] = 0.1 * torch.randn(n, generator=g)
        sd[prefix + ".running_mean"] = 0.05 * torch.randn(n, generator=g)
        sd[prefix + ".running_var"] = 0.5 + torch.rand(n, generator=g)
    else:
        sd[prefix + ".weight"] = torch.ones(n)
        sd[prefix + ".bias"] = torch.zeros(n)
        sd[prefix + ".running_mean"] = torch.zeros(n)
        sd[prefix + ".running_var"] = torch.ones(n)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def make_hotpath_state(seed: int = 17, base_channels: int = 8, randomize_bn: bool = True,
                       channels=None) -> dict:
    """state_dict (CPU fp32) for cost_regularizer / detail_detection / dynamic_upsampling /
    soft_attention / refinement with the reference's keys and shapes."""
    g = torch.Generator().manual_seed(seed)
    ch = tuple(channels) if channels is not None else tuple(c * base_channels // 8 for c in LEVEL_CHANNELS)
    sd = {}
    c0 = ch[0]
    # cost_regularizer (submodule.py:608-662): 7x Conv3d c0->c0 + one c0->1, all with BN
    for blk, n in (("conv0", 2), ("conv1", 3), ("conv2", 3)):
        for i in range(n):
            cout = 1 if (blk == "conv2" and i == 2) else c0
            p = f"cost_regularizer.{blk}.{i}"
            sd[p + ".conv.weight"] = _conv(g, (cout, c0, 3, 3, 3))
            _bn(sd, p + ".bn", cout, g, randomize_bn)
    for l in range(3):
        c, cprev = ch[l + 1], ch[l]
        # detail_detection (submodule.py:347-364)
        p = f"detail_detection.{l}"
        sd[p + ".deconv.0.conv.weight"] = _conv(g, (8, cprev, 3, 3)).transpose(0, 1).contiguous()  # ConvTranspose2d: (in, out, k, k)
        sd[p + ".deconv.0.conv.bias"] = 0.01 * torch.randn(8, generator=g)
        sd[p + ".deconv.1.conv.weight"] = _conv(g, (3, 8, 3, 3)); _bn(sd, p + ".deconv.1.bn", 3, g, randomize_bn)
        sd[p + ".conv_sub.0.conv.weight"] = _conv(g, (8, c, 3, 3))
        sd[p + ".conv_sub.0.conv.bias"] = 0.01 * torch.randn(8, generator=g)
        sd[p + ".conv_sub.1.conv.weight"] = _conv(g, (3, 8, 3, 3)); _bn(sd, p + ".conv_sub.1.bn", 3, g, randomize_bn)
        sd[p + ".conv.0.conv.weight"] = _conv(g, (3, 3, 3, 3)); _bn(sd, p + ".conv.0.bn", 3, g, randomize_bn)
        sd[p + ".conv.1.conv.weight"] = _conv(g, (1, 3, 1, 1)); _bn(sd, p + ".conv.1.bn", 1, g, randomize_bn)
    for l in range(3):
        c = ch[l + 1]
        # dynamic_upsampling (submodule.py:566-575)
        p = f"dynamic_upsampling.{l}.weight_learning"
        for i, cin in enumerate((9 * c + 1, 81, 81)):
            sd[f"{p}.{i}.conv.weight"] = _conv(g, (81, cin, 3, 3)); _bn(sd, f"{p}.{i}.bn", 81, g, randomize_bn)
    for l in range(3):
        c = ch[l + 1]
        # soft_attention (submodule.py:593-600)
        p = f"soft_attention.{l}.conv"
        for i, (cin, cout) in enumerate(((c + 4, base_channels), (base_channels, base_channels), (base_channels, 1))):
            sd[f"{p}.{i}.conv.weight"] = _conv(g, (cout, cin, 3, 3)); _bn(sd, f"{p}.{i}.bn", cout, g, randomize_bn)
    for l in range(3):
        c = ch[l + 1]
        # refinement (submodule.py:666-716): in 2c+1 -> c,c,c,c/2,c/2,c/2 -> 1 (bias, no BN)
        p = f"refinement.{l}.conv"
        chain = [(2 * c + 1, c), (c, c), (c, c), (c, c // 2), (c // 2, c // 2), (c // 2, c // 2)]
        for i, (cin, cout) in enumerate(chain):
            sd[f"{p}.{i}.conv.weight"] = _conv(g, (cout, cin, 3, 3)); _bn(sd, f"{p}.{i}.bn", cout, g, randomize_bn)
        sd[f"{p}.6.conv.weight"] = _conv(g, (1, c // 2, 3, 3))
        sd[f"{p}.6.conv.bias"] = 0.01 * torch.randn(1, generator=g)
    return sd


def make_features(B: int, H: int, W: int, seed: int = 17, scale: float = 0.3, channels=LEVEL_CHANNELS,
                  device="cpu"):
    """Synthetic left/right feature pyramids {stage0..3} for a padded H x W image (multiples of 27),
    N(0,1)*scale, NCHW fp32."""
    assert H % 27 == 0 and W % 27 == 0, "pad to multiples of 27 first (demo.py:75-81)"
    g = torch.Generator().manual_seed(seed)
    left, right = {}, {}
    for s, c in enumerate(channels):
        f = 3 ** (3 - s)
        shape = (B, c, H // f, W // f)
        left[f"stage{s}"] = (torch.randn(shape, generator=g) * scale).to(device).contiguous()
        right[f"stage{s}"] = (torch.randn(shape, generator=g) * scale).to(device).contiguous()
    return left, right


# --------------------------------------------------------------------------------------
# feature extractor: keys and shapes of FeatExtNetChannelPlus(8, 4, 3)
# (modules/submodule.py:245-309), listed explicitly so this file does not depend on any module class
# --------------------------------------------------------------------------------------
def featext_layout(base_channels: int = 8):
    """[(key prefix, kind, shape)]: kind 'conv' (Conv2d weight [out,in,k,k] + BN) or 'deconv'
    (ConvTranspose2d weight [in,out,k,k] + BN)."""
    c = base_channels
    c1, c2, c3 = 3 * c, 9 * c, 27 * c
    L = [("conv0.0", "conv", (c, 3, 3, 3)), ("conv0.1", "conv", (c, c, 3, 3)), ("addition_trans0", "conv", (c, c, 1, 1)),
         ("conv1.0", "conv", (c1, c, 3, 3)), ("conv1.1", "conv", (c1, c1, 3, 3)), ("conv1.2", "conv", (c1, c1, 3, 3)),
         ("addition_trans1", "conv", (c1, c1, 1, 1)),
         ("deconv1.deconv", "deconv", (c1, c, 3, 3)), ("deconv1.conv.0", "conv", (c, 2 * c, 3, 3)), ("deconv1.conv.1", "conv", (c, c, 3, 3)),
         ("conv2.0", "conv", (c2, c1, 3, 3)), ("conv2.1", "conv", (c2, c2, 3, 3)), ("conv2.2", "conv", (c2, c2, 3, 3)),
         ("addition_trans2", "conv", (c2, c2, 1, 1)),
         ("deconv2.deconv", "deconv", (c2, c1, 3, 3)), ("deconv2.conv.0", "conv", (c1, 2 * c1, 3, 3)), ("deconv2.conv.1", "conv", (c1, c1, 3, 3)),
         ("conv3_1", "conv", (c3, c2, 3, 3)), ("conv3_2.0", "conv", (c3, c3, 3, 3)), ("conv3_2.1", "conv", (c3, c3, 3, 3)),
         ("addition_ctx_collection.0.stages.c0", "conv", (c3, c3, 1, 1)),
         ("addition_ctx_collection.0.stages.c1", "conv", (c3, c3, 3, 3)),
         ("addition_ctx_collection.0.stages.c2", "conv", (c3, c3, 3, 3)),
         ("addition_ctx_collection.0.stages.c3", "conv", (c3, c3, 3, 3)),
         ("addition_ctx_collection.1", "conv", (c3, 4 * c3, 1, 1)),
         ("addition_fusion", "conv", (c3, 2 * c3, 1, 1)),
         ("deconv3.deconv", "deconv", (c3, c2, 3, 3)), ("deconv3.conv.0", "conv", (c2, 2 * c2, 3, 3)), ("deconv3.conv.1", "conv", (c2, c2, 3, 3))]
    return L


def make_featext_state(seed: int = 17, base_channels: int = 8, randomize_bn: bool = True) -> dict:
    """state_dict (CPU fp32) of the reference's `feature_extractor` sub-module (keys WITHOUT that prefix)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for prefix, kind, shape in featext_layout(base_channels):
        if kind == "conv":
            sd[prefix + ".conv.weight"] = _conv(g, shape)
            n = shape[0]
        else:
            k = shape[2] * shape[3]
            sd[prefix + ".conv.weight"] = torch.randn(shape, generator=g) * math.sqrt(2.0 / (k * shape[1]))
            n = shape[1]
        _bn(sd, prefix + ".bn", n, g, randomize_bn)
    return sd
