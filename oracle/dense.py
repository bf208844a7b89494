"""oracle/dense.py -- TEST INFRASTRUCTURE: torch (fp32, CPU or any device) restatement of the
coarse dense stage.

Follows modules/submodule.py:376-390 (candidates 0..D-1), :479-522 (warp by grid_sample with the
align_corners=True-style normalisation but the default align_corners=False sampler, zeroing of the
left operand where w < d, per-channel product), :608-662 + :90-123 (8x Conv3d 3^3 + BN(eval) + ReLU
with one residual; the last layer 216->1 keeps its BatchNorm3d(1), only the ReLU is off) and
:766-777 (soft-argmin).  Convolution / batch-norm arithmetic itself is ATen's (the reference's
own dependency); the bilinear gather is restated explicitly so it does not lean on grid_sample.

Parity pinning: tests/golden/*.npz hold outputs of the UNMODIFIED reference modules run in the
build container (tests/golden/make_golden.py); tests/test_oracle_golden.py checks this file
against them.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _sample_coords(n_out, size, shift, dtype, device):
    """fp32 op sequence of the reference: normalise with (size-1)/2 (submodule.py:497-499), then
    grid_sample's align_corners=False un-normalisation ((g+1)*size-1)/2."""
    pos = torch.arange(n_out, dtype=dtype, device=device)
    g = (pos - shift) / ((size - 1.0) / 2.0) - 1.0
    return ((g + 1.0) * size - 1.0) / 2.0


def bilinear_zero_pad(img, ix, iy):
    """img [B,C,H,W]; ix, iy broadcastable to [B,Ho,Wo] float pixel coords -> [B,C,Ho,Wo].
    4-tap bilinear with zero padding (taps outside the image contribute 0)."""
    B, C, H, W = img.shape
    ix, iy = torch.broadcast_tensors(ix, iy)
    x0 = torch.floor(ix); y0 = torch.floor(iy)
    wx1 = ix - x0; wy1 = iy - y0
    wx0 = 1.0 - wx1; wy0 = 1.0 - wy1
    out = 0
    flat = img.reshape(B, C, H * W)
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xi = (x0 + dx).long(); yi = (y0 + dy).long()
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1))
            idx = idx.expand(B, *idx.shape[-2:]).reshape(B, 1, -1).expand(B, C, -1)
            v = torch.gather(flat, 2, idx).reshape(B, C, *ix.shape[-2:])
            out = out + v * (wx * wy * ok.to(img.dtype)).unsqueeze(1)
    return out


def cost_volume(L, R, D):
    """[B,C,H,W] x2 -> [B,C,D,H,W]; vol = (w>=d ? L : 0) * bilinear0(R, y', x'(w-d))."""
    B, C, H, W = R.shape
    dt, dev = R.dtype, R.device
    iy = _sample_coords(H, H, 0.0, dt, dev).view(1, H, 1)
    w = torch.arange(W, dtype=dt, device=dev)
    vols = []
    for d in range(int(D)):
        ix = _sample_coords(W, W, float(d), dt, dev).view(1, 1, W)
        right = bilinear_zero_pad(R, ix.expand(1, H, W), iy.expand(1, H, W))
        left = L * (w >= d).to(dt).view(1, 1, 1, W)
        vols.append(left * right)
    return torch.stack(vols, dim=2)


def conv3d_unit(x, P, prefix, relu=True):
    x = F.conv3d(x, P[prefix + ".conv.weight"], padding=1)
    x = F.batch_norm(x, P[prefix + ".bn.running_mean"], P[prefix + ".bn.running_var"],
                     P[prefix + ".bn.weight"], P[prefix + ".bn.bias"], False, 0.0, BN_EPS)
    return F.relu(x) if relu else x


def cost_regularizer(vol, P, prefix="cost_regularizer"):
    """[B,C,D,H,W] -> [B,D,H,W]  (submodule.py:650-662)."""
    x = conv3d_unit(vol, P, f"{prefix}.conv0.0")
    o0 = conv3d_unit(x, P, f"{prefix}.conv0.1")
    x = conv3d_unit(o0, P, f"{prefix}.conv1.0")
    x = conv3d_unit(x, P, f"{prefix}.conv1.1")
    x = conv3d_unit(x, P, f"{prefix}.conv1.2") + o0
    x = conv3d_unit(x, P, f"{prefix}.conv2.0")
    x = conv3d_unit(x, P, f"{prefix}.conv2.1")
    x = conv3d_unit(x, P, f"{prefix}.conv2.2", relu=False)
    return x.squeeze(1)


def disparity_regression(cost, D=None):
    """softmax over the candidate axis, expectation of d = 0..D-1 (submodule.py:766-777)."""
    D = cost.shape[1] if D is None else D
    p = torch.softmax(cost, dim=1)
    d = torch.arange(D, dtype=cost.dtype, device=cost.device).view(1, D, 1, 1)
    return (p * d).sum(1)


def dense_stage(L, R, D, P):
    vol = cost_volume(L, R, D)
    cost = cost_regularizer(vol, P)
    return disparity_regression(cost, D), cost, vol
