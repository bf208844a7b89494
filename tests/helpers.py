"""Shared synthetic-input generators for the parity tests (seeded)."""
from __future__ import annotations

import torch


def make_feats(B, C, H, W, seed=17, scale=0.3, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    L = (torch.randn(B, C, H, W, generator=g) * scale).to(device)
    R = (torch.randn(B, C, H, W, generator=g) * scale).to(device)
    return L.contiguous(), R.contiguous()


def make_masks(B, H, W, rho_l=0.1, rho_r=0.1, seed=18, device="cpu", clustered=False):
    g = torch.Generator().manual_seed(seed)
    if clustered:
        # dilated edges of a smooth random field: masks hug structures like real detail masks
        f = torch.randn(B, 1, max(H // 8, 1) + 2, max(W // 8, 1) + 2, generator=g)
        f = torch.nn.functional.interpolate(f, size=(H, W), mode="bilinear", align_corners=False)[:, 0]
        gx = (f[:, :, 1:] - f[:, :, :-1]).abs()
        gx = torch.nn.functional.pad(gx, (0, 1))
        thr_l = torch.quantile(gx.flatten(), 1 - rho_l)
        ml = (gx > thr_l).float()
        shift = 3
        mr = torch.roll(ml, -shift, dims=2)
    else:
        ml = (torch.rand(B, H, W, generator=g) < rho_l).float()
        mr = (torch.rand(B, H, W, generator=g) < rho_r).float()
    return ml.to(device).contiguous(), mr.to(device).contiguous()


# (name, B, C, H, W, D): reference level shapes (SURVEY.md section 8 header) and odd ones
LEVEL_SHAPES = [
    ("sceneflow_s1", 1, 72, 60, 108, 24),
    ("sceneflow_s2", 1, 24, 180, 324, 72),
    ("sceneflow_s3_band", 1, 8, 64, 972, 216),
    ("kitti_s1", 2, 72, 42, 141, 24),      # W % 4 != 0 -> scalar staging path
    ("kitti_s3_band", 1, 8, 16, 1269, 216),
    ("tiny_ragged", 3, 5, 7, 33, 9),
    ("single_column", 1, 4, 3, 1, 5),
    ("d_exceeds_w", 1, 6, 4, 20, 64),
]
