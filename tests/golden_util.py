"""Loading of the committed golden cases (tests/golden/*.npz, made by make_golden.py from the
unmodified reference) and regeneration of their seeded inputs."""
from pathlib import Path

import numpy as np
import torch

from decnet_b200.params import make_features, make_hotpath_state

GOLD = Path(__file__).resolve().parent / "golden"
CASES = ["pipeline_detail", "pipeline_masks", "pipeline_skip3"]


def load_case(name, device="cpu"):
    z = np.load(GOLD / f"{name}.npz")
    B, H, W, max_disp, seed, use_detail, skip = (int(v) for v in z["meta"])
    thold = float(z["thold"])
    P = make_hotpath_state(seed)
    left, right = make_features(B, H, W, seed=seed, device=device)
    g = torch.Generator().manual_seed(seed + 1)
    rho = float(z["mask_rho"])
    lmasks = [(torch.rand(B, H // f, W // f, generator=g) < rho).float().to(device) for f in (9, 3, 1)]
    rmasks = [(torch.rand(B, H // f, W // f, generator=g) < rho).float().to(device) for f in (9, 3, 1)]
    cfg = dict(max_disp=max_disp, use_detail=bool(use_detail), thold=thold, skip_stage_id=skip, B=B, H=H, W=W)
    return z, P, left, right, lmasks, rmasks, cfg


def gold_list(z, key, device="cpu"):
    out, i = [], 0
    while f"{key}{i}" in z:
        out.append(torch.from_numpy(z[f"{key}{i}"]).to(device)); i += 1
    return out


def chain_close(got, want, rel=2e-4, abs_=1e-3):
    """Tolerance for values that went through several random-init stages: fp32 noise is amplified
    with the magnitude of the activations (hundreds of px at random init), so scale with it."""
    tol = abs_ + rel * float(want.abs().max())
    return float((got - want).abs().max()) <= tol
