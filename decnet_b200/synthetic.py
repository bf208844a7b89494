"""Synthetic workloads of the named shapes (BASELINE.json configs)."""
from __future__ import annotations

import math

import torch

from .params import make_features, make_hotpath_state

WORKLOADS = {
    # name: (H_pad, W_pad, max_disp, skip_stage_id)   -- padded to multiples of 27 (demo.py:75-81)
    "sceneflow": (540, 972, 216, 4),       # 540x960 -> 540x972, "192" -> 216
    "kitti": (378, 1269, 216, 4),          # 376x1248 -> 378x1269
    "middlebury": (2025, 2916, 783, 3),    # ~2000x2900, "768" -> 783, finest stage skipped (demo.sh:5)
    "tiny": (108, 162, 216, 4),
}


@torch.no_grad()
def calibrate_mask_density(model, left_feats, right_feats, rho=0.10):
    """Random-init detail detectors give erratic mask densities (0-66 %, SURVEY.md hard part 7).
    Shift the last BatchNorm bias of every detector (one scalar per level) so that the LEFT mask of
    this input has density `rho` at the model's threshold.  Setup-time only, outside any timed region."""
    logit_t = math.log(model.thold / (1.0 - model.thold))
    pre_l = left_feats["stage0"]
    dens = []
    for l in range(model.num_stage - 1):
        if l + 1 >= model.skip_stage_id:
            break
        det = model.detail_detection[l]
        cur = left_feats[f"stage{l + 1}"]
        logits, _, _ = det(cur, pre_l)
        q = torch.quantile(logits.flatten()[:: max(1, logits.numel() // 2_000_000)].float(), 1.0 - rho)
        unit = det.conv[1]
        unit.bn.bias.add_(logit_t - q)                  # in-place: bumps the version the weight caches are keyed on
        logits, _, _ = det(cur, pre_l)
        dens.append(float((torch.sigmoid(logits) > model.thold).float().mean()))
        pre_l = cur
    return dens


def build_workload(name="sceneflow", batch=8, seed=17, device="cuda", rho=0.10, precision="fp32"):
    """(model, left_feats, right_feats, info) for one rank: random-init DecNet hot path + synthetic
    N(0,1)*0.3 feature pyramids of the named shape, mask density calibrated to `rho`."""
    from .model import DecompMatching
    H, W, max_disp, skip = WORKLOADS[name]
    model = DecompMatching(max_disp=max_disp, skip_stage_id=skip, use_detail=True, thold=0.9, precision=precision)
    model.load_state_dict(make_hotpath_state(seed))
    model = model.to(device)
    left, right = make_features(batch, H, W, seed=seed, device=device)
    dens = calibrate_mask_density(model, left, right, rho)
    info = {"H": H, "W": W, "max_disp": max_disp, "skip_stage_id": skip, "batch": batch,
            "left_mask_density": [round(d, 4) for d in dens]}
    return model, left, right, info
