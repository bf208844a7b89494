"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Runs the reference model's own forward (modules/SparseDenseNetRefinementMask.py:102-236,
is_check=True / is_eval=True taps) on CPU with
  * hot-path weights from decnet_b200.params.make_hotpath_state(seed) (reproducible anywhere),
  * synthetic feature pyramids from decnet_b200.params.make_features(seed) injected in place of
    the feature extractor (out of scope),
  * the reference's own SpaMat/SpaVar Python wrappers on top of the CPU oracle extension
    (the CUDA kernels cannot run here; they are pinned separately on the GPU box via oracle/_ref).
Only OUTPUTS (and seeds) are stored; inputs and weights are regenerated from the seeds.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from decnet_b200.params import make_features, make_hotpath_state  # noqa: E402
from oracle import ref_loader  # noqa: E402


class FixedFeatures(torch.nn.Module):
    """Stands in for feature_extractor: first call -> left pyramid, second -> right."""

    def __init__(self, left, right):
        super().__init__()
        self.seq = [left, right]
        self.calls = 0

    def forward(self, x):
        out = self.seq[self.calls % 2]
        self.calls += 1
        return out


def run_case(name, B, H, W, max_disp, seed, use_detail, thold, skip_stage_id=4, mask_rho=0.25):
    torch.manual_seed(seed)
    torch.set_num_threads(8)
    model = ref_loader.build_reference_model(max_disp=max_disp, use_detail=use_detail, thold=thold,
                                             skip_stage_id=skip_stage_id)
    sd = make_hotpath_state(seed)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("feature_extractor") for k in missing), [k for k in missing if not k.startswith("feature_extractor")][:5]
    left, right = make_features(B, H, W, seed=seed)
    model.feature_extractor = FixedFeatures(left, right)
    taps = {"var": [], "lmask": [], "rmask": [], "vol": [], "ldetail": [], "rdetail": []}
    for m in model.sparse_var:
        m.register_forward_hook(lambda mod, args, out: taps["var"].append(out.detach().clone()))
    for m in model.sparse_matching:
        m.register_forward_pre_hook(lambda mod, args: (taps["lmask"].append(args[2].detach().clone()),
                                                       taps["rmask"].append(args[3].detach().clone())) and None)
    model.get_cost_volume.register_forward_hook(lambda mod, args, out: taps["vol"].append(out.detach().clone()))
    g = torch.Generator().manual_seed(seed + 1)
    lmasks = [(torch.rand(B, H // f, W // f, generator=g) < mask_rho).float() for f in (9, 3, 1)]
    rmasks = [(torch.rand(B, H // f, W // f, generator=g) < mask_rho).float() for f in (9, 3, 1)]
    img = torch.zeros(B, 3, H, W)
    with torch.no_grad():
        (pred_list, dense_list, sparse_list, fusion_list, residual_list, _, _, soft_list,
         _, _, cost) = model(img, img, None, lmasks, rmasks, is_check=True, is_eval=True)
    out = {"meta": np.array([B, H, W, max_disp, seed, int(use_detail), skip_stage_id], dtype=np.int64),
           "thold": np.float64(thold), "mask_rho": np.float64(mask_rho),
           "cost": cost.numpy(), "vol": taps["vol"][0].numpy()}
    for k, lst in (("pred", pred_list), ("dense", dense_list), ("sparse", sparse_list), ("fusion", fusion_list),
                   ("residual", residual_list), ("soft_mask", soft_list), ("var", taps["var"]),
                   ("lmask", taps["lmask"]), ("rmask", taps["rmask"])):
        for i, t in enumerate(lst):
            out[f"{k}{i}"] = t.detach().numpy().astype(np.float32)
    path = Path(__file__).resolve().parent / f"{name}.npz"
    np.savez_compressed(path, **out)
    print(name, "->", path, f"{path.stat().st_size / 1e3:.0f} kB;",
          "mask density:", [round(float(m.mean()), 3) for m in taps["lmask"]])


if __name__ == "__main__":
    # learned-detector masks (shipped use_detail=1, thold 0.9 -> demo.sh:1); random-init densities are erratic
    run_case("pipeline_detail", B=2, H=108, W=162, max_disp=216, seed=17, use_detail=True, thold=0.9)
    # given masks (use_detail=False), the demo.py:161-162 path
    run_case("pipeline_masks", B=1, H=108, W=189, max_disp=216, seed=23, use_detail=False, thold=0.9)
    # Middlebury-style: finest stage skipped -> bicubic (demo.sh:5 skip_stage_id=3)
    run_case("pipeline_skip3", B=1, H=81, W=108, max_disp=243, seed=29, use_detail=True, thold=0.5, skip_stage_id=3)
