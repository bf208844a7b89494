"""Generate tests/golden/pipeline_images.npz: the UNMODIFIED reference model end to end -- its own feature extractor
included -- on a seeded image pair (build container only).  Weights: make_featext_state(seed) + make_hotpath_state(seed),
strictly loaded (every key of the reference model is covered).  Only outputs are stored.

    python tests/golden/make_golden_images.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from decnet_b200.params import make_featext_state, make_hotpath_state  # noqa: E402
from oracle import ref_loader  # noqa: E402

SEED, B, H, W, MAX_DISP = 37, 1, 108, 162, 216


def make_images(seed=SEED, B=B, H=H, W=W):
    g = torch.Generator().manual_seed(seed + 3)
    left = torch.randn(B, 3, H, W, generator=g)
    right = torch.roll(left, shifts=-3, dims=3) + 0.05 * torch.randn(B, 3, H, W, generator=g)     # a rough 3-px shift
    return left, right


if __name__ == "__main__":
    torch.set_num_threads(8)
    model = ref_loader.build_reference_model(max_disp=MAX_DISP, use_detail=True, thold=0.9)
    sd = {f"feature_extractor.{k}": v for k, v in make_featext_state(SEED).items()}
    sd.update(make_hotpath_state(SEED))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing[:5], unexpected[:5])
    left, right = make_images()
    masks = {"l": [], "r": []}
    for m in model.sparse_matching:
        m.register_forward_pre_hook(lambda mod, args: (masks["l"].append(args[2].detach().clone()),
                                                       masks["r"].append(args[3].detach().clone())) and None)
    with torch.no_grad():
        dummy = [torch.zeros(1)] * 3        # the lists are indexed even when the learned detector supplies the masks (:121-122)
        out = model(left, right, None, dummy, dummy, is_check=True, is_eval=True)
    pred_list = out[0]
    res = {"meta": np.array([SEED, B, H, W, MAX_DISP], dtype=np.int64)}
    for i, p in enumerate(pred_list):
        res[f"pred{i}"] = p.numpy().astype(np.float32)
    for i, m in enumerate(masks["l"]):
        res[f"lmask{i}"] = np.packbits(m.numpy().astype(bool))
        res[f"lshape{i}"] = np.array(m.shape, dtype=np.int64)
    path = Path(__file__).resolve().parent / "pipeline_images.npz"
    np.savez_compressed(path, **res)
    print("->", path, path.stat().st_size, "bytes; pred max", [float(p.abs().max()) for p in pred_list],
          "mask density", [round(float(m.mean()), 3) for m in masks["l"]])
