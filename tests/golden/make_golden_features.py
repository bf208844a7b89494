"""Generate tests/golden/features.npz from the UNMODIFIED reference feature extractor (build container only).

The reference model is built as in make_golden.py; its `feature_extractor` sub-module
(modules/submodule.py:245-343) gets the seeded weights of decnet_b200.params.make_featext_state
(strict load: the key list is checked against the reference's own) and is run on a seeded image.
Only the four output maps are stored; image and weights are regenerated from the seed.

    python tests/golden/make_golden_features.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from decnet_b200.params import make_featext_state  # noqa: E402
from oracle import ref_loader  # noqa: E402

SEED, B, H, W = 31, 1, 81, 108


def make_image(seed=SEED, B=B, H=H, W=W):
    g = torch.Generator().manual_seed(seed + 7)
    return torch.randn(B, 3, H, W, generator=g)


if __name__ == "__main__":
    torch.set_num_threads(8)
    model = ref_loader.build_reference_model()
    fe = model.feature_extractor
    sd = make_featext_state(SEED)
    want = {k for k in fe.state_dict().keys()}
    assert want == set(sd.keys()), (sorted(want - set(sd.keys()))[:5], sorted(set(sd.keys()) - want)[:5])
    fe.load_state_dict(sd, strict=True)
    fe.eval()
    with torch.no_grad():
        out = fe(make_image())
    path = Path(__file__).resolve().parent / "features.npz"
    np.savez_compressed(path, meta=np.array([SEED, B, H, W], dtype=np.int64),
                        **{k: v.numpy().astype(np.float32) for k, v in out.items()})
    print("->", path, f"{path.stat().st_size / 1e3:.0f} kB", {k: tuple(v.shape) for k, v in out.items()},
          {k: float(v.abs().max()) for k, v in out.items()})
