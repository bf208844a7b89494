"""CPU: the numpy restatement of the reference's image-space detector (oracle/detail.py) against the masks the
UNMODIFIED reference function produced (tests/golden/detail_masks.npz, made with cv2 by make_golden_detail.py)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

from make_golden_detail import make_detail_image  # noqa: E402
from oracle.detail import detail_masks  # noqa: E402


def gold_masks():
    z = np.load(ROOT / "tests" / "golden" / "detail_masks.npz")
    seed = int(z["meta"][0])
    out = []
    for b in range(2):
        ms = []
        for i in range(3):
            shp = tuple(int(v) for v in z[f"shape{b}_{i}"])
            ms.append(np.unpackbits(z[f"mask{b}_{i}"])[: shp[0] * shp[1]].reshape(shp).astype(bool))
        out.append((seed + b, ms))
    return out


def test_detail_oracle_matches_reference_masks_bit_exactly():
    for seed, want in gold_masks():
        got = detail_masks(make_detail_image(seed))
        for g, w in zip(got, want):
            assert g.shape == w.shape and np.array_equal(g, w)
