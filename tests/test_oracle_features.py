"""CPU: the feature-extractor restatement (oracle/features.py) against the golden outputs of the unmodified
reference module, and the key contract of the drop-in class."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

from decnet_b200.params import make_featext_state  # noqa: E402
from make_golden_features import make_image  # noqa: E402
from oracle.features import feature_pyramid  # noqa: E402

GOLD = ROOT / "tests" / "golden" / "features.npz"


def test_feature_oracle_matches_reference_golden():
    z = np.load(GOLD)
    seed, B, H, W = (int(v) for v in z["meta"])
    out = feature_pyramid(make_image(seed, B, H, W), make_featext_state(seed))
    for k in ("stage0", "stage1", "stage2", "stage3"):
        want = torch.from_numpy(z[k])
        assert out[k].shape == want.shape
        assert float((out[k] - want).abs().max()) <= 1e-4 * max(1.0, float(want.abs().max())), k


def test_dropin_class_has_the_reference_keys():
    from decnet_b200.features import FeatExtNetChannelPlus
    m = FeatExtNetChannelPlus(8)
    sd = make_featext_state(3)
    res = m.load_state_dict(sd, strict=True)          # same keys as the reference module accepted strictly
    assert not res.missing_keys and not res.unexpected_keys
    assert m.out_channels == [216, 72, 24, 8]
    try:
        m(torch.zeros(1, 3, 27, 27))
    except RuntimeError as e:
        assert "no CPU path" in str(e)
    else:
        raise AssertionError("CPU input must be refused")
