"""GPU parity of the feature-extractor drop-in against the golden outputs of
the unmodified reference module, and of its tensor-core route against its fp32 route at SceneFlow size."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests" / "golden"))

from decnet_b200.params import make_featext_state  # noqa: E402
from make_golden_features import make_image  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = ROOT / "tests" / "golden" / "features.npz"


def _model(seed, precision="fp32"):
    from decnet_b200.features import FeatExtNetChannelPlus
    m = FeatExtNetChannelPlus(8, precision=precision)
    m.load_state_dict(make_featext_state(seed), strict=True)
    return m.cuda()


@pytest.mark.parametrize("tf32", [False, True])
def test_features_match_reference_golden(tf32):
    """precision "fp32" (3xTF32 tensor-core layers, the default): <= 1e-4 of the map's scale (accumulation order only).
    precision "tf32" (what the reference runs on a GPU by default): 10-bit mantissa operands through up to 14 layers,
    tolerance 1e-2 of the scale (measured ~2e-3)."""
    z = np.load(GOLD)
    seed, B, H, W = (int(v) for v in z["meta"])
    out = _model(seed, "tf32" if tf32 else "fp32")(make_image(seed, B, H, W).cuda())
    tol = 1e-2 if tf32 else 1e-4
    for k in ("stage0", "stage1", "stage2", "stage3"):
        want = torch.from_numpy(z[k]).cuda()
        assert out[k].shape == want.shape
        err = float((out[k] - want).abs().max())
        assert err <= tol * max(1.0, float(want.abs().max())), (k, err)


def test_features_full_size_tensor_core_route_vs_fp32_route():
    """540x972 (SceneFlow padded), B=2: every level of the pyramid, tf32 mode vs the default 3xTF32 mode."""
    from decnet_b200.model import set_precision
    m = _model(5)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(2, 3, 540, 972, device="cuda", generator=g)
    b = m(x)
    set_precision(m, "tf32")
    a = m(x)
    for k, shape in (("stage0", (2, 216, 20, 36)), ("stage1", (2, 72, 60, 108)), ("stage2", (2, 24, 180, 324)),
                     ("stage3", (2, 8, 540, 972))):
        assert tuple(a[k].shape) == shape
        assert float((a[k] - b[k]).abs().max()) <= 1e-2 * max(1.0, float(b[k].abs().max())), k


def test_extract_pair_two_streams_equals_two_calls():
    """Both views with the right one on a forked stream: same bits as two calls on one stream, eagerly and
    replayed from a captured graph (two branches)."""
    from decnet_b200.features import extract_pair
    m = _model(5)
    g = torch.Generator(device="cuda").manual_seed(2)
    xl = torch.randn(2, 3, 108, 216, device="cuda", generator=g)
    xr = torch.randn(2, 3, 108, 216, device="cuda", generator=g)
    wl, wr = m(xl), m(xr)
    fl, fr = extract_pair(m, xl, xr)
    torch.cuda.synchronize()
    for k in wl:
        assert torch.equal(fl[k], wl[k]) and torch.equal(fr[k], wr[k])
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        extract_pair(m, xl, xr)
        with torch.cuda.graph(graph, stream=side):
            gl, gr = extract_pair(m, xl, xr)
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    for k in wl:
        assert torch.equal(gl[k], wl[k]) and torch.equal(gr[k], wr[k])
