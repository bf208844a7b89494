"""GPU, BASELINE.json's full sizes: the oracle is too slow there, so parity is checked through
size-independent properties of the sparse ops and of the pipeline (SceneFlow B=8, KITTI, Middlebury)."""
import pytest
import torch

from helpers import make_feats, make_masks

pytestmark = pytest.mark.gpu


def _fused(L, R, ml, mr, D):
    from decnet_b200 import ops
    return ops.spamat_spavar_forward(L, R, ml, mr, D)


@pytest.mark.parametrize("name,B,C,H,W,D", [("sceneflow_s3", 8, 8, 540, 972, 216), ("sceneflow_s2", 8, 24, 180, 324, 72),
                                             ("kitti_s3", 4, 8, 378, 1269, 216), ("middlebury_s2", 1, 24, 675, 972, 261)])
def test_sparse_properties_full_size(name, B, C, H, W, D):
    from decnet_b200 import ops
    L, R = make_feats(B, C, H, W, device="cuda")
    ml, mr = make_masks(B, H, W, 0.1, 0.1, device="cuda", clustered=True)
    out, var, ssim, mx = _fused(L, R, ml, mr, D)
    unmasked = ml == 0
    # 1. unmasked pixels are exactly zero in every output
    for t in (out, var, ssim, mx):
        assert t[unmasked].abs().max().item() == 0
    # 2. ranges: 0 <= out <= D-1 (+eps), var >= 0, sum_sim >= 1e-6 (>= 1 + 1e-6 when the max is a real candidate), max >= 1e-6
    m = ~unmasked
    assert out[m].min().item() >= 0 and out[m].max().item() <= D - 1 + 1e-3
    assert var[m].min().item() >= 0 and ssim[m].min().item() >= 1e-6 * (1 - 1e-6) and mx[m].min().item() >= 1e-6 * (1 - 1e-6)
    # 3. candidate counts: masked pixels without a candidate give out = var = 1, sum_sim = max = 1e-6
    cnt, _ = ops.candidate_signature(ml, mr, D)
    none = m & (cnt == 0)
    if none.any():
        assert torch.all(out[none] == 1.0) and torch.all(var[none] == 1.0)
        assert torch.allclose(ssim[none], torch.full_like(ssim[none], 1e-6)) and torch.allclose(mx[none], torch.full_like(mx[none], 1e-6))
    # 4. fused == separate ops (same kernels, same arithmetic): bit-exact
    o2, s2, m2 = ops.spamat_forward(L, R, ml, mr, D)
    v2, s3, m3 = ops.spavar_forward(L, R, ml, mr, o2, D)
    assert torch.equal(o2, out) and torch.equal(s2, ssim) and torch.equal(m2, mx)
    assert torch.equal(s3, ssim) and torch.equal(m3, mx) and torch.allclose(v2, var, rtol=1e-6, atol=1e-6)
    # 5. batch permutation equivariance (rows are independent work items)
    perm = torch.randperm(B, device="cuda")
    op, vp, sp, mp = _fused(L[perm].contiguous(), R[perm].contiguous(), ml[perm].contiguous(), mr[perm].contiguous(), D)
    assert torch.equal(op, out[perm]) and torch.equal(vp, var[perm]) and torch.equal(mp, mx[perm])
    # 6. softmax shift invariance: scaling R by 0 makes every cost 0 -> uniform weights over the candidates,
    #    so out = mean of the candidate disparities and max_cost = the 1e-6 floor
    oz, vz, sz, mz = _fused(L, torch.zeros_like(R), ml, mr, D)
    has = m & (cnt > 0)
    assert torch.allclose(mz[has], torch.full_like(mz[has], 1e-6))
    assert torch.allclose(sz[has], cnt[has].float() * torch.exp(torch.tensor(-1e-6)).item() + 1e-6, rtol=1e-5)
    # 7. the two staged load paths (TMA, cp.async) agree exactly where both apply
    if W % 4 == 0:
        from decnet_b200 import _lib
        res = []
        for path in (1, 2):
            _lib.lib().decnet_set_sparse_path(path)
            try:
                res.append(_fused(L, R, ml, mr, D))
            finally:
                _lib.lib().decnet_set_sparse_path(0)
        for x, y in zip(*res):
            assert torch.equal(x, y)
    # 8. every forward kernel (staged rows: one row per CTA / persistent, both load paths; sector-gather: one row
    #    per CTA / software-pipelined) agrees with the default: max_cost bit for bit (same FMA chain), the sums to rounding (lanes per pixel differ)
    from decnet_b200 import _lib
    combos = [(1, 1), (1, 2), (0, 3), (0, 4)] + ([(2, 1), (2, 2)] if W % 4 == 0 else [])
    for path, variant in combos:
        _lib.lib().decnet_set_sparse_path(path)
        _lib.lib().decnet_set_sparse_variant(variant)
        try:
            o3, v3, s3, m3 = _fused(L, R, ml, mr, D)
            # a shape whose staged rows leave no room for the mask staging falls back to one row per CTA
            assert _lib.lib().decnet_last_sparse_variant() in (variant, 1 if variant == 2 else variant)
        finally:
            _lib.lib().decnet_set_sparse_path(0)
            _lib.lib().decnet_set_sparse_variant(0)
        assert torch.equal(m3, mx)
        assert torch.allclose(o3, out, rtol=1e-6, atol=1e-5) and torch.allclose(s3, ssim, rtol=1e-6, atol=1e-7)
        assert torch.allclose(v3, var, rtol=1e-5, atol=1e-4)
        assert torch.equal(o3 == 0, out == 0)


def test_sparse_translation_property():
    """Shifting both views and both masks by k columns (zero fill) shifts the outputs by k columns:
    candidates are relative (w - d), so nothing else may change away from the left border."""
    B, C, H, W, D, k = 2, 8, 64, 972, 216, 40
    L, R = make_feats(B, C, H, W, device="cuda")
    ml, mr = make_masks(B, H, W, 0.1, 0.1, device="cuda")
    out, var, ssim, mx = _fused(L, R, ml, mr, D)

    def shift(t):
        s = torch.zeros_like(t)
        s[..., k:] = t[..., :-k]
        return s.contiguous()
    o2, v2, s2, m2 = _fused(shift(L), shift(R), shift(ml), shift(mr), D)
    assert torch.equal(o2[..., k:], out[..., :-k]) and torch.equal(v2[..., k:], var[..., :-k])
    assert torch.equal(m2[..., k:], mx[..., :-k]) and torch.equal(s2[..., k:], ssim[..., :-k])


@pytest.mark.parametrize("workload,batch", [("sceneflow", 8), ("kitti", 4), ("middlebury", 1)])
def test_pipeline_full_size_runs_and_is_deterministic(workload, batch):
    from decnet_b200.synthetic import build_workload
    model, left, right, info = build_workload(workload, batch, rho=0.1)
    a = model(left, right)[0]
    b = model(left, right)[0]
    assert a.shape == (batch, info["H"], info["W"]) and torch.isfinite(a).all()
    assert torch.equal(a, b)                                   # no atomics / races on the path
    dens = info["left_mask_density"]
    assert all(abs(d - 0.1) < 0.02 for d in dens), dens


@pytest.mark.parametrize("workload,batch", [("sceneflow", 2), ("kitti", 2)])
def test_two_stream_step_equals_single_stream(workload, batch):
    """Masks + sparse ops on the forked second stream (DecompMatching.overlap, the default) change the order of
    execution only: same bits as the single-stream order, eagerly and replayed from a captured graph."""
    from decnet_b200.synthetic import build_workload
    model, left, right, info = build_workload(workload, batch, rho=0.1)
    model.overlap = False
    want = model(left, right)[0].clone()
    model.overlap = True
    got = model(left, right)[0]
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        model(left, right)
        with torch.cuda.graph(graph, stream=side):
            out = model(left, right)[0]
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)


def test_middlebury_bands_match_single_device():
    """BASELINE.json configs[3] at full size: 8 row bands (simulated in one process) vs one device."""
    from decnet_b200 import bands
    from decnet_b200.synthetic import build_workload
    model, left, right, info = build_workload("middlebury", 1, rho=0.1)
    want = model(left, right)[0]
    full = bands.forward_bands(model, left, right, bands.LocalTransport(8))
    got = full[0]
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 1e-3 + 5e-3 * scale
    assert float((got - want).abs().mean()) <= 1e-3 + 1e-4 * scale
