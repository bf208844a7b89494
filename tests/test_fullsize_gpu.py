"""GPU, BASELINE.json's full sizes (SceneFlow 540x972 B=8, KITTI 378x1269, Middlebury 2025x2916 D=783): parity against the
CPU oracle at those sizes -- the whole stage loop of one SceneFlow and one KITTI pair, Middlebury's 1/9 and 1/3 levels op by
op, the C/OpenMP sparse oracle on the B=8 finest level and on the C=32/64 sweep shapes -- plus size-independent properties
of the sparse ops and of the pipeline."""
import math

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


# ---------------------------------------------------------------------------------------------------------------
# Full-size parity against the oracle (the CPU restatement of the reference, oracle/): masks bit-equal, sparse / var
# <= 1e-3, coarse disparity <= 0.05 px EPE (bf16 aggregation), fp32-class stages chained from the oracle's own coarse
# disparity at the chained tolerance of test_glue_gpu.py.
# ---------------------------------------------------------------------------------------------------------------
def _cpu(d):
    return {k: v.detach().cpu() for k, v in d.items()}


def _var_close(got, want):
    return torch.allclose(got, want, atol=1e-3, rtol=1e-4)


@pytest.mark.parametrize("workload", ["sceneflow", "kitti"])
def test_pipeline_full_size_vs_oracle(workload):
    """One pair at BASELINE.json's size through oracle.pipeline.forward (learned detectors, thold 0.9, masks calibrated to
    10 %) against DecompMatching on the product route.  KITTI's 1269 / 423 / 141 widths take the pitch-padded route."""
    from decnet_b200.synthetic import build_workload
    from golden_util import chain_close
    from oracle import pipeline as opipe
    model, left, right, info = build_workload(workload, 1, rho=0.1)
    pred, taps = model(left, right, is_check=True)
    P = _cpu(model.state_dict())                               # after the density calibration: what the model really runs
    with torch.no_grad():
        want, otaps = opipe.forward(P, _cpu(left), _cpu(right), info["max_disp"], use_detail=True, thold=0.9,
                                    skip_stage_id=info["skip_stage_id"])
    for l in range(3):
        assert torch.equal(taps["left_mask"][l].cpu(), otaps["left_mask"][l]), ("left mask", l)     # mask selection: bit-exact
        assert torch.equal(taps["right_mask"][l].cpu(), otaps["right_mask"][l]), ("right mask", l)
        # the sparse ops see identical features and identical masks: exact quantities
        assert torch.allclose(taps["sparse"][l].cpu(), otaps["sparse"][l], atol=1e-3, rtol=1e-5), ("sparse", l)
        assert _var_close(taps["var"][l].cpu(), otaps["var"][l]), ("var", l)
        dens = float(otaps["left_mask"][l].mean())
        assert abs(dens - 0.1) < 0.02, dens
    epe0 = float((taps["pred"][0].cpu() - otaps["pred"][0]).abs().mean())
    assert epe0 <= 0.05, f"coarse EPE delta {epe0}"
    # fp32-class stages: chained from the oracle's coarse disparity
    pred_f, taps_f = model(left, right, is_check=True, coarse_pred=otaps["pred"][0].cuda())
    for key in ("dense", "fusion", "residual", "soft_mask", "pred"):
        for i, (g_, w_) in enumerate(zip(taps_f[key][-3:] if key == "pred" else taps_f[key], otaps[key][-3:] if key == "pred" else otaps[key])):
            g_ = g_.cpu()
            assert chain_close(g_, w_, rel=3e-3, abs_=3e-3), f"{key}[{i}] max diff {(g_ - w_).abs().max()} of {w_.abs().max()}"
            assert float((g_ - w_).abs().mean()) <= 1e-3 + 2e-4 * float(w_.abs().max()), (key, i)


def test_middlebury_levels_vs_oracle_teacher_forced():
    """Middlebury 2025x2916, D = 783 (BASELINE.json configs[3]): the 1/9 (225x324, C72, D87) and 1/3 (675x972, C24, D261)
    levels op by op against the oracle, each op fed with OUR upstream values (the CPU oracle's 4 TFLOP coarse aggregation
    at this size is what makes a chained comparison too slow; the coarse stage's parity is size-independent and gated on
    SceneFlow / KITTI above)."""
    from decnet_b200.synthetic import build_workload
    from oracle import glue as og, sparse as osp
    model, left, right, info = build_workload("middlebury", 1, rho=0.1)
    pred, taps = model(left, right, is_check=True)
    P = _cpu(model.state_dict())
    assert len(taps["dense"]) == 2                           # finest level skipped (bicubic)
    for l in range(2):
        s = l + 1
        Lf, Rf = left[f"stage{s}"].cpu(), right[f"stage{s}"].cpu()
        D = info["max_disp"] // 3 ** (3 - s)
        tol = lambda w: 1e-3 + 1e-5 * float(w.abs().max())
        with torch.no_grad():
            # mask selection: identical except where a logit sits ON the threshold within fp32 summation-order noise
            # (1.5 M logits per level here; the oracle's ATen convs and our kernels add in different orders)
            logit_t = math.log(0.9 / 0.1)
            masks = []
            for mine, cur, prev in ((taps["left_mask"][l], Lf, left[f"stage{s - 1}"]), (taps["right_mask"][l], Rf, right[f"stage{s - 1}"])):
                lg = og.detail_logits(cur, prev.cpu(), P, f"detail_detection.{l}")
                want_m = og.threshold_mask(torch.sigmoid(lg), 0.9)
                flips = mine.cpu() != want_m
                assert int(flips.sum()) <= 4, int(flips.sum())
                if flips.any():
                    assert float((lg[flips] - logit_t).abs().max()) <= 2e-5 * max(1.0, float(lg.abs().max())), "a flipped pixel away from the threshold"
                masks.append(mine.cpu())                           # downstream ops are teacher-forced with OUR masks
            lm, rm = masks
            dense = og.dynamic_upsampling(taps["pred"][s - 1].cpu(), Lf, P, f"dynamic_upsampling.{l}")
            assert float((taps["dense"][l].cpu() - dense).abs().max()) <= tol(dense), ("dense", l)
            sp, _, mx = osp.spamat_forward(Lf, Rf, lm, rm, D)
            vr, _, _ = osp.spavar_forward(Lf, Rf, lm, rm, sp, D)
            assert torch.allclose(taps["sparse"][l].cpu(), sp, atol=1e-3, rtol=1e-5), ("sparse", l)
            assert _var_close(taps["var"][l].cpu(), vr), ("var", l)
            m = og.soft_attention(Lf, taps["dense"][l].cpu(), taps["sparse"][l].cpu(), lm, taps["var"][l].cpu(), P, f"soft_attention.{l}")
            assert float((taps["soft_mask"][l].cpu() - m).abs().max()) <= 1e-3, ("soft mask", l)
            p_, r_ = og.refinement(Lf, Rf, taps["fusion"][l].cpu(), P, f"refinement.{l}", s)
            assert float((taps["residual"][l].cpu() - r_).abs().max()) <= tol(p_), ("residual", l)
            assert float((taps["pred"][s].cpu() - p_).abs().max()) <= tol(p_), ("pred", l)
    up = og.bicubic_skip(taps["pred"][2].cpu(), left["stage3"].shape[-2:])
    assert float((pred.cpu() - up).abs().max()) <= 1e-3 + 1e-5 * float(up.abs().max())


@pytest.mark.parametrize("name,B,C,H,W,D,rho", [("sceneflow_s3_b8", 8, 8, 540, 972, 216, 0.1), ("sceneflow_s3_b8_dense", 8, 8, 540, 972, 216, 0.3),
                                                 ("sweep_c32_third", 2, 32, 180, 324, 72, 0.1), ("sweep_c64_third", 2, 64, 180, 324, 72, 0.3),
                                                 ("sweep_c32_full", 1, 32, 540, 972, 216, 0.1), ("sweep_c64_full", 1, 64, 540, 972, 216, 0.03),
                                                 ("kitti_s3_b8", 8, 8, 378, 1269, 216, 0.1), ("middlebury_s2", 1, 24, 675, 972, 261, 0.1)])
def test_sparse_full_size_vs_c_oracle(name, B, C, H, W, D, rho):
    """The C/OpenMP restatement of the reference kernels (oracle/sparse_oracle.c) at BASELINE.json's sizes: configs[1]'s
    finest level at B = 8, the C = 32 / 64 sweep shapes of configs[4], KITTI's and Middlebury's largest sparse levels."""
    from decnet_b200 import ops
    from oracle import sparse as osp
    L, R = make_feats(B, C, H, W, device="cuda")
    ml, mr = make_masks(B, H, W, rho, rho, device="cuda", clustered=(name.endswith("b8")))
    out, var, ssim, mx = ops.spamat_spavar_forward(L, R, ml, mr, D)
    o_out, o_ssim, o_mx = osp.spamat_forward(L, R, ml, mr, D)
    o_var, _, _ = osp.spavar_forward(L, R, ml, mr, o_out, D)
    assert torch.equal(mx.cpu(), o_mx)                                       # same FMA chain: bit-identical
    assert torch.allclose(out.cpu(), o_out, atol=1e-3, rtol=1e-5)
    assert torch.allclose(ssim.cpu(), o_ssim, atol=1e-3, rtol=1e-5)
    assert _var_close(var.cpu(), o_var)
    cnt, hsh = ops.candidate_signature(ml, mr, D)
    o_cnt, o_hsh = osp.candidate_signature(ml, mr, D)
    assert torch.equal(cnt.cpu(), o_cnt) and torch.equal(hsh.cpu(), o_hsh)    # candidate index sets: bit-exact
