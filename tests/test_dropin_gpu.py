"""GPU: drop-in proof on hardware.  The UNMODIFIED reference model class (modules/SparseDenseNetRefinementMask.py, from
/root/reference in the build container or from the staged copy baseline/_ref on the GPU box) runs on a B200 on top of
libdecnet_b200.so at the two integration levels of INTEGRATION.md:

  ext      section 1: the reference's own functions/SpaMat.py / SpaVar.py call `decnet_b200.ext` through the pybind
           modules' call surface (modules/SparseMatching/functions/SpaMat.py:4,24-28); everything else is the reference's
           PyTorch / cuDNN code (fp32, TF32 off);
  modules  section 2: the classes the model file imports by name (SparseDenseNetRefinementMask.py:9-12) are replaced by
           ours before construction -- the reference's stage loop then drives our kernels.

Both are compared with tests/golden/pipeline_images.npz (the same model run entirely on the CPU in the build container) and
with the pipeline-level route (our extractor + DecompMatching)."""
import contextlib
import io
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests" / "golden"))

from golden_util import chain_close  # noqa: E402
from make_golden_images import make_images  # noqa: E402
from oracle import ref_loader  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="reference tree not staged (baseline/_ref)")]

SWAPPED = ("GetCostVolume", "CostRegNetNoDown", "disparity_regression", "DynamicUpsampling", "SoftAttention", "Refinement",
           "GenerateSparseMask", "FeatExtNetChannelPlus", "SpaMat", "SpaVar")


def _state(seed):
    from decnet_b200.params import make_featext_state, make_hotpath_state
    sd = {f"feature_extractor.{k}": v for k, v in make_featext_state(seed).items()}
    sd.update(make_hotpath_state(seed))
    return sd


def _reference_model(level, seed, max_disp):
    import decnet_b200
    from decnet_b200 import ext, features, model as dm
    ref_loader.install(ext.SpaMat, ext.SpaVar)
    # `from ..build.lib import SpaMat` binds at first import; rebind in case another test imported the package before
    sys.modules["modules.SparseMatching.functions.SpaMat"].SpaMat = ext.SpaMat
    sys.modules["modules.SparseVar.functions.SpaVar"].SpaVar = ext.SpaVar
    # NOT `import modules.SparseDenseNetRefinementMask as M`: modules/__init__.py:3 re-exports the class under the
    # sub-module's name, so that attribute is the class; the module object lives in sys.modules
    M = sys.modules["modules.SparseDenseNetRefinementMask"]
    saved = {n: getattr(M, n) for n in SWAPPED}
    if level == "modules":                                     # INTEGRATION.md section 2, verbatim
        M.SpaMat, M.SpaVar = decnet_b200.SpaMat, decnet_b200.SpaVar
        M.GetCostVolume, M.CostRegNetNoDown, M.disparity_regression = dm.GetCostVolume, dm.CostRegNetNoDown, dm.disparity_regression
        M.DynamicUpsampling, M.SoftAttention, M.Refinement, M.GenerateSparseMask = (dm.DynamicUpsampling, dm.SoftAttention,
                                                                                    dm.Refinement, dm.GenerateSparseMask)
        M.FeatExtNetChannelPlus = features.FeatExtNetChannelPlus
    try:
        model = ref_loader.build_reference_model(max_disp=max_disp, use_detail=True, thold=0.9)
    finally:
        for n, v in saved.items():
            setattr(M, n, v)
    missing, unexpected = model.load_state_dict(_state(seed), strict=False)
    assert not missing and not unexpected, (missing[:5], unexpected[:5])
    return model.cuda().eval()


def _run(model, left, right):
    masks = {"l": [], "r": []}
    hooks = [m.register_forward_pre_hook(lambda mod, args: (masks["l"].append(args[2].detach().clone()),
                                                            masks["r"].append(args[3].detach().clone())) and None)
             for m in model.sparse_matching]
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False     # the reference's own layers in fp32
    try:
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            dummy = [torch.zeros(1, device="cuda")] * 3
            out = model(left, right, None, dummy, dummy, is_check=True, is_eval=True)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        for h in hooks:
            h.remove()
    return out[0], masks


def _golden():
    z = np.load(ROOT / "tests" / "golden" / "pipeline_images.npz")
    seed, B, H, W, max_disp = (int(v) for v in z["meta"])
    gmask = []
    for i in range(3):
        shp = tuple(int(v) for v in z[f"lshape{i}"])
        gmask.append(torch.from_numpy(np.unpackbits(z[f"lmask{i}"])[: int(np.prod(shp))].reshape(shp).astype(np.float32)).cuda())
    return z, seed, B, H, W, max_disp, gmask


@pytest.mark.parametrize("level", ["ext", "modules"])
def test_unmodified_reference_model_runs_on_our_library(level):
    from decnet_b200 import _lib
    from decnet_b200.features import FeatExtNetChannelPlus
    from decnet_b200.model import DecompMatching
    from decnet_b200.params import make_featext_state, make_hotpath_state
    z, seed, B, H, W, max_disp, gmask = _golden()
    left, right = (t.cuda() for t in make_images(seed, B, H, W))
    model = _reference_model(level, seed, max_disp)
    assert type(model).__name__ == "SparseDenseNetRefinementMask" and type(model).__module__ == "modules.SparseDenseNetRefinementMask"
    _lib.lib().decnet_reset_launch_count()
    preds, masks = _run(model, left, right)
    torch.cuda.synchronize()
    launches = int(_lib.lib().decnet_launch_count())
    assert launches >= (6 if level == "ext" else 60), launches          # our kernels did the work (SpaMat+SpaVar x 3 levels at least)
    # against the reference run entirely on the CPU (golden): coarse stage, masks, every later stage
    want0 = torch.from_numpy(z["pred0"]).cuda()
    assert float((preds[0] - want0).abs().mean()) <= (1e-3 if level == "ext" else 0.05)     # modules: bf16 aggregation budget
    for i in range(3):
        agree = float((masks["l"][i] == gmask[i]).float().mean())
        assert agree >= 0.999, (i, agree)
    for i in range(1, 4):
        want = torch.from_numpy(z[f"pred{i}"]).cuda()
        diff = (preds[i] - want).abs()
        scale = max(1.0, float(want.abs().max()))
        if level == "ext":
            assert chain_close(preds[i], want, rel=1e-2, abs_=5e-2), (i, float(diff.max()), scale)
        assert float(diff.mean()) <= (5e-3 if level == "ext" else 2e-2) * scale, (i, float(diff.mean()), scale)
        assert float(diff.median()) <= (5e-4 if level == "ext" else 1e-2) * scale, (i, float(diff.median()), scale)
    # against the pipeline-level route: our extractor + DecompMatching, same weights
    fe = FeatExtNetChannelPlus(8)
    fe.load_state_dict(make_featext_state(seed), strict=True)
    hot = DecompMatching(max_disp=max_disp, use_detail=True, thold=0.9)
    hot.load_state_dict(make_hotpath_state(seed))
    fe, hot = fe.cuda(), hot.cuda()
    ours, taps = hot(fe(left), fe(right), is_check=True)
    for i in range(3):
        assert float((taps["left_mask"][i] == masks["l"][i]).float().mean()) >= 0.999, i
    d = (ours - preds[-1]).abs()
    scale = max(1.0, float(preds[-1].abs().max()))
    if level == "modules":
        # same kernels under the reference's stage loop (separate SpaMat / SpaVar calls, materialised concatenations)
        assert float(d.max()) <= 1e-3 + 5e-3 * scale and float(d.mean()) <= 1e-3 + 1e-4 * scale, (float(d.max()), float(d.mean()), scale)
    else:
        assert float(d.mean()) <= 2e-2 * scale, (float(d.mean()), scale)


def test_checkpoint_reload_through_the_reference_parent_takes_effect():
    """our units inside the reference's model class; a second load_state_dict after a forward (moving from one
    checkpoint to another) goes through the PARENT's loader and must change the result (caches keyed on parameter identity)."""
    z, seed, B, H, W, max_disp, _ = _golden()
    left, right = (t.cuda() for t in make_images(seed, B, H, W))
    model = _reference_model("modules", seed, max_disp)
    a, _ = _run(model, left, right)
    model.load_state_dict(_state(seed + 1), strict=False)
    b, _ = _run(model, left, right)
    fresh = _reference_model("modules", seed + 1, max_disp)
    c, _ = _run(fresh, left, right)
    assert not torch.allclose(a[-1], b[-1])
    assert torch.equal(b[-1], c[-1])


def test_reference_test_loss_func_pins_epe_kernel():
    """modules/loss.py:427-437 `test_loss_func` (it calls .cuda() itself) against decnet_epe_3px and the oracle restatement."""
    from decnet_b200 import ext, ops
    from oracle import codec
    ref_loader.install(ext.SpaMat, ext.SpaVar)
    from modules.loss import test_loss_func
    g = torch.Generator(device="cuda").manual_seed(2)
    pred = torch.rand(2, 108, 162, device="cuda", generator=g) * 300 - 10
    gt = torch.rand(2, 108, 162, device="cuda", generator=g) * 250 - 20
    r_epe, r_l3 = test_loss_func(pred, gt, 192.0)
    epe, l3 = ops.epe_3px(pred, gt, 192.0)
    o_epe, o_l3 = codec.epe_3px(pred.cpu(), gt.cpu(), 192.0)
    assert abs(float(epe) - float(r_epe)) <= 1e-4 * float(r_epe) and abs(float(l3) - float(r_l3)) <= 1e-3
    assert abs(float(o_epe) - float(r_epe)) <= 1e-4 * float(r_epe) and abs(float(o_l3) - float(r_l3)) <= 1e-3
