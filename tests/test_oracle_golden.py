"""CPU: pin the torch restatement (oracle/dense.py, oracle/glue.py, oracle/pipeline.py) against
the golden vectors produced by the UNMODIFIED reference model (tests/golden/make_golden.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from golden_util import CASES, chain_close, gold_list, load_case
from oracle import dense as odense
from oracle import glue as oglue
from oracle import pipeline as opipe


@pytest.mark.parametrize("name", CASES)
def test_oracle_pipeline_matches_reference_golden(name):
    z, P, left, right, lmasks, rmasks, cfg = load_case(name)
    torch.set_num_threads(8)
    with torch.no_grad():
        pred, taps = opipe.forward(P, left, right, cfg["max_disp"], lmasks, rmasks, cfg["use_detail"],
                                   cfg["thold"], cfg["skip_stage_id"])
    # dense stage: raw volume and regularised cost within 1e-3 abs (north star), usually ~1e-5
    assert torch.allclose(taps["vol"], torch.from_numpy(z["vol"]), atol=1e-4, rtol=1e-4)
    assert torch.allclose(taps["cost"], torch.from_numpy(z["cost"]), atol=1e-3, rtol=1e-4)
    # mask selection is bit-exact
    for got, want in zip(taps["left_mask"], gold_list(z, "lmask")):
        assert torch.equal(got, want)
    for got, want in zip(taps["right_mask"], gold_list(z, "rmask")):
        assert torch.equal(got, want)
    # chained through every stage: fp32 noise grows with the activations (see chain_close)
    for key in ("pred", "dense", "sparse", "fusion", "residual", "soft_mask", "var"):
        want = gold_list(z, key)
        assert len(want) == len(taps[key]), key
        for i, (g_, w_) in enumerate(zip(taps[key], want)):
            assert chain_close(g_, w_), f"{key}[{i}] max diff {(g_ - w_).abs().max()} of {w_.abs().max()}"


@pytest.mark.parametrize("name", CASES)
def test_oracle_ops_teacher_forced(name):
    """Each op fed with the reference's own intermediate (no error accumulation): <= 1e-3 abs."""
    z, P, left, right, lmasks, rmasks, cfg = load_case(name)
    torch.set_num_threads(8)
    D0 = cfg["max_disp"] // 27
    with torch.no_grad():
        vol = torch.from_numpy(z["vol"])
        assert torch.allclose(odense.cost_regularizer(vol, P), torch.from_numpy(z["cost"]), atol=1e-3, rtol=1e-4)
        pred = gold_list(z, "pred")
        assert torch.allclose(odense.disparity_regression(torch.from_numpy(z["cost"]), D0), pred[0], atol=1e-4)
        dense, sparse, var = gold_list(z, "dense"), gold_list(z, "sparse"), gold_list(z, "var")
        soft, fusion, resid = gold_list(z, "soft_mask"), gold_list(z, "fusion"), gold_list(z, "residual")
        lm = gold_list(z, "lmask")
        for l in range(len(dense)):
            s = l + 1
            Lf, Rf = left[f"stage{s}"], right[f"stage{s}"]
            tol = 1e-3 + 1e-5 * float(dense[l].abs().max())
            d = oglue.dynamic_upsampling(pred[s - 1], Lf, P, f"dynamic_upsampling.{l}")
            assert float((d - dense[l]).abs().max()) <= tol
            m = oglue.soft_attention(Lf, dense[l], sparse[l], lm[l], var[l], P, f"soft_attention.{l}")
            assert torch.allclose(m, soft[l], atol=1e-4)
            assert float((oglue.blend(dense[l], sparse[l], soft[l]) - fusion[l]).abs().max()) <= tol
            p_, r_ = oglue.refinement(Lf, Rf, fusion[l], P, f"refinement.{l}", s)
            assert float((r_ - resid[l]).abs().max()) <= tol
            assert float((p_ - pred[s]).abs().max()) <= tol
