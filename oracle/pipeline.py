"""oracle/pipeline.py -- TEST INFRASTRUCTURE: the stage loop of
SparseDenseNetRefinementMask.forward after feature extraction
(modules/SparseDenseNetRefinementMask.py:118-212), restated over a plain parameter dict and
feature pyramids, with the sparse ops supplied as callables (C oracle on CPU by default).
Returns every intermediate (the reference's `is_check` taps, :224-225).
"""
from __future__ import annotations

import torch

from . import dense, glue
from . import sparse as osp


def _spamat_cpu(L, R, ml, mr, D):
    return osp.spamat_forward(L, R, ml, mr, D)[0].to(L.device)


def _spavar_cpu(L, R, ml, mr, disp, D):
    return osp.spavar_forward(L, R, ml, mr, disp, D)[0].to(L.device)


def forward(P, left_feats, right_feats, max_disp, left_masks=None, right_masks=None, use_detail=True,
            thold=0.9, skip_stage_id=4, num_stage=4, spamat=_spamat_cpu, spavar=_spavar_cpu):
    taps = {"pred": [], "dense": [], "sparse": [], "var": [], "soft_mask": [], "fusion": [],
            "residual": [], "left_mask": [], "right_mask": [], "left_detail": [], "right_detail": []}
    pred = None
    for s in range(num_stage):
        Lf, Rf = left_feats[f"stage{s}"], right_feats[f"stage{s}"]
        D = max_disp // (3 ** (num_stage - s - 1))
        if s == 0:
            pred, cost, vol = dense.dense_stage(Lf, Rf, D, P)
            taps["cost"] = cost; taps["vol"] = vol
            preL, preR = Lf, Rf
        elif s >= skip_stage_id:
            pred = glue.bicubic_skip(pred, Lf.shape[-2:])
        else:
            l = s - 1
            if use_detail:
                ld = torch.sigmoid(glue.detail_logits(Lf, preL, P, f"detail_detection.{l}"))
                rd = torch.sigmoid(glue.detail_logits(Rf, preR, P, f"detail_detection.{l}"))
                preL, preR = Lf, Rf
                lm, rm = glue.threshold_mask(ld, thold), glue.threshold_mask(rd, thold)
                taps["left_detail"].append(ld); taps["right_detail"].append(rd)
            else:
                lm, rm = left_masks[l], right_masks[l]
            dense_d = glue.dynamic_upsampling(pred, Lf, P, f"dynamic_upsampling.{l}")
            sp = spamat(Lf, Rf, lm, rm, D)
            var = spavar(Lf, Rf, lm, rm, sp, D)
            m = glue.soft_attention(Lf, dense_d, sp, lm, var, P, f"soft_attention.{l}")
            fused = glue.blend(dense_d, sp, m)
            pred, res = glue.refinement(Lf, Rf, fused, P, f"refinement.{l}", s)
            for k, v in (("dense", dense_d), ("sparse", sp), ("var", var), ("soft_mask", m),
                         ("fusion", fused), ("residual", res), ("left_mask", lm), ("right_mask", rm)):
                taps[k].append(v)
        taps["pred"].append(pred)
    return pred, taps
