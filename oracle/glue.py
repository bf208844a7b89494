"""oracle/glue.py -- TEST INFRASTRUCTURE: torch restatement of the per-level glue around the
sparse ops.

Follows modules/submodule.py:347-372 (GenerateSparseMask), SparseDenseNetRefinementMask.py:158-170
(sigmoid + threshold), submodule.py:566-589 (DynamicUpsampling), :593-604 + model :197-202
(SoftAttention + blend), :719-762 (Refinement: warp by fractional disparity + 7 convs, dilated at
stages 2/3 :697-716), model :143-144 (bicubic skip stage), utils/Wavelet.py:8-123 (Haar; the
reference's filter pickle is absent -> orthonormal Haar substituted, PARITY UNPINNED for a7).
Conv / BN arithmetic is ATen's.  Pinned by tests/golden (see oracle/dense.py header).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .dense import BN_EPS, _sample_coords, bilinear_zero_pad


def conv_bn(x, P, prefix, relu=True, bn=True, padding=1, dilation=1, stride=1):
    x = F.conv2d(x, P[prefix + ".conv.weight"], P.get(prefix + ".conv.bias"), stride=stride,
                 padding=padding, dilation=dilation)
    if bn:
        x = F.batch_norm(x, P[prefix + ".bn.running_mean"], P[prefix + ".bn.running_var"],
                         P[prefix + ".bn.weight"], P[prefix + ".bn.bias"], False, 0.0, BN_EPS)
    return F.relu(x) if relu else x


# ---- a5: detail logits ---------------------------------------------------------------
def detail_logits(cur, prev, P, prefix):
    pre = F.relu(F.conv_transpose2d(prev, P[prefix + ".deconv.0.conv.weight"],
                                    P[prefix + ".deconv.0.conv.bias"], stride=3))
    pre = conv_bn(pre, P, prefix + ".deconv.1", relu=False)
    c = conv_bn(cur, P, prefix + ".conv_sub.0", relu=True, bn=False)
    c = conv_bn(c, P, prefix + ".conv_sub.1", relu=False)
    res = (c - pre) ** 2
    x = conv_bn(res, P, prefix + ".conv.0", relu=False)
    x = conv_bn(x, P, prefix + ".conv.1", relu=False, padding=0)
    return x.squeeze(1)


# ---- a6: mask selection --------------------------------------------------------------
def threshold_mask(prob, thold):
    """m = prob > thold ? 1 : 0 with NaN kept (the reference's two boolean index_puts leave NaN)."""
    out = prob.clone()
    out[prob <= thold] = 0.0
    out[prob > thold] = 1.0
    return out


# ---- a8: dynamic up-sampling ---------------------------------------------------------
def dynup_pack(disp, Lf):
    """conv input [B, 1+9C, h, w]: ch0 = disp, ch 1+c*9+ky*3+kx = Lf[c, 3h+ky, 3w+kx]."""
    B, h, w = disp.shape
    C = Lf.shape[1]
    x = Lf.reshape(B, C, h, 3, w, 3).permute(0, 1, 3, 5, 2, 4).reshape(B, C * 9, h, w)
    return torch.cat((disp.unsqueeze(1), x), dim=1)


def dynup_glue(logits, disp):
    """logits [B,81,h,w] (channel = sub*9 + k, sub = i*3+j, k = ky*3+kx), disp [B,h,w]
    -> [B,3h,3w]: out[3h+i,3w+j] = 3 * sum_k softmax_k(logits[sub]) * disp_reppad[h+ky-1, w+kx-1]."""
    B, _, h, w = logits.shape
    wts = torch.softmax(logits.reshape(B, 9, 9, h, w), dim=2)
    pad = F.pad(disp.unsqueeze(1), (1, 1, 1, 1), mode="replicate")[:, 0]
    nb = torch.stack([pad[:, ky:ky + h, kx:kx + w] for ky in range(3) for kx in range(3)], dim=1)  # [B,9,h,w]
    res = (wts * nb.unsqueeze(1)).sum(2)                               # [B,9(sub),h,w]
    out = res.reshape(B, 3, 3, h, w).permute(0, 3, 1, 4, 2).reshape(B, 3 * h, 3 * w)
    return out * 3.0


def dynamic_upsampling(disp, Lf, P, prefix):
    x = dynup_pack(disp, Lf)
    x = conv_bn(x, P, prefix + ".weight_learning.0")
    x = conv_bn(x, P, prefix + ".weight_learning.1")
    x = conv_bn(x, P, prefix + ".weight_learning.2", relu=False)
    return dynup_glue(x, disp)


# ---- a13: soft attention + blend -----------------------------------------------------
def soft_attention(Lf, dense, sparse, lmask, var, P, prefix):
    x = torch.cat((Lf, dense.unsqueeze(1), sparse.unsqueeze(1), lmask.unsqueeze(1), -var.unsqueeze(1)), dim=1)
    x = conv_bn(x, P, prefix + ".conv.0")
    x = conv_bn(x, P, prefix + ".conv.1")
    x = conv_bn(x, P, prefix + ".conv.2", relu=False)
    return torch.sigmoid(x).squeeze(1)


def blend(dense, sparse, m):
    return dense * (1 - m) + m * sparse


# ---- a14: refinement -----------------------------------------------------------------
REFINE_DILATIONS = {1: (1, 1, 1, 1, 1, 1, 1), 2: (2, 1, 4, 1, 6, 1, 1), 3: (3, 1, 6, 1, 9, 1, 1)}


def warp_by_disparity(R, disp):
    """warped[b,c,h,w] = bilinear0(R[b,c], y'(h), x'(w - disp[b,h,w]))  (submodule.py:719-745)."""
    B, C, H, W = R.shape
    dt, dev = R.dtype, R.device
    pos = torch.arange(W, dtype=dt, device=dev).view(1, 1, W)
    g = (pos - disp) / ((W - 1.0) / 2.0) - 1.0
    ix = ((g + 1.0) * W - 1.0) / 2.0
    iy = _sample_coords(H, H, 0.0, dt, dev).view(1, H, 1).expand(B, H, W)
    return bilinear_zero_pad(R, ix, iy)


def refinement(Lf, Rf, disp, P, prefix, stage_id):
    warped = warp_by_disparity(Rf, disp)
    x = torch.cat((Lf, warped, disp.unsqueeze(1)), dim=1)
    dil = REFINE_DILATIONS[stage_id]
    for i in range(6):
        x = conv_bn(x, P, f"{prefix}.conv.{i}", padding=dil[i], dilation=dil[i])
    res = conv_bn(x, P, f"{prefix}.conv.6", relu=False, bn=False).squeeze(1)
    return disp + res, res


# ---- a15: skip stage -----------------------------------------------------------------
def bicubic_skip(pred, size):
    return F.interpolate(pred.unsqueeze(1) * 3, list(size), mode="bicubic").squeeze(1)


# ---- a7: Haar wavelet lost-detail masks (parity unpinned, see header) ------------------
def haar_analysis(x):
    """x [B,1,H,W] (H, W even) -> LL, LH, HL, HH each [B,1,H/2,W/2]; orthonormal Haar
    (2x2 stride-2 analysis, the `rec2` filter bank of utils/Wavelet.py:29-51)."""
    x = x[:, :, : x.shape[2] // 2 * 2, : x.shape[3] // 2 * 2]      # stride-2 conv without padding drops an odd tail
    a = x[:, :, 0::2, 0::2]; b = x[:, :, 0::2, 1::2]
    c = x[:, :, 1::2, 0::2]; d = x[:, :, 1::2, 1::2]
    ll = (a + b + c + d) * 0.5
    lh = (a - b + c - d) * 0.5
    hl = (a + b - c - d) * 0.5
    hh = (a - b - c + d) * 0.5
    return ll, lh, hl, hh


def haar_detail_masks(x, levels):
    """Per x2 level: v = max(|LH|,|HL|,|HH|), min-max normalise per image, threshold t = smallest of
    {0.1,...,1.0} with >= 85% of pixels <= t, mask = v >= t (utils/Wavelet.py:66-123).  Returns
    (masks list fine->coarse, final LL)."""
    import numpy as np
    masks = []
    ll = x
    for _ in range(levels):
        ll, lh, hl, hh = haar_analysis(ll)
        v = torch.maximum(torch.maximum(lh.abs(), hl.abs()), hh.abs())
        B = v.shape[0]
        flat = v.reshape(B, -1)
        mn = flat.min(1).values.view(B, 1, 1, 1); mx = flat.max(1).values.view(B, 1, 1, 1)
        vn = (v - mn) / (mx - mn)
        ms = []
        for bi in range(B):
            t_sel = 1.0
            n = vn[bi].numel()
            for t in (np.arange(0, 1, 0.1) + 0.1):
                if float((vn[bi] <= float(np.float32(t))).sum()) / n >= 0.85:
                    t_sel = float(np.float32(t)); break
            ms.append((vn[bi] >= t_sel).to(x.dtype))
        masks.append(torch.stack(ms, 0))
    return masks, ll
