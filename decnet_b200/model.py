"""Host-side mirror of the reference's hot-path modules, backed by libdecnet_b200.so.

Same class names, constructor arguments, forward signatures and state_dict keys as
modules/submodule.py (GetCostVolume :428, CostRegNetNoDown :608, disparity_regression :766,
GenerateSparseMask :347, DynamicUpsampling :566, SoftAttention :593, Refinement :666) so that a
reference checkpoint loads unchanged and the reference's stage loop can call them.  What runs:

  * hand-written sm_100a kernels (through the C ABI) for everything SURVEY.md section 8 marks
    "ours": cost volume, 3-D aggregation (tcgen05 implicit GEMM), soft-argmin, mask threshold,
    dynamic-upsampling pack + glue, SpaMat / SpaVar, soft-attention pack, sigmoid + blend,
    disparity warp + refinement pack;
  * cuDNN (through torch) for the tiny 2-D conv stacks the north star leaves to the library
    (a5 / a8 / a13 / a14 convs), with eval-mode BatchNorm folded into the conv weights.

Inference only (eval-mode BN).  CUDA tensors only: there is no CPU path.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .modules import SpaMat, SpaVar

BN_EPS = 1e-5
# direct sm_100a kernels for the tiny-channel 2-D convs (section 8f rank 1); False = cuDNN everywhere
USE_NATIVE_CONV2D = True
# TF32 tcgen05 implicit GEMM for the GEMM-sized 2-D convs whenever torch.backends.cudnn.allow_tf32 is on
_CUDNN_FUSED_RELU = True
USE_TF32_TCGEN05 = True


# --------------------------------------------------------------------------------------
# units with the reference's parameter names (conv.weight / conv.bias / bn.*)
# --------------------------------------------------------------------------------------
class Conv2dUnit(nn.Module):
    """Conv2d [+ BN(eval)] [+ ReLU]; keys as modules/submodule.py:15-49."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, relu=True, bn=True,
                 padding=0):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, dilation=dilation,
                              padding=padding, bias=not bn)
        self.bn = nn.BatchNorm2d(out_channels) if bn else None
        self.relu = relu
        self._folded = None

    def folded(self):
        if self._folded is None:
            w = self.conv.weight.detach()
            b = self.conv.bias.detach() if self.conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
            if self.bn is not None:
                scale = self.bn.weight.detach() / torch.sqrt(self.bn.running_var + BN_EPS)
                w = w * scale.view(-1, 1, 1, 1)
                b = (b - self.bn.running_mean) * scale + self.bn.bias.detach()
            self._folded = (w.contiguous(), b.contiguous())
        return self._folded

    def native(self):
        """Packed weights for the direct-conv kernel when this layer is one of its shapes (else None)."""
        c = self.conv
        k, d = c.kernel_size[0], c.dilation[0]
        ok = (c.kernel_size[0] == c.kernel_size[1] and c.stride == (1, 1) and c.dilation[0] == c.dilation[1]
              and c.padding == (d * (k // 2), d * (k // 2)) and c.groups == 1
              and ops.conv2d_small_supported(c.in_channels, c.out_channels, k))
        if not ok:
            return None
        if getattr(self, "_native", None) is None or self._native[0] is not self._folded:
            w, b = self.folded()
            self._native = (self._folded, ops.pack_conv2d_weights(w), b.float().contiguous())
        return self._native

    def tensor_core(self, x):
        """Packed weights for the NCHW TF32 tcgen05 kernel when this layer / input is one of its shapes and
        TF32 is allowed (PyTorch's default for cuDNN convolutions, i.e. what the reference runs), else None.
        Thin layers (Cin < 8 with one output, or wide dilations on 4 channels) stay on the fp32 direct
        kernel, which is faster there (scripts/exp_conv2d_tc.py)."""
        c = self.conv
        d = c.dilation[0]
        cin, cout = c.in_channels, c.out_channels
        ok = (USE_TF32_TCGEN05 and torch.backends.cudnn.allow_tf32 and c.kernel_size == (3, 3) and c.stride == (1, 1)
              and c.dilation == (d, d) and c.padding == (d, d) and c.groups == 1
              and ((d <= 4 and (cin >= 8 or cout >= 3)) or (d <= 8 and cin >= 8))
              and ops.conv2d_tf32_supported(cin, cout, x.shape[2], x.shape[3], d))
        if not ok:
            return None
        if getattr(self, "_tc", None) is None or self._tc[0] is not self._folded:
            w, b = self.folded()
            self._tc = (self._folded,) + ops.pack_conv2d_tf32_nchw_weights(w, b)
        return self._tc

    def forward_cat(self, srcs, w_valid=None):
        """forward(torch.cat(srcs, 1)) with single-channel maps given as [B,H,W]; on the tensor-core path the
        concatenation is never materialised (the kernel reads each source through its own tensor map).
        w_valid: the tensors are right-padded to a 16-byte row pitch (see pad_pitch); columns >= w_valid of the
        result are zeros."""
        x0 = srcs[0]
        chans = tuple(1 if t.dim() == 3 else t.shape[1] for t in srcs)
        c = self.conv
        d = c.dilation[0]
        ok = (len(srcs) <= 3 and x0.is_cuda and x0.dtype == torch.float32 and USE_NATIVE_CONV2D and USE_TF32_TCGEN05
              and torch.backends.cudnn.allow_tf32 and c.kernel_size == (3, 3) and c.stride == (1, 1)
              and c.dilation == (d, d) and c.padding == (d, d) and c.groups == 1 and sum(chans) == c.in_channels
              and ops.conv2d_tf32_supported(ops.padded_cat_channels(chans), c.out_channels, x0.shape[-2], x0.shape[-1], d))
        if not ok:
            return self.forward(torch.cat([t.unsqueeze(1) if t.dim() == 3 else t for t in srcs], 1), w_valid=w_valid)
        cache = getattr(self, "_tc_cat", None)
        if cache is None or cache[0] is not self._folded or cache[1] != chans:
            w, b = self.folded()
            self._tc_cat = (self._folded, chans) + ops.pack_conv2d_tf32_nchw_weights(w, b, chans)
        return ops.conv2d_tf32_nchw_cat([t.contiguous() for t in srcs], self._tc_cat[2], self._tc_cat[3],
                                        c.out_channels, d, self.relu, w_valid or 0)

    def forward(self, x, addend=None, w_valid=None):
        fast = x.is_cuda and x.dtype == torch.float32 and USE_NATIVE_CONV2D
        tc = self.tensor_core(x) if (fast and addend is None) else None
        if tc is not None:
            c = self.conv
            return ops.conv2d_tf32_nchw_cat([x.contiguous()], tc[1], tc[2], c.out_channels, c.dilation[0], self.relu,
                                            w_valid or 0)
        if w_valid:
            out = self.forward(x, addend)
            out[..., w_valid:] = 0                       # keep the pitch padding at zero behind a non-tensor-core layer
            return out
        nat = self.native() if fast else None
        if nat is not None:
            c = self.conv
            return ops.conv2d_small(x.contiguous(), nat[1], nat[2], c.out_channels, c.kernel_size[0], c.dilation[0],
                                    self.relu, addend)
        w, b = self.folded()
        global _CUDNN_FUSED_RELU
        if _CUDNN_FUSED_RELU and self.relu and addend is None and x.is_cuda:
            # library layers (strided / 1x1 / wide convs): cuDNN's fused conv + bias + ReLU, one kernel instead of three
            try:
                return torch.cudnn_convolution_relu(x, w, b, self.conv.stride, self.conv.padding, self.conv.dilation, 1)
            except RuntimeError:
                _CUDNN_FUSED_RELU = False
        x = F.conv2d(x, w, b, stride=self.conv.stride, padding=self.conv.padding, dilation=self.conv.dilation)
        x = F.relu_(x) if self.relu else x
        return x if addend is None else x + addend.unsqueeze(1)


class Deconv2dUnit(nn.Module):
    """ConvTranspose2d [+ BN(eval)] + ReLU; keys as modules/submodule.py:52-87 (bn=False: bias, as in
    GenerateSparseMask.deconv.0; bn=True: no bias, as in Deconv2dBlock.deconv of the feature extractor)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, bn=False):
        super().__init__()
        self.conv = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride=stride, bias=not bn)
        self.bn = nn.BatchNorm2d(out_channels) if bn else None
        self._folded = None

    def folded(self):
        if self._folded is None:
            w = self.conv.weight.detach()                       # [Cin, Cout, k, k]
            b = self.conv.bias.detach() if self.conv.bias is not None else torch.zeros(w.shape[1], device=w.device)
            if self.bn is not None:
                scale = self.bn.weight.detach() / torch.sqrt(self.bn.running_var + BN_EPS)
                w = w * scale.view(1, -1, 1, 1)
                b = (b - self.bn.running_mean) * scale + self.bn.bias.detach()
            self._folded = (w.contiguous(), b.contiguous())
        return self._folded

    def forward(self, x):
        c = self.conv
        w, b = self.folded()
        if (USE_NATIVE_CONV2D and x.is_cuda and x.dtype == torch.float32 and c.kernel_size == (3, 3)
                and c.stride == (3, 3) and c.padding == (0, 0) and ops.deconv3x3s3_supported(c.out_channels)):
            return ops.deconv3x3s3(x.contiguous(), w, b, True)
        return F.relu_(F.conv_transpose2d(x, w, b, stride=c.stride))


class Conv3dUnit(nn.Module):
    """Conv3d 3^3 pad 1 (no bias) + BN3d(eval) [+ ReLU]; keys as modules/submodule.py:90-123."""

    def __init__(self, in_channels, out_channels, relu=True):
        super().__init__()
        self.conv = nn.Conv3d(in_channels, out_channels, 3, padding=1, bias=False)
        self.bn = nn.BatchNorm3d(out_channels)
        self.relu = relu

    def scale_bias(self):
        scale = self.bn.weight.detach() / torch.sqrt(self.bn.running_var + BN_EPS)
        return scale, self.bn.bias.detach() - self.bn.running_mean * scale


def pad_pitch(t):
    """(tensor right-padded with zero columns to a width that is a multiple of 4, original width or None).
    The tensor-core Conv2d reads images through TMA, which needs 16-byte row strides: a KITTI-sized level
    (W = 1269, 423, 141) runs its conv stacks on such padded copies and crops the result (unpad_pitch)."""
    W = t.shape[-1]
    if W % 4 == 0 or not (USE_NATIVE_CONV2D and USE_TF32_TCGEN05 and torch.backends.cudnn.allow_tf32 and t.is_cuda):
        return t, None
    return F.pad(t, (0, 4 - W % 4)), W


def unpad_pitch(t, w_valid):
    return t if w_valid is None else t[..., :w_valid].contiguous()


def _reset_folded(module):
    for m in module.modules():
        if hasattr(m, "_folded"):
            m._folded = None
        if hasattr(m, "_packed"):
            m._packed = None
        if hasattr(m, "_native"):
            m._native = None
        if hasattr(m, "_tc"):
            m._tc = None
        if hasattr(m, "_tc_cat"):
            m._tc_cat = None
        if hasattr(m, "_head"):
            m._head = None


# --------------------------------------------------------------------------------------
# a1/a2: cost volume
# --------------------------------------------------------------------------------------
class GetCostVolume(nn.Module):
    """Drop-in for modules/submodule.py:428-562, restricted to what the shipped model uses:
    warp_ope="homgrp", cost_func="cor", stage-0 candidates 0..D-1 (get_disp_samples :376-390).
    forward(left, right, disp_samples=[B,D,H,W] | max_disp=D) -> [B,C,D,H,W] fp32."""

    def __init__(self, warp_ope="homgrp", cost_func="cor"):
        super().__init__()
        if warp_ope != "homgrp" or cost_func != "cor":
            raise NotImplementedError("decnet_b200 implements the shipped configuration only: homgrp + cor")
        self.warp_ope, self.cost_func = warp_ope, cost_func

    def forward(self, left_feature_map, right_feature_map, **kargs):
        if kargs.get("disp_samples") is not None:
            D = int(kargs["disp_samples"].shape[1])
        else:
            D = int(kargs["max_disp"])
        return ops.cost_volume(left_feature_map.contiguous(), right_feature_map.contiguous(), D)


def get_disp_samples(max_dis, feature_map, stage_id=0, **_unused):
    """Stage-0 branch of modules/submodule.py:376-390 (the only one the model executes)."""
    B, _, H, W = feature_map.shape
    return torch.arange(int(max_dis), dtype=feature_map.dtype, device=feature_map.device) \
        .view(1, -1, 1, 1).expand(B, -1, H, W)


def disparity_regression(cost_vol, disp_samples=None):
    """Drop-in for modules/submodule.py:766-777 with integer candidates 0..D-1."""
    return ops.softargmin(cost_vol.contiguous())


# --------------------------------------------------------------------------------------
# a3: 3-D aggregation
# --------------------------------------------------------------------------------------
class CostRegNetNoDown(nn.Module):
    """Drop-in for modules/submodule.py:608-662 (cost_func 'cor').  forward([B,C,D,H,W]) -> [B,D,H,W].

    `impl`:
      "tcgen05"  hand-written bf16 implicit-GEMM kernels (decnet_conv3d_*), fp32 accumulate
      "cudnn"    torch/cuDNN Conv3d with folded BN (fp32 or bf16 per `dtype`): bring-up/compare only
    """

    def __init__(self, in_channels, base_channels=None, cost_func="cor", down_scale=3, impl="tcgen05",
                 dtype=torch.float32):
        super().__init__()
        if cost_func != "cor":
            raise NotImplementedError("cost_func 'cor' only")
        c = in_channels
        self.conv0 = nn.Sequential(Conv3dUnit(c, c), Conv3dUnit(c, c))
        self.conv1 = nn.Sequential(Conv3dUnit(c, c), Conv3dUnit(c, c), Conv3dUnit(c, c))
        self.conv2 = nn.Sequential(Conv3dUnit(c, c), Conv3dUnit(c, c), Conv3dUnit(c, 1, relu=False))
        self.impl = impl
        self.dtype = dtype
        self._folded = None
        self._packed = None

    def units(self):
        return [*self.conv0, *self.conv1, *self.conv2]

    # ---- cuDNN bring-up path --------------------------------------------------------
    def _fold(self):
        if self._folded is None:
            out = []
            for u in self.units():
                s, b = u.scale_bias()
                w = (u.conv.weight.detach() * s.view(-1, 1, 1, 1, 1)).to(self.dtype)
                out.append((w.contiguous(memory_format=torch.channels_last_3d), b.to(self.dtype), u.relu))
            self._folded = out
        return self._folded

    def _forward_cudnn(self, x):
        f = self._fold()
        x = x.to(self.dtype).contiguous(memory_format=torch.channels_last_3d)

        def run(x, i):
            w, b, relu = f[i]
            y = F.conv3d(x, w, b, padding=1)
            return F.relu_(y) if relu else y
        x = run(x, 0); o0 = run(x, 1)
        x = run(o0, 2); x = run(x, 3); x = run(x, 4) + o0
        x = run(x, 5); x = run(x, 6); x = run(x, 7)
        return x.squeeze(1).float().contiguous()

    def forward(self, x):
        if self.impl == "cudnn":
            return self._forward_cudnn(x)
        from . import conv3d
        return conv3d.cost_regularizer_forward(self, x)


# --------------------------------------------------------------------------------------
# a5: learned lost-detail detector (convs stay cuDNN)
# --------------------------------------------------------------------------------------
class GenerateSparseMask(nn.Module):
    """Drop-in for modules/submodule.py:347-372."""

    def __init__(self, in_channels, down_scale=3):
        super().__init__()
        self.deconv = nn.Sequential(Deconv2dUnit(in_channels * down_scale, 8, 3, 3),
                                    Conv2dUnit(8, 3, 3, padding=1, relu=False))
        self.conv_sub = nn.Sequential(Conv2dUnit(in_channels, 8, 3, padding=1, relu=True, bn=False),
                                      Conv2dUnit(8, 3, 3, padding=1, relu=False))
        self.conv = nn.Sequential(Conv2dUnit(3, 3, 3, padding=1, relu=False),
                                  Conv2dUnit(3, 1, 1, padding=0, relu=False))

    def forward(self, cur_fea, pre_fea):
        pre = self.deconv(pre_fea)
        cur = self.conv_sub(cur_fea)
        res = (cur - pre) ** 2
        return self.conv(res).squeeze(1), cur, pre

    def masks_pair(self, cur_l, pre_l, cur_r, pre_r, thold):
        """Left and right masks `sigmoid(forward(.)) > thold` with the tail fused: one kernel for both
        squared differences, and one for 1x1 conv + BN + sigmoid + threshold of both views (the comparison
        runs on the logit against the exact float where torch.sigmoid crosses `thold`)."""
        # widths that are not a multiple of 4 (KITTI) run on right-padded copies, the padding kept at zero (pad_pitch)
        cur_l, wv = pad_pitch(cur_l)
        cur_r, _ = pad_pitch(cur_r)
        pl = self.deconv[1](pad_pitch(self.deconv[0](pre_l))[0], w_valid=wv)
        pr = self.deconv[1](pad_pitch(self.deconv[0](pre_r))[0], w_valid=wv)
        cl = self.conv_sub[1](self.conv_sub[0](cur_l, w_valid=wv), w_valid=wv)
        cr = self.conv_sub[1](self.conv_sub[0](cur_r, w_valid=wv), w_valid=wv)
        rl, rr = ops.sqdiff_pair(cl, pl, cr, pr)
        xl, xr = self.conv[0](rl, w_valid=wv), self.conv[0](rr, w_valid=wv)
        if getattr(self, "_head", None) is None or self._head[0] is not self.conv[1]._folded:
            w, b = self.conv[1].folded()                    # [1,3,1,1], [1]
            self._head = (self.conv[1]._folded, [float(v) for v in w.flatten().cpu()], float(b.cpu()))
        ml, mr = ops.detail_head(xl.contiguous(), xr.contiguous(), self._head[1], self._head[2],
                                 ops.sigmoid_logit_threshold(thold, xl.device))
        return unpad_pitch(ml, wv), unpad_pitch(mr, wv)


# --------------------------------------------------------------------------------------
# a8: dynamic up-sampling
# --------------------------------------------------------------------------------------
class DynamicUpsampling(nn.Module):
    """Drop-in for modules/submodule.py:566-589: pack kernel -> 3 cuDNN convs -> glue kernel."""

    def __init__(self, in_channels, down_scale=3):
        super().__init__()
        assert down_scale == 3, "the reference hard-codes x3 (SURVEY.md D1)"
        n = down_scale ** 2 * 9
        self._packed = None
        self.weight_learning = nn.Sequential(Conv2dUnit(in_channels * 9 + 1, n, 3, padding=1),
                                             Conv2dUnit(n, n, 3, padding=1),
                                             Conv2dUnit(n, n, 3, padding=1, relu=False))

    def _tf32_pack(self):
        """Weights of the three convs for the TF32 tcgen05 implicit GEMM (cached; reset with the folds)."""
        if getattr(self, "_packed", None) is None:
            units = list(self.weight_learning)
            cin0 = units[0].conv.in_channels
            cp = (cin0 + 7) // 8 * 8
            packed = []
            for u in units:
                w, b = u.folded()
                wp, bp, np_ = ops.pack_conv2d_tf32_weights(w, b, cp)
                packed.append((wp, bp, u.relu))
                cp = np_
            self._packed = ((cin0 + 7) // 8 * 8, packed)
        return self._packed

    def prepack(self, left_fea):
        """The feature channels of the conv input (everything but the disparity channel), or None when the TF32
        tensor-core route is off.  Independent of the disparity: the stage loop runs it ahead, on its second stream."""
        if USE_TF32_TCGEN05 and torch.backends.cudnn.allow_tf32:
            cp0, _ = self._tf32_pack()
            return ops.dynup_pack_nhwc(None, left_fea.contiguous(), cp0, pad=True)
        return None

    def forward(self, disp_map, left_fea, packed=None):
        disp_map = disp_map.contiguous()
        if USE_TF32_TCGEN05 and torch.backends.cudnn.allow_tf32:
            # TF32 allowed (PyTorch's default for convolutions): the three 81-channel convs run as tcgen05
            # implicit GEMMs on channels-last fp32, between channels-last pack / glue kernels
            cp0, packed_w = self._tf32_pack()
            # the tensors carry a one-pixel zero border, so one TMA fill per row tap serves the three column taps
            if packed is not None:
                x = ops.dynup_set_disp_nhwc(packed, disp_map, pad=True)        # feature channels packed ahead (prepack)
            else:
                x = ops.dynup_pack_nhwc(disp_map, left_fea.contiguous(), cp0, pad=True)
            for i, (wp, bp, relu) in enumerate(packed_w):
                x = ops.conv2d_tf32_nhwc_halo(x, wp, bp, relu, round_out=i + 1 < len(packed_w))
            return ops.dynup_glue_nhwc(x, disp_map, pad=True)
        x = ops.dynup_pack(disp_map, left_fea.contiguous())
        logits = self.weight_learning(x)
        return ops.dynup_glue(logits.contiguous(), disp_map)


# --------------------------------------------------------------------------------------
# a13: soft attention (+ blend)
# --------------------------------------------------------------------------------------
class SoftAttention(nn.Module):
    """Drop-in for modules/submodule.py:593-604.  forward(x=[B,C+4,H,W]) -> sigmoid mask like the
    reference; `logits()` + ops.blend() is the fused route the pipeline uses."""

    def __init__(self, in_channels, base_channels):
        super().__init__()
        self.conv = nn.Sequential(Conv2dUnit(in_channels, base_channels, 3, padding=1),
                                  Conv2dUnit(base_channels, base_channels, 3, padding=1),
                                  Conv2dUnit(base_channels, 1, 3, padding=1, relu=False))

    def logits(self, x):
        return self.conv(x)

    def logits_cat(self, left_fea, aux):
        """logits(cat(left_fea, aux)) with aux = [dense, sparse, left_mask, -var] as one [B,4,H,W] tensor."""
        left_fea, wv = pad_pitch(left_fea)
        aux, _ = pad_pitch(aux)
        x = self.conv[0].forward_cat([left_fea, aux], w_valid=wv)
        return unpad_pitch(self.conv[2](self.conv[1](x, w_valid=wv), w_valid=wv), wv)

    def forward(self, x):
        return torch.sigmoid(self.conv(x))


# --------------------------------------------------------------------------------------
# a14: refinement
# --------------------------------------------------------------------------------------
_REFINE_DIL = {0: (1,) * 6, 1: (1,) * 6, 2: (2, 1, 4, 1, 6, 1), 3: (3, 1, 6, 1, 9, 1)}


class Refinement(nn.Module):
    """Drop-in for modules/submodule.py:666-762: warp+pack kernel -> 7 cuDNN convs -> add."""

    def __init__(self, in_channels, base_channels=None, stage_id=-1, down_scale=3):
        super().__init__()
        c = in_channels
        dil = _REFINE_DIL[stage_id]
        chain = [(2 * c + 1, c), (c, c), (c, c), (c, c // 2), (c // 2, c // 2), (c // 2, c // 2)]
        layers = [Conv2dUnit(ci, co, 3, padding=d, dilation=d) for (ci, co), d in zip(chain, dil)]
        layers.append(Conv2dUnit(c // 2, 1, 3, padding=1, relu=False, bn=False))
        self.conv = nn.Sequential(*layers)

    def _wide_pack(self, n_wide):
        """Weights of the first `n_wide` layers for the zero-bordered channels-last TF32 kernel (cached)."""
        if getattr(self, "_packed", None) is None:
            units = list(self.conv)
            cp = (units[0].conv.in_channels + 7) // 8 * 8
            packed = []
            for u in units[:n_wide]:
                w, b = u.folded()
                wp, bp, np_ = ops.pack_conv2d_tf32_weights(w, b, cp)
                packed.append((wp, bp, u.relu))
                cp = np_
            self._packed = ((units[0].conv.in_channels + 7) // 8 * 8, packed)
        return self._packed

    def forward(self, left_fea, right_fea, disp_map):
        disp_map = disp_map.contiguous()
        units = list(self.conv)
        c0 = units[0].conv
        C = left_fea.shape[1]
        if (USE_NATIVE_CONV2D and USE_TF32_TCGEN05 and torch.backends.cudnn.allow_tf32 and C >= 48
                and all(u.conv.dilation == (1, 1) for u in units[:4])):
            # wide level (1/9: 145 -> 72 -> 72 -> 72 -> 36 channels): GEMM-sized layers, too wide for the resident-weight NCHW
            # kernel -> zero-bordered channels-last TF32 kernel between two layout bridges, then back to NCHW for the rest
            cp0, packed = self._wide_pack(4)
            warped = ops.warp_bilinear(right_fea.contiguous(), disp_map)
            x = ops.nchw_cat_to_nhwc_pad([left_fea.contiguous(), warped, disp_map], cp0)
            for i, (wp, bp, relu) in enumerate(packed):
                x = ops.conv2d_tf32_nhwc_halo(x, wp, bp, relu, round_out=i + 1 < len(packed))
            x = ops.nhwc_pad_to_nchw(x, units[3].conv.out_channels)
            for unit in units[4:-1]:
                x = unit(x)
            residual = self.conv[-1](x).squeeze(1)
            return disp_map + residual, residual
        Wp = (left_fea.shape[3] + 3) // 4 * 4
        wv = None
        if (USE_NATIVE_CONV2D and USE_TF32_TCGEN05 and torch.backends.cudnn.allow_tf32
                and ops.conv2d_tf32_supported(ops.padded_cat_channels((C, C, 1)), c0.out_channels, left_fea.shape[2],
                                              Wp, c0.dilation[0])):
            # first conv reads (left, warped right, disparity) as three sources: only the warp is materialised
            warped = ops.warp_bilinear(right_fea.contiguous(), disp_map)
            lp, wv = pad_pitch(left_fea.contiguous())
            x = units[0].forward_cat([lp, pad_pitch(warped)[0], pad_pitch(disp_map)[0]], w_valid=wv)
        else:
            x = units[0](ops.refine_pack(left_fea.contiguous(), right_fea.contiguous(), disp_map))
        for unit in units[1:-1]:
            x = unit(x, w_valid=wv)
        residual = unpad_pitch(self.conv[-1](x, w_valid=wv), wv).squeeze(1)
        return disp_map + residual, residual


# --------------------------------------------------------------------------------------
# a16: the stage loop
# --------------------------------------------------------------------------------------
class DecompMatching(nn.Module):
    """The decomposed-matching hot path: body of SparseDenseNetRefinementMask.forward after feature
    extraction (modules/SparseDenseNetRefinementMask.py:118-212), same hyper-parameters and the
    same sub-module names, so `load_state_dict(reference_state, strict=False)` picks up every
    hot-path weight (feature_extractor.* is ignored: out of scope).

    forward(left_feats, right_feats, left_mask_list=None, right_mask_list=None, is_check=False)
      left_feats/right_feats: {"stage0".."stage3"} NCHW fp32 CUDA tensors (C = 216,72,24,8)
      returns [pred] like the reference's inference path, or (pred, taps) with is_check.
    """

    def __init__(self, max_disp=216, base_channels=8, num_stage=4, down_scale=3, skip_stage_id=4,
                 use_detail=True, thold=0.9, conv3d_impl="tcgen05", channels=None):
        super().__init__()
        assert down_scale == 3 and num_stage == 4, "shipped configuration (demo.sh:1)"
        assert max_disp % (down_scale ** (num_stage - 1)) == 0, "max_disp must be a multiple of 27"
        self.max_disp, self.num_stage, self.down_scale = max_disp, num_stage, down_scale
        self.skip_stage_id, self.use_detail, self.thold = skip_stage_id, use_detail, thold
        ch = list(channels) if channels is not None else [27 * base_channels, 9 * base_channels,
                                                          3 * base_channels, base_channels]
        self.channels = ch
        self.get_cost_volume = GetCostVolume("homgrp", "cor")
        self.sparse_matching = nn.ModuleList([SpaMat() for _ in range(num_stage - 1)])
        self.sparse_var = nn.ModuleList([SpaVar() for _ in range(num_stage - 1)])
        self.cost_regularizer = CostRegNetNoDown(ch[0], ch[0] * 2, "cor", down_scale, impl=conv3d_impl)
        self.detail_detection = nn.ModuleList([GenerateSparseMask(ch[i + 1], down_scale) for i in range(num_stage - 1)])
        self.dynamic_upsampling = nn.ModuleList([DynamicUpsampling(ch[i + 1], down_scale) for i in range(num_stage - 1)])
        self.soft_attention = nn.ModuleList([SoftAttention(ch[i + 1] + 4, base_channels) for i in range(num_stage - 1)])
        self.refinement = nn.ModuleList([Refinement(ch[i + 1], base_channels // (2 ** i), stage_id=i + 1,
                                                    down_scale=down_scale) for i in range(num_stage - 1)])
        self.eval()

    def load_state_dict(self, state_dict, strict=False, **kw):
        sd = {k[7:] if k.startswith("module.") else k: v for k, v in state_dict.items()}   # demo.py:124-135
        sd = {k: v for k, v in sd.items() if not k.startswith("feature_extractor")}
        res = super().load_state_dict(sd, strict=strict, **kw)
        _reset_folded(self)
        return res

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        _reset_folded(self)
        return out

    # Masks (a5/a6) and the sparse ops (a9/a10) of every level depend on the features only, not on the disparity
    # coming up from the coarser level: they run on a second stream, forked at the start of forward() and joined
    # where the soft attention first needs them, so that they fill the SMs the main chain leaves idle (the last,
    # partial round of tiles of the persistent conv kernels, launch gaps).  Captured in a CUDA graph this becomes two
    # parallel branches.  `overlap = False` restores the single-stream order.
    overlap = True

    @torch.no_grad()
    def forward(self, left_feats, right_feats, left_mask_list=None, right_mask_list=None, is_check=False):
        taps = {k: [] for k in ("pred", "dense", "sparse", "var", "soft_mask", "fusion", "residual",
                                "left_mask", "right_mask", "left_detail", "right_detail")} if is_check else None
        pred = None
        branch = events = packs = pack_events = side = None
        if self.overlap and not is_check:
            dev = left_feats["stage0"].device
            main = torch.cuda.current_stream(dev)
            if getattr(self, "_side", None) is None or self._side.device != dev:
                self._side = torch.cuda.Stream(dev)
            side = self._side
            side.wait_stream(main)                       # fork: the features are ready on the main stream
            with torch.cuda.stream(side):
                branch, events, packs, pack_events = [], [], [], []
                # first what the main chain needs first: the feature channels of every level's up-sampling input
                for s in range(1, self.num_stage):
                    pk = None
                    if s < self.skip_stage_id:
                        pk = self.dynamic_upsampling[s - 1].prepack(left_feats[f"stage{s}"])
                    ev = None
                    if pk is not None:
                        ev = torch.cuda.Event()
                        ev.record(side)
                    packs.append(pk); pack_events.append(ev)
                pre = (left_feats["stage0"].contiguous(), right_feats["stage0"].contiguous())
                for s in range(1, self.num_stage):
                    if s >= self.skip_stage_id:
                        branch.append(None); events.append(None)
                        continue
                    l = s - 1
                    Lf, Rf = left_feats[f"stage{s}"].contiguous(), right_feats[f"stage{s}"].contiguous()
                    D = self.max_disp // (self.down_scale ** (self.num_stage - s - 1))
                    if self.use_detail:
                        lm, rm = self.detail_detection[l].masks_pair(Lf, pre[0], Rf, pre[1], self.thold)
                        pre = (Lf, Rf)
                    else:
                        lm, rm = left_mask_list[l].contiguous(), right_mask_list[l].contiguous()
                    sparse, var, _, _ = ops.spamat_spavar_forward(Lf, Rf, lm, rm, D)
                    branch.append((lm, rm, sparse, var))
                    ev = torch.cuda.Event()
                    ev.record(side)
                    events.append(ev)
        pre_l = pre_r = None
        for s in range(self.num_stage):
            Lf = left_feats[f"stage{s}"].contiguous()
            Rf = right_feats[f"stage{s}"].contiguous()
            D = self.max_disp // (self.down_scale ** (self.num_stage - s - 1))
            if s == 0:
                pred, cost = self.dense_stage(Lf, Rf, D)
                if is_check:
                    taps["cost"] = cost
                pre_l, pre_r = Lf, Rf
            elif s >= self.skip_stage_id:
                # SparseDenseNetRefinementMask.py:143-144 (Middlebury's finest level); single ATen kernel
                pred = F.interpolate(pred.unsqueeze(1) * self.down_scale, list(Lf.shape[-2:]), mode="bicubic").squeeze(1)
            else:
                l = s - 1
                if branch is not None:
                    if pack_events[l] is not None:
                        torch.cuda.current_stream(Lf.device).wait_event(pack_events[l])
                    dense = self.dynamic_upsampling[l](pred, Lf, packed=packs[l])
                    torch.cuda.current_stream(Lf.device).wait_event(events[l])      # join for this level
                    lm, rm, sparse, var = branch[l]
                else:
                    if self.use_detail:
                        lm, rm = self.detail_detection[l].masks_pair(Lf, pre_l, Rf, pre_r, self.thold)
                        if is_check:
                            ld, _, _ = self.detail_detection[l](Lf, pre_l)
                            rd, _, _ = self.detail_detection[l](Rf, pre_r)
                            taps["left_detail"].append(torch.sigmoid(ld).contiguous())
                            taps["right_detail"].append(torch.sigmoid(rd).contiguous())
                        pre_l, pre_r = Lf, Rf
                    else:
                        lm, rm = left_mask_list[l].contiguous(), right_mask_list[l].contiguous()
                    dense = self.dynamic_upsampling[l](pred, Lf)
                    sparse, var, _, _ = ops.spamat_spavar_forward(Lf, Rf, lm, rm, D)     # SpaMat + SpaVar, one pass
                aux = ops.attn_pack(None, dense, sparse, lm, var)                  # [dense, sparse, mask, -var]
                logit = self.soft_attention[l].logits_cat(Lf, aux).squeeze(1).contiguous()
                soft, fused = ops.blend(logit, dense, sparse, want_mask=is_check)
                pred, residual = self.refinement[l](Lf, Rf, fused)
                if is_check:
                    for k, v in (("dense", dense), ("sparse", sparse), ("var", var), ("soft_mask", soft),
                                 ("fusion", fused), ("residual", residual), ("left_mask", lm), ("right_mask", rm)):
                        taps[k].append(v)
            if is_check:
                taps["pred"].append(pred)
        if side is not None:
            # join (a captured graph needs every forked stream back).  The side stream's tensors are consumed on the
            # main stream; their memory returns to the side stream's pool and is reused only by the next forward's
            # branch, which starts behind that forward's fork, i.e. behind everything enqueued here.
            torch.cuda.current_stream(pred.device).wait_stream(side)
        return (pred, taps) if is_check else [pred]

    def dense_stage(self, Lf, Rf, D):
        """a1-a4: cost volume -> 3-D aggregation -> soft-argmin.  Returns (pred [B,H,W], cost [B,D,H,W])."""
        if self.cost_regularizer.impl == "tcgen05":
            from . import conv3d
            cost = conv3d.dense_cost(self.cost_regularizer, Lf, Rf, D)
        else:
            vol = ops.cost_volume(Lf, Rf, D)
            cost = self.cost_regularizer(vol)
        return ops.softargmin(cost), cost
