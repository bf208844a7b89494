"""the step BEFORE the hot path -- `FeatExtNetChannelPlus`
(modules/submodule.py:245-343), the U-shaped extractor that turns an image into the
{stage0..stage3} feature pyramid the decomposed-matching path consumes.

Same class name, constructor arguments, sub-module names and state_dict keys as the reference
(`feature_extractor.conv0.0.conv.weight`, `...deconv2.deconv.bn.running_var`, ...), inference
only (BN folded).  The layers are the units of model.py, so each one runs on the kernel that
fits it:
  * 3x3 stride-1 convs up to 24 channels (full and 1/3 resolution: where the pixels are):
    `conv2d_tcgen05_kernel` (TF32 tensor cores, NCHW); the `cat(x_up, x_pre)` in front of each
    Deconv2dBlock conv is read as two sources, never materialised;
  * ConvTranspose 3x3 stride 3 to 8 / 24 channels: `deconv3x3s3_kernel`;
  * 1x1 convs at full resolution: `conv2d_small_kernel`;
  * the stride-3 convs, the 72 / 216-channel layers and the ASPP at 1/9 and 1/27 resolution
    (6 480 and 720 pixels per SceneFlow image) stay on cuDNN with folded weights: library GEMMs,
    about 7 GFLOP per image in total.
The module refuses CPU tensors like the rest of the package; `oracle/features.py` is the CPU
restatement used by the tests.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .model import Conv2dUnit, Deconv2dUnit, set_precision


class Deconv2dBlock(nn.Module):
    """Drop-in for modules/submodule.py:162-177: deconv(x) -> cat(x_up, x_pre) -> 2 convs."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1):
        super().__init__()
        self.deconv = Deconv2dUnit(in_channels, out_channels, kernel_size, stride, bn=True)
        self.conv = nn.Sequential(Conv2dUnit(out_channels * 2, out_channels, 3, 1, padding=1),
                                  Conv2dUnit(out_channels, out_channels, 3, 1, padding=1))

    def forward(self, x_pre, x):
        x_up = self.deconv(x)
        y = self.conv[0].forward_cat([x_up, x_pre])
        return self.conv[1](y), x_up


class ASPP(nn.Module):
    """Drop-in for modules/submodule.py:222-241 (no image-pool branch in the shipped version)."""

    def __init__(self, in_ch, out_ch, rates):
        super().__init__()
        self.stages = nn.Module()
        self.stages.add_module("c0", Conv2dUnit(in_ch, out_ch, 1, stride=1, padding=0, dilation=1))
        for i, rate in enumerate(rates):
            self.stages.add_module(f"c{i + 1}", Conv2dUnit(in_ch, out_ch, 3, stride=1, padding=rate, dilation=rate))

    def forward(self, x):
        return torch.cat([stage(x) for stage in self.stages.children()], dim=1)


class FeatExtNetChannelPlus(nn.Module):
    """Drop-in for modules/submodule.py:245-343 with num_stage = 4, down_scale = 3 (the shipped config)."""

    def __init__(self, base_channels, num_stage=4, down_scale=3, precision="fp32"):
        super().__init__()
        assert num_stage == 4 and down_scale == 3, "the shipped configuration (demo.sh:1) is 4 stages, x3"
        c, s = base_channels, down_scale
        self.base_channels, self.num_stage, self.down_scale = c, num_stage, s
        self.conv0 = nn.Sequential(Conv2dUnit(3, c, 3, 1, padding=1), Conv2dUnit(c, c, 3, 1, padding=1))
        self.addition_trans0 = Conv2dUnit(c, c, 1, stride=1, padding=0)
        self.conv1 = nn.Sequential(Conv2dUnit(c, c * s, 3, stride=s, padding=1),
                                   Conv2dUnit(c * s, c * s, 3, 1, padding=1), Conv2dUnit(c * s, c * s, 3, 1, padding=1))
        self.addition_trans1 = Conv2dUnit(c * s, c * s, 1, stride=1, padding=0)
        self.deconv1 = Deconv2dBlock(c * s, c, kernel_size=3, stride=3)
        self.conv2 = nn.Sequential(Conv2dUnit(c * s, c * s ** 2, 3, stride=3, padding=1),
                                   Conv2dUnit(c * s ** 2, c * s ** 2, 3, 1, padding=1),
                                   Conv2dUnit(c * s ** 2, c * s ** 2, 3, 1, padding=1))
        self.addition_trans2 = Conv2dUnit(c * s ** 2, c * s ** 2, 1, stride=1, padding=0)
        self.deconv2 = Deconv2dBlock(c * s ** 2, c * s, kernel_size=3, stride=3)
        self.conv3_1 = Conv2dUnit(c * s ** 2, c * s ** 3, 3, stride=3, padding=1)
        self.conv3_2 = nn.Sequential(Conv2dUnit(c * s ** 3, c * s ** 3, 3, 1, padding=1),
                                     Conv2dUnit(c * s ** 3, c * s ** 3, 3, 1, padding=1))
        self.addition_ctx_collection = nn.Sequential(ASPP(c * s ** 3, c * s ** 3, [4, 8, 12]),
                                                     Conv2dUnit(4 * c * s ** 3, c * s ** 3, 1, stride=1, padding=0))
        self.addition_fusion = Conv2dUnit(2 * c * s ** 3, c * s ** 3, 1, stride=1, padding=0)
        self.deconv3 = Deconv2dBlock(c * s ** 3, c * s ** 2, kernel_size=3, stride=3)
        self.out_channels = [c * s ** 3, c * s ** 2, c * s, c]
        # the strided / wide / ASPP layers at 1/9 and 1/27 resolution have no kernel of ours yet: library layers
        for m in self.modules():
            if isinstance(m, (Conv2dUnit, Deconv2dUnit)):
                m.library_ok = True
        set_precision(self, precision)
        self.precision = precision
        self.eval()

    @torch.no_grad()
    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("decnet_b200 has no CPU path: FeatExtNetChannelPlus needs a CUDA tensor "
                               "(oracle/features.py is the CPU restatement for tests)")
        x = x.contiguous().float()
        conv0 = self.conv0(x)                                   # [B,  c, H,    W   ]
        conv1 = self.conv1(conv0)                               # [B, 3c, H/3,  W/3 ]
        conv2 = self.conv2(conv1)                               # [B, 9c, H/9,  W/9 ]
        conv3_1 = self.conv3_1(conv2)                           # [B,27c, H/27, W/27]
        conv3_2 = self.conv3_2(conv3_1)
        ctx = self.addition_ctx_collection(conv3_1)
        conv3 = self.addition_fusion(torch.cat((conv3_2, ctx), dim=1))
        out = {"stage0": conv3}
        res, _ = self.deconv3(self.addition_trans2(conv2), conv3)
        out["stage1"] = res
        res, _ = self.deconv2(self.addition_trans1(conv1), res)
        out["stage2"] = res
        res, _ = self.deconv1(self.addition_trans0(conv0), res)
        out["stage3"] = res
        return out


def extract_pair(fe, left, right, prepare=None):
    """Features of both views, the right view on a forked second stream (the two extractions are independent;
    demo.py:166-167 runs them back to back inside the model).  `prepare` is applied to each input first
    (e.g. the uint8 -> padded / normalised image kernel).  Returns (left_feats, right_feats), joined on the
    current stream; capturable in a CUDA graph (two branches)."""
    main = torch.cuda.current_stream(left.device)
    side = getattr(fe, "_side", None)
    if side is None or side.device != left.device:
        side = fe._side = torch.cuda.Stream(left.device)
    side.wait_stream(main)                               # fork
    with torch.cuda.stream(side):
        fr = fe(prepare(right) if prepare is not None else right)
    fl = fe(prepare(left) if prepare is not None else left)
    # join; the right view's tensors live in the side stream's pool and are reused only behind the next fork
    main.wait_stream(side)
    return fl, fr
