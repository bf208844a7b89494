"""oracle/features.py -- TEST INFRASTRUCTURE ONLY (never imported by decnet_b200/).

CPU restatement of the reference's feature extractor `FeatExtNetChannelPlus.forward`
(/root/reference/modules/submodule.py:311-343) in functional torch, reading the reference's state_dict
(keys without the `feature_extractor.` prefix):
  unit(x, p)      = relu(batch_norm_eval(conv2d(x, W_p)))                 submodule.py:15-49  (Conv2dUnit)
  deconv(x, p)    = relu(batch_norm_eval(conv_transpose2d(x, W_p, s=3)))  submodule.py:52-87  (Deconv2dUnit)
  block(pre, x)   = unit(unit(cat(deconv(x), pre)))                       submodule.py:162-177 (Deconv2dBlock)
  aspp(x)         = cat(unit_1x1(x), unit_d4(x), unit_d8(x), unit_d12(x)) submodule.py:222-241
Pinned by tests/golden/features.npz (outputs of the UNMODIFIED reference module, made by
tests/golden/make_golden_features.py in the build container).
"""
import torch
import torch.nn.functional as F

EPS = 1e-5


def _unit(x, sd, p, stride=1, padding=0, dilation=1):
    y = F.conv2d(x, sd[p + ".conv.weight"], None, stride=stride, padding=padding, dilation=dilation)
    y = F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"], sd[p + ".bn.bias"],
                     False, 0.0, EPS)
    return F.relu(y)


def _deconv(x, sd, p):
    y = F.conv_transpose2d(x, sd[p + ".conv.weight"], None, stride=3)
    y = F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"], sd[p + ".bn.bias"],
                     False, 0.0, EPS)
    return F.relu(y)


def _block(pre, x, sd, p):
    up = _deconv(x, sd, p + ".deconv")
    y = _unit(torch.cat((up, pre), 1), sd, p + ".conv.0", padding=1)
    return _unit(y, sd, p + ".conv.1", padding=1)


def feature_pyramid(x, sd):
    """x [B,3,H,W] (H, W multiples of 27) -> {"stage0".."stage3"} like the reference's feature_extractor."""
    c0 = _unit(_unit(x, sd, "conv0.0", padding=1), sd, "conv0.1", padding=1)
    c1 = _unit(c0, sd, "conv1.0", stride=3, padding=1)
    c1 = _unit(_unit(c1, sd, "conv1.1", padding=1), sd, "conv1.2", padding=1)
    c2 = _unit(c1, sd, "conv2.0", stride=3, padding=1)
    c2 = _unit(_unit(c2, sd, "conv2.1", padding=1), sd, "conv2.2", padding=1)
    c31 = _unit(c2, sd, "conv3_1", stride=3, padding=1)
    c32 = _unit(_unit(c31, sd, "conv3_2.0", padding=1), sd, "conv3_2.1", padding=1)
    a = "addition_ctx_collection.0.stages."
    ctx = torch.cat([_unit(c31, sd, a + "c0")] + [_unit(c31, sd, a + f"c{i + 1}", padding=r, dilation=r)
                                                   for i, r in enumerate((4, 8, 12))], 1)
    ctx = _unit(ctx, sd, "addition_ctx_collection.1")
    c3 = _unit(torch.cat((c32, ctx), 1), sd, "addition_fusion")
    out = {"stage0": c3}
    r = _block(_unit(c2, sd, "addition_trans2"), c3, sd, "deconv3")
    out["stage1"] = r
    r = _block(_unit(c1, sd, "addition_trans1"), r, sd, "deconv2")
    out["stage2"] = r
    r = _block(_unit(c0, sd, "addition_trans0"), r, sd, "deconv1")
    out["stage3"] = r
    return out
