"""A few launches of one hot kernel for `ncu --set full`: argv[1] in {conv3d, thin, sparse_multi}."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from decnet_b200 import conv3d as c3, ops
which = sys.argv[1]
g = torch.Generator(device="cuda").manual_seed(0)
if which == "conv3d":
    B, D, H, W, C = 8, 8, 20, 36, 224
    x = torch.randn(B, D, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(27, C, C, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
    bias = torch.zeros(C, device="cuda")
    for _ in range(4):
        c3.conv3d_layer(x, w, bias, C, True)
elif which == "thin":
    x = torch.randn(8, 8, 540, 972, device="cuda", generator=g)
    w = torch.randn(8, 8, 3, 3, device="cuda", generator=g) * 0.1
    wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, torch.zeros(8, device="cuda"), split=True)
    for _ in range(4):
        ops.conv2d_tf32_nchw_cat([x], wp, bp, 8, 1, True, split=True)
else:
    from helpers import make_feats
    lv = []
    for i, (C, H, W, D) in enumerate(((8, 540, 972, 216), (24, 180, 324, 72), (72, 60, 108, 24))):
        L, R = make_feats(8, C, H, W, seed=17 + i, device="cuda")
        pl = torch.rand(8, H, W, device="cuda", generator=g)
        pr = torch.rand(8, H, W, device="cuda", generator=g)
        ml, mr = ops.mask_threshold(pl, pr, 0.9)
        lv.append((L, R, ml, mr, D))
    for _ in range(4):
        ops.spamat_spavar_forward_levels(lv)
torch.cuda.synchronize()
