"""A few launches of one glue kernel at the finest level (B = 8, C = 8, 540x972) for ncu: argv[1] in {warp, deconv, dynup_pack}."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from decnet_b200 import ops
which = sys.argv[1]
g = torch.Generator(device="cuda").manual_seed(0)
B, C, H, W = 8, 8, 540, 972
if which == "warp":
    R = torch.randn(B, C, H, W, device="cuda", generator=g)
    d = torch.rand(B, H, W, device="cuda", generator=g) * 200
    for _ in range(4):
        ops.warp_bilinear(R, d)
elif which == "deconv":
    x = torch.randn(B, 24, H // 3, W // 3, device="cuda", generator=g)
    w = torch.randn(24, 8, 3, 3, device="cuda", generator=g) * 0.1
    b = torch.zeros(8, device="cuda")
    for _ in range(4):
        ops.deconv3x3s3(x, w, b, True)
torch.cuda.synchronize()
