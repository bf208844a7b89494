"""A few launches of conv2d_nhwc_halo_kernel (81 -> 81 channels, 180x324, B = 8) for ncu: argv[1] = fp32 | tf32."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from decnet_b200 import ops
split = (sys.argv[1] if len(sys.argv) > 1 else "fp32") == "fp32"
g = torch.Generator(device="cuda").manual_seed(0)
B, h, w, cin, cout = 8, 180, 324, 81, 81
cp = 88
x = torch.zeros(B, h + 2, w + 2, cp, device="cuda")
x[:, 1:-1, 1:-1, :cin] = torch.randn(B, h, w, cin, device="cuda", generator=g)
wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.05
wp, bp, np_ = ops.pack_conv2d_tf32_weights(wt, torch.zeros(cout, device="cuda"), cp, split=split)
xx = x if split else ops.rna_tf32(x)
for _ in range(4):
    y = ops.conv2d_tf32_nhwc_halo(xx, wp, bp, True, split=split)
torch.cuda.synchronize()
