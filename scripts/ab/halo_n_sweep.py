"""MMA-bound time of conv2d_nhwc_halo_kernel against N (output channels): knock-out mask 9 (converters idle, no weight TMA
after the first ring fill) leaves A fills + all MMAs.  Tuning only."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import _lib, ops

g = torch.Generator(device="cuda").manual_seed(0)
B, h, w, cin = 8, 180, 324, 81
cp = 88
x = torch.zeros(B, h + 2, w + 2, cp, device="cuda")
x[:, 1:-1, 1:-1, :cin] = torch.randn(B, h, w, cin, device="cuda", generator=g)
for cout in (16, 32, 48, 64, 80, 96):
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.05
    wp, bp, np_ = ops.pack_conv2d_tf32_weights(wt, torch.zeros(cout, device="cuda"), cp, split=True)
    row = []
    for mask in (0, 9, 15):
        _lib.lib().decnet_conv2d_nhwc_set_variant(100 + mask)
        for _ in range(3):
            ops.conv2d_tf32_nhwc_halo(x, wp, bp, True, split=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.conv2d_tf32_nhwc_halo(x, wp, bp, True, split=True)
        e1.record(); torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 10 * 1e3)
    _lib.lib().decnet_conv2d_nhwc_set_variant(0)
    print(f"np {np_:3d}: full {row[0]:7.1f} us, MMAs + A fills only {row[1]:7.1f} us, A fills only {row[2]:7.1f} us")
