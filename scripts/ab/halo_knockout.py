"""Knock-out timing of conv2d_nhwc_halo_kernel (tuning only; results are wrong with any mask set): which role is on the
critical path?  mask bits: 1 converters idle, 2 no correction MMAs, 4 no hi*hi MMAs, 8 no weight TMA after the first ring
fill."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import _lib, ops

g = torch.Generator(device="cuda").manual_seed(0)
B, h, w, cin, cout = 8, 180, 324, 81, 81
cp = 88
x = torch.zeros(B, h + 2, w + 2, cp, device="cuda")
x[:, 1:-1, 1:-1, :cin] = torch.randn(B, h, w, cin, device="cuda", generator=g)
wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.05
for split in (True, False):
    wp, bp, np_ = ops.pack_conv2d_tf32_weights(wt, torch.zeros(cout, device="cuda"), cp, split=split)
    xx = x if split else ops.rna_tf32(x)
    for mask in ([0, 1, 2, 4, 8, 3, 6, 7, 9, 10, 12, 11, 14, 15] if split else [0, 4, 8, 12]):
        _lib.lib().decnet_conv2d_nhwc_set_variant(100 + mask)
        for _ in range(3):
            ops.conv2d_tf32_nhwc_halo(xx, wp, bp, True, split=split)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv2d_tf32_nhwc_halo(xx, wp, bp, True, split=split)
        e1.record(); torch.cuda.synchronize()
        print(f"split_kind {ops.SPLIT_KIND if split else 0} mask {mask:2d}: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us")
    _lib.lib().decnet_conv2d_nhwc_set_variant(0)
