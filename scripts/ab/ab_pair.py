"""One-tile-per-CTA kernel against the pair kernel (variant 2) in the current split kind: time and max difference."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import _lib, ops

g = torch.Generator(device="cuda").manual_seed(0)
for (B, h, w, cin, cout) in [(8, 180, 324, 81, 81), (8, 180, 324, 81, 64), (8, 180, 324, 81, 48), (8, 180, 324, 145, 64), (16, 180, 324, 24, 96)]:
    cp = (cin + 7) // 8 * 8
    x = torch.zeros(B, h + 2, w + 2, cp, device="cuda")
    x[:, 1:-1, 1:-1, :cin] = torch.randn(B, h, w, cin, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.05
    wp, bp, np_ = ops.pack_conv2d_tf32_weights(wt, torch.zeros(cout, device="cuda"), cp, split=True)
    outs = {}
    for variant in (0, 2):
        _lib.lib().decnet_conv2d_nhwc_set_variant(variant)
        for _ in range(3):
            y = ops.conv2d_tf32_nhwc_halo(x, wp, bp, True, split=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            y = ops.conv2d_tf32_nhwc_halo(x, wp, bp, True, split=True)
        e1.record(); torch.cuda.synchronize()
        outs[variant] = (y, e0.elapsed_time(e1) / 20 * 1e3)
    _lib.lib().decnet_conv2d_nhwc_set_variant(0)
    print(f"kind {ops.SPLIT_KIND} B{B} {h}x{w} {cin}->{cout}: single {outs[0][1]:7.1f} us, pair {outs[2][1]:7.1f} us, max diff {(outs[0][0] - outs[2][0]).abs().max().item():.2e}")
