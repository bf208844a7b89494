"""A/B of the channels-last kernel's fp32-class modes: time and error against an fp64 convolution.
Run once per kind: DECNET_SPLIT_KIND=1 python scripts/ab/ab_split_kind.py ; DECNET_SPLIT_KIND=2 python scripts/ab/ab_split_kind.py"""
import sys
from pathlib import Path
import torch
import torch.nn.functional as F
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import ops

g = torch.Generator(device="cuda").manual_seed(3)
print("SPLIT_KIND", ops.SPLIT_KIND)
for (B, H, W, Cin, Cout, scale_x, scale_w) in [(8, 180, 324, 81, 81, 1.0, 1.0), (8, 180, 324, 73, 81, 1.0, 1.0), (2, 60, 108, 649, 81, 1.0, 1.0),
                                               (4, 100, 120, 81, 81, 300.0, 1e-3), (4, 100, 120, 24, 96, 1e-3, 30.0)]:
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g) * scale_x
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5 * scale_w
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    want = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1))
    cp = (Cin + 7) // 8 * 8
    xn = ops.nchw_cat_to_nhwc_pad([x], cp, round_tf32=False)
    wp, bp, np_ = ops.pack_conv2d_tf32_weights(w, b, cp, split=True)
    got = ops.conv2d_tf32_nhwc_halo(xn, wp, bp, True, split=True)
    err = (got[:, 1:-1, 1:-1, :Cout].permute(0, 3, 1, 2).double() - want).abs()
    ref32 = F.relu(F.conv2d(x, w, b, padding=1))
    e32 = (ref32.double() - want).abs()
    for _ in range(3):
        ops.conv2d_tf32_nhwc_halo(xn, wp, bp, True, split=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv2d_tf32_nhwc_halo(xn, wp, bp, True, split=True)
    e1.record(); torch.cuda.synchronize()
    print(f"B{B} {H}x{W} {Cin}->{Cout} x*{scale_x:g} w*{scale_w:g}: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us  max|err| {err.max().item():.3e} "
          f"mean {err.mean().item():.3e} (cuDNN fp32: max {e32.max().item():.3e} mean {e32.mean().item():.3e}) scale {want.abs().max().item():.3g}")
