"""conv2d_tcgen05_kernel (thin NCHW 3x3 convs) on the layer shapes of the hot path, 3xTF32 and TF32: us per launch and the
fraction of the fp32 HBM floor (read Cin + write Cout channels once)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import ops

B = 8
torch.manual_seed(0)
rows = []
for (H, W, srcs, Cout, dil) in [(540, 972, [8], 8, 1), (540, 972, [8, 8, 1], 8, 1), (540, 972, [8], 8, 2), (540, 972, [8], 8, 4),
                                (540, 972, [8, 5], 8, 1), (540, 972, [8], 1, 1), (180, 324, [24], 24, 1), (180, 324, [24, 24, 1], 24, 1),
                                (180, 324, [24], 24, 3), (180, 324, [24], 8, 1), (60, 108, [72], 72, 1)]:
    xs = [torch.randn(B, c, H, W, device="cuda") for c in srcs]
    cin = sum(srcs)
    w = torch.randn(Cout, cin, 3, 3, device="cuda") * 0.1
    b = torch.randn(Cout, device="cuda")
    line = f"{H}x{W} {'+'.join(map(str, srcs)):>7s}->{Cout:2d} d{dil}:"
    for split in (True, False):
        if not ops.conv2d_tf32_supported(cin, Cout, H, W, dil, split):
            line += "   unsupported"
            continue
        wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b, srcs, split=split)[:2]
        for _ in range(3):
            ops.conv2d_tf32_nchw_cat(xs, wp, bp, Cout, dil, True, split=split)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv2d_tf32_nchw_cat(xs, wp, bp, Cout, dil, True, split=split)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        by = 4.0 * B * H * W * (cin + Cout)
        line += f"  {'3x' if split else 'tf32'} {us:7.1f} us ({by / us / 1e3 / 6550:4.2f} of HBM)"
    print(line)
