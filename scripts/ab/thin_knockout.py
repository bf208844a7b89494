"""Knock-out timing of conv2d_tcgen05_kernel (tuning only, results wrong with a mask set): 1 converters idle, 2 epilogue skips
its math and stores, 4 no MMAs."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import _lib, ops

B, H, W = 8, 540, 972
torch.manual_seed(0)
for srcs, Cout, dil in (([8], 8, 1), ([8, 8, 1], 8, 1), ([8], 8, 4)):
    xs = [torch.randn(B, c, H, W, device="cuda") for c in srcs]
    cin = sum(srcs)
    w = torch.randn(Cout, cin, 3, 3, device="cuda") * 0.1
    b = torch.randn(Cout, device="cuda")
    for split in (True, False):
        wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b, srcs, split=split)[:2]
        line = f"{'+'.join(map(str, srcs))}->{Cout} d{dil} {'split' if split else 'tf32 '}:"
        for mask in (0, 1, 2, 4, 3, 5, 6, 7):
            _lib.lib().decnet_conv2d_tf32_debug(mask, None)
            for _ in range(3):
                ops.conv2d_tf32_nchw_cat(xs, wp, bp, Cout, dil, True, split=split)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.conv2d_tf32_nchw_cat(xs, wp, bp, Cout, dil, True, split=split)
            e1.record(); torch.cuda.synchronize()
            line += f"  m{mask} {e0.elapsed_time(e1) / 20 * 1e3:6.1f}"
        _lib.lib().decnet_conv2d_tf32_debug(0, None)
        print(line)
