import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import torch
from decnet_b200 import ops, _lib
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B, H, W = 8, 540, 972
for Cin, Cout, dil in ((8, 8, 1), (17, 8, 3), (8, 4, 1), (8, 3, 1), (4, 1, 1), (12, 8, 1)):
    x = torch.randn(B, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.1
    b = torch.randn(Cout, device="cuda")
    wp = ops.pack_conv2d_weights(w)
    _lib.lib().decnet_conv2d_set_variant(variant)
    for _ in range(3): ops.conv2d_small(x, wp, b, Cout, 3, dil, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.conv2d_small(x, wp, b, Cout, 3, dil, True)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    fl = 2.0 * B * H * W * Cin * Cout * 9
    by = 4.0 * B * H * W * (Cin + Cout)
    print(f"variant {variant} {Cin:3d}->{Cout:2d} dil {dil}: {us:7.1f} us  {fl/us/1e6:6.2f} TFLOP/s  {by/us/1e3:7.1f} GB/s")
