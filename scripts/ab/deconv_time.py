import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
for (B, Cin, h, w, Cout) in ((8, 216, 20, 36, 8), (8, 72, 60, 108, 8), (8, 24, 180, 324, 8), (16, 72, 60, 108, 24)):
    x = torch.randn(B, Cin, h, w, device="cuda", generator=g)
    wt = torch.randn(Cin, Cout, 3, 3, device="cuda", generator=g) * 0.1
    b = torch.zeros(Cout, device="cuda")
    for _ in range(3):
        ops.deconv3x3s3(x, wt, b, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.deconv3x3s3(x, wt, b, True)
    e1.record(); torch.cuda.synchronize()
    print(f"deconv3x3s3 B={B} {Cin}->{Cout} {h}x{w}: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us")
