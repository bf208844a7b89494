import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
for (B, C, H, W) in ((8, 8, 540, 972), (8, 24, 180, 324), (8, 72, 60, 108)):
    R = torch.randn(B, C, H, W, device="cuda", generator=g)
    L = torch.randn(B, C, H, W, device="cuda", generator=g)
    d = torch.rand(B, H, W, device="cuda", generator=g) * (W / 5)
    for name, fn in (("warp", lambda: ops.warp_bilinear(R, d)), ("refine_pack", lambda: ops.refine_pack(L, R, d))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name} C={C} {H}x{W}: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us")
