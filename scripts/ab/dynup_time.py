import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
for (B, C, h, w) in ((8, 8, 180, 324), (8, 24, 60, 108), (8, 72, 20, 36)):
    Lf = torch.randn(B, C, 3 * h, 3 * w, device="cuda", generator=g)
    disp = torch.rand(B, h, w, device="cuda", generator=g) * 50
    cp = (9 * C + 1 + 7) // 8 * 8
    logits = torch.randn(B, h + 2, w + 2, 96, device="cuda", generator=g)
    for name, fn in (("dynup_pack", lambda: ops.dynup_pack_nhwc(disp, Lf, cp, round_tf32=False, pad=True)),
                     ("dynup_glue", lambda: ops.dynup_glue_nhwc(logits, disp, pad=True))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name} C={C} coarse {h}x{w}: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us")
