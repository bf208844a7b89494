import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
for (B, C, H, W, D) in ((8, 216, 20, 36, 8), (8, 216, 14, 47, 8), (1, 216, 74, 111, 29)):
    L = torch.randn(B, C, H, W, device="cuda", generator=g)
    R = torch.randn(B, C, H, W, device="cuda", generator=g)
    for _ in range(3):
        ops.cost_volume_bf16_ndhwc(L, R, D, 224)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.cost_volume_bf16_ndhwc(L, R, D, 224)
    e1.record(); torch.cuda.synchronize()
    print(f"costvol bf16 B={B} C={C} {H}x{W} D={D}: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us")
