"""Per-level timing of the glue kernels (refine_pack / deconv / dynup pack+glue / attn_pack) at SceneFlow size, B=8,
with the bytes each must move, to see which are far from the HBM roofline."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from decnet_b200 import ops

def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters

B = 8
levels = {1: (72, 60, 108), 2: (24, 180, 324), 3: (8, 540, 972)}
g = torch.Generator(device="cuda").manual_seed(0)
for s, (C, H, W) in levels.items():
    L = torch.randn(B, C, H, W, device="cuda", generator=g)
    R = torch.randn(B, C, H, W, device="cuda", generator=g)
    disp = torch.rand(B, H, W, device="cuda", generator=g) * (W / 5)
    n = B * H * W * 4
    us = timeit(lambda: ops.refine_pack(L, R, disp))
    print(f"s{s} refine_pack C={C}: {us:.1f} us  ({(n * (4 * C + 2)) / us / 1e3:.0f} GB/s of {n * (4 * C + 2) / 1e6:.0f} MB)")
    m = torch.rand(B, H, W, device="cuda", generator=g)
    us = timeit(lambda: ops.attn_pack(L, disp, disp, m, m))
    print(f"s{s} attn_pack   C={C}: {us:.1f} us  ({(n * (2 * C + 8)) / us / 1e3:.0f} GB/s of {n * (2 * C + 8) / 1e6:.0f} MB)")
    # deconv: previous level features [B, 3C, H/3, W/3] -> [B, 8, H, W]
    xp = torch.randn(B, 3 * C, H // 3, W // 3, device="cuda", generator=g)
    wd = torch.randn(3 * C, 8, 3, 3, device="cuda", generator=g) * 0.1
    bd = torch.zeros(8, device="cuda")
    us = timeit(lambda: ops.deconv3x3s3(xp, wd, bd, True))
    byts = xp.numel() * 4 + B * 8 * H * W * 4
    print(f"s{s} deconv3x3s3 Cin={3 * C}: {us:.1f} us  ({byts / us / 1e3:.0f} GB/s of {byts / 1e6:.0f} MB)")
    dprev = torch.rand(B, H // 3, W // 3, device="cuda", generator=g) * 10
    cp = (9 * C + 1 + 7) // 8 * 8
    us = timeit(lambda: ops.dynup_pack_nhwc(dprev, L, cp))
    byts = L.numel() * 4 + B * (H // 3) * (W // 3) * cp * 4
    print(f"s{s} dynup_pack_nhwc C={C} cp={cp}: {us:.1f} us  ({byts / us / 1e3:.0f} GB/s of {byts / 1e6:.0f} MB)")
    lg = torch.randn(B, H // 3, W // 3, 96, device="cuda", generator=g)
    us = timeit(lambda: ops.dynup_glue_nhwc(lg, dprev))
    byts = lg.numel() * 4 + n
    print(f"s{s} dynup_glue_nhwc: {us:.1f} us  ({byts / us / 1e3:.0f} GB/s of {byts / 1e6:.0f} MB)")
