"""conv2d_nhwc_halo (padded channels-last, one A tile per row tap) vs the per-tap kernel: equality and time."""
import sys
from pathlib import Path
import torch
import torch.nn.functional as F
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from decnet_b200 import ops

def run(B, h, w, ci, co, iters=0):
    g = torch.Generator(device="cuda").manual_seed(1)
    cp = (ci + 7) // 8 * 8
    x = torch.zeros(B, h, w, cp, device="cuda")
    x[..., :ci] = torch.randn(B, h, w, ci, device="cuda", generator=g)
    x = ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    wt = torch.randn(co, ci, 3, 3, device="cuda", generator=g) * (2.0 / (9 * ci)) ** 0.5
    bias = torch.randn(co, device="cuda", generator=g) * 0.1
    wp, bp, np_ = ops.pack_conv2d_tf32_weights(wt, bias, cp)
    want = ops.conv2d_tf32_nhwc(x, wp, bp, True)
    xp = F.pad(x, (0, 0, 1, 1, 1, 1)).contiguous()
    got = ops.conv2d_tf32_nhwc_halo(xp, wp, bp, True)
    torch.cuda.synchronize()
    border = torch.cat([got[:, 0].flatten(), got[:, -1].flatten(), got[:, :, 0].flatten(), got[:, :, -1].flatten()]).abs().max().item()
    err = (got[:, 1:-1, 1:-1] - want).abs().max().item()
    msg = f"B={B} {h}x{w} {ci}->{co}: max |halo - per-tap| {err:.2e} (scale {want.abs().max().item():.2f}), border max {border:.1e}"
    if iters:
        def t(fn):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / iters
        msg += f"   halo {t(lambda: ops.conv2d_tf32_nhwc_halo(xp, wp, bp, True)):.1f} us   per-tap {t(lambda: ops.conv2d_tf32_nhwc(x, wp, bp, True)):.1f} us"
    print(msg, flush=True)

run(1, 9, 14, 16, 16)
run(2, 20, 36, 73, 81)
run(8, 180, 324, 73, 81, iters=10)
run(8, 180, 324, 81, 81, iters=10)
run(8, 60, 108, 217, 81, iters=10)
run(8, 20, 36, 649, 81, iters=10)
if len(sys.argv) > 1:
    for shp in [(2, 13, 37, 73, 81), (1, 13, 37, 73, 81), (2, 13, 36, 73, 81), (2, 13, 37, 32, 16), (2, 12, 37, 73, 81)]:
        run(*shp)
