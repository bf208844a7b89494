"""Per-role wait cycles of CTA 0 of conv2d_tcgen05_kernel (decnet_conv2d_tf32_debug): who waits for whom."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import _lib, ops

B, H, W = 8, 540, 972
torch.manual_seed(0)
for srcs, Cout, dil in (([8], 8, 1), ([8, 8, 1], 8, 1)):
    xs = [torch.randn(B, c, H, W, device="cuda") for c in srcs]
    cin = sum(srcs)
    w = torch.randn(Cout, cin, 3, 3, device="cuda") * 0.1
    b = torch.randn(Cout, device="cuda")
    for split in (True, False):
        wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b, srcs, split=split)[:2]
        for _ in range(3):
            ops.conv2d_tf32_nchw_cat(xs, wp, bp, Cout, dil, True, split=split)
        prof = torch.zeros(16, dtype=torch.int64, device="cuda")
        _lib.lib().decnet_conv2d_tf32_debug(0, prof.data_ptr())
        ops.conv2d_tf32_nchw_cat(xs, wp, bp, Cout, dil, True, split=split)
        torch.cuda.synchronize()
        _lib.lib().decnet_conv2d_tf32_debug(0, None)
        p = prof.cpu().view(4, 4)
        tiles = (B * 35 * 17 + 147) // 148
        print(f"{'+'.join(map(str, srcs))}->{Cout} {'split' if split else 'tf32 '}: alive {int(p[1,0])} cycles (~{int(p[1,0]) // tiles} per tile); "
              f"producer waits for a free stage {100 * int(p[0,1]) / int(p[0,0]):.0f} %; MMA warp waits for a TMEM slot {100 * int(p[1,1]) / int(p[1,0]):.0f} %, "
              f"for a converted stage {100 * int(p[1,2]) / int(p[1,0]):.0f} %; converter waits for a landed stage {100 * int(p[2,1]) / int(p[2,0]):.0f} %; "
              f"epilogue waits for an accumulator {100 * int(p[3,1]) / int(p[3,0]):.0f} %")
