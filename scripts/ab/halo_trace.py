"""clock64 timeline of CTA 0 of conv2d_nhwc_halo_kernel (81 -> 81, 180x324, B = 8): where a stage's time goes.  argv[1] = mask."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import _lib, ops

mask = int(sys.argv[1]) if len(sys.argv) > 1 else 0
g = torch.Generator(device="cuda").manual_seed(0)
B, h, w, cin, cout = 8, 180, 324, 81, 81
cp = 88
x = torch.zeros(B, h + 2, w + 2, cp, device="cuda")
x[:, 1:-1, 1:-1, :cin] = torch.randn(B, h, w, cin, device="cuda", generator=g)
wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.05
wp, bp, np_ = ops.pack_conv2d_tf32_weights(wt, torch.zeros(cout, device="cuda"), cp, split=True)
for _ in range(3):
    ops.conv2d_tf32_nhwc_halo(x, wp, bp, True, split=True)
tr = torch.zeros(2560, dtype=torch.int64, device="cuda")
_lib.lib().decnet_conv2d_nhwc_set_variant(100 + mask)
_lib.lib().decnet_conv2d_nhwc_debug_trace(tr.data_ptr())
ops.conv2d_tf32_nhwc_halo(x, wp, bp, True, split=True)
torch.cuda.synchronize()
_lib.lib().decnet_conv2d_nhwc_debug_trace(None)
_lib.lib().decnet_conv2d_nhwc_set_variant(0)
t = tr.cpu()
st = t[:2048].view(256, 8)
dr = t[2048:].view(256, 2)
t0 = int(st[0, 0])
print(f"split kind {ops.SPLIT_KIND} mask {mask}; cycles relative to the first TMA issue")
print("stage  issue   landed  conv'd  seen    mma_out | fill  conv  wait  issue | period")
for i in range(36, 72):
    a = [int(v) - t0 for v in st[i, :5]]
    prev = int(st[i - 1, 4]) - t0
    print(f"{i:4d} {a[0]:7d} {a[1]:7d} {a[2]:7d} {a[3]:7d} {a[4]:7d} | {a[1]-a[0]:5d} {a[2]-a[1]:5d} {a[3]-a[2]:5d} {a[4]-a[3]:5d} | {a[4]-prev:5d}")
print("drain  full   drained   dt")
for i in range(16, 28):
    a = [int(v) - t0 for v in dr[i]]
    print(f"{i:4d} {a[0]:7d} {a[1]:7d} {a[1]-a[0]:5d}")
per = (int(st[200, 4]) - int(st[20, 4])) / 180
print(f"mean period per stage over stages 20..200: {per:.0f} cycles")
