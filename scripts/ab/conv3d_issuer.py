"""Issuer-side timing of conv3d_tcgen05_kernel (decnet_conv3d_debug_timing): cycles per 64-channel stage, share spent waiting for
operands, per CTA (CTAs with 2 whole tiles + a half item against CTAs with 2 whole tiles)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import conv3d as c3, _lib
B, D, H, W, cp = 8, 8, 20, 36, 224
x = torch.randn(B, D, H, W, cp, device="cuda").to(torch.bfloat16)
dbg = torch.zeros(148 * 4, dtype=torch.int64, device="cuda")
np_ = 224
w = (torch.randn(27, np_, cp, device="cuda") * 0.01).to(torch.bfloat16)
bias = torch.zeros(np_, device="cuda")
out = torch.empty(B, D, H, W, np_, device="cuda", dtype=torch.bfloat16)
for _ in range(5):
    c3.conv3d_layer(x, w, bias, np_, True, out=out)
torch.cuda.synchronize()
_lib.lib().decnet_conv3d_debug_timing(dbg.data_ptr())
c3.conv3d_layer(x, w, bias, np_, True, out=out)
torch.cuda.synchronize()
_lib.lib().decnet_conv3d_debug_timing(None)
d = dbg.view(148, 4).cpu().double()
for its in sorted(set(d[:, 3].tolist())):
    q = d[d[:, 3] == its]
    print(f"{len(q):3d} CTAs with {int(its)} stages: alive {q[:,0].mean():9.0f} cycles = {q[:,0].mean()/its:6.1f} per stage, waiting for operands "
          f"{100*q[:,1].mean()/q[:,0].mean():4.1f} %, {q[:,2].mean()/1e3:6.1f} us")
