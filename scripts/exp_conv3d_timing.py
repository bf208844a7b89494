import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import torch
from decnet_b200 import conv3d as c3, _lib
B, D, H, W, cp = 8, 8, 20, 36, 224
x = torch.randn(B, D, H, W, cp, device="cuda").to(torch.bfloat16)
dbg = torch.zeros(148 * 4, dtype=torch.int64, device="cuda")
for np_ in (16, 224):
    w = (torch.randn(27, np_, cp, device="cuda") * 0.01).to(torch.bfloat16)
    bias = torch.zeros(np_, device="cuda")
    out = torch.empty(B, D, H, W, np_, device="cuda", dtype=torch.bfloat16)
    for _ in range(5): c3.conv3d_layer(x, w, bias, np_, True, out=out)
    torch.cuda.synchronize()
    _lib.lib().decnet_conv3d_debug_timing(dbg.data_ptr())
    c3.conv3d_layer(x, w, bias, np_, True, out=out)
    torch.cuda.synchronize()
    _lib.lib().decnet_conv3d_debug_timing(None)
    d = dbg.view(148, 4).cpu().double()
    three = d[:, 3] == 324
    q = d[three]
    print(f"np={np_:3d} 3-unit CTAs n={int(three.sum())}: {q[:,0].mean()/324:6.1f} cyc/stage, wait(full) {q[:,1].mean()/324:6.1f}, wait(tmem_empty) total {q[:,2].mean():8.0f} cyc")
