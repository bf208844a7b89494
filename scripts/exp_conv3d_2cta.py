"""1-CTA vs CTA-pair aggregation kernel: layer time and the issuer's cycle split (alive / blocked on operands / blocked on TMEM)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import torch
from decnet_b200 import conv3d as c3, _lib
B, D, H, W, cp, np_ = 8, 8, 20, 36, 224, 224
x = torch.randn(B, D, H, W, cp, device="cuda").to(torch.bfloat16)
w = (torch.randn(27, np_, cp, device="cuda") * 0.01).to(torch.bfloat16)
bias = torch.zeros(np_, device="cuda")
out = torch.empty(B, D, H, W, np_, device="cuda", dtype=torch.bfloat16)
outs = {}
dbg = torch.zeros(148 * 4, dtype=torch.int64, device="cuda")
for variant in (1, 2, 12, 22, 32):
    _lib.lib().decnet_conv3d_set_variant(variant)
    for _ in range(5):
        c3.conv3d_layer(x, w, bias, np_, True, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        c3.conv3d_layer(x, w, bias, np_, True, out=out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 50
    dbg.zero_()
    _lib.lib().decnet_conv3d_debug_timing(dbg.data_ptr())
    c3.conv3d_layer(x, w, bias, np_, True, out=out)
    torch.cuda.synchronize()
    _lib.lib().decnet_conv3d_debug_timing(None)
    outs[variant] = out.clone()
    d = dbg.view(148, 4).cpu().double()
    act = d[d[:, 3] > 0]
    print(f"variant {variant}: {us:.1f} us/layer; issuing CTAs {act.shape[0]}; stages/issuer {act[:,3].mean():.0f}; cycles/stage {(act[:,0] / act[:,3]).mean():.0f}; "
          f"col1/stage {(act[:,1] / act[:,3]).mean():.0f}; col2 {(act[:,2]).mean():.0f}")
_lib.lib().decnet_conv3d_set_variant(0)
print('pair == single:', torch.equal(outs[1], outs[2]), '(variants 12/22/32: pair kernel without A / B / both loads, timing only)')
