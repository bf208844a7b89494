"""Where does the TF32 NHWC conv (DynamicUpsampling.weight_learning) spend its time?  Per-CTA issuer cycles vs
cycles blocked on operand (full) barriers, for the three SceneFlow levels at B=8."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import torch
from decnet_b200 import ops, _lib
B = 8
dbg = torch.zeros(148 * 4, dtype=torch.int64, device="cuda")
import ctypes
for variant, (h, w, cin) in [(v, s) for v in (0,) for s in ((180, 324, 73), (60, 108, 217))]:
    _lib.lib().decnet_conv3d_set_variant(variant); print("variant", variant, "(3 = no TMA refills: stale operands, timing only)")
    for ci, co in ((cin, 81), (81, 81)):
        cp = (ci + 7) // 8 * 8
        x = torch.randn(B, h, w, cp, device="cuda")
        wt = torch.randn(co, ci, 3, 3, device="cuda") * 0.05
        wp, bp, np_ = ops.pack_conv2d_tf32_weights(wt, torch.zeros(co, device="cuda"), cp)
        for _ in range(3):
            ops.conv2d_tf32_nhwc(x, wp, bp, True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.conv2d_tf32_nhwc(x, wp, bp, True)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        _lib.lib().decnet_conv3d_debug_timing(dbg.data_ptr())
        ops.conv2d_tf32_nhwc(x, wp, bp, True)
        torch.cuda.synchronize()
        _lib.lib().decnet_conv3d_debug_timing(None)
        d = dbg.view(148, 4).cpu().double()
        stages = d[:, 3].mean()
        flops = 2.0 * B * h * w * 9 * cp * np_
        print(f"{h}x{w} cp={cp} np={np_}: {us:7.1f} us  {flops / us / 1e6:6.1f} TF/s(padded)  stages/CTA {stages:6.0f}  "
              f"cyc/stage {d[:,0].mean() / stages:6.1f}  blocked-on-operands/stage {d[:,1].mean() / stages:6.1f}")
