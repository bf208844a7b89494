"""TF32 vs 3xTF32 ("fp32" precision) cost of the two tensor-core Conv2d kernels at the layer shapes of the SceneFlow step
(B = 8): us per launch, CUDA events, inputs larger than L2 or L2 flushed between launches."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from decnet_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


out = {"nchw": [], "nhwc_halo": []}
g = torch.Generator(device=dev).manual_seed(0)
B = 8
for (cin, cout, H, W, d) in [(8, 8, 540, 972, 1), (24, 8, 540, 972, 3), (8, 8, 540, 972, 6), (8, 4, 540, 972, 1), (16, 8, 540, 972, 1),
                             (56, 24, 180, 324, 2), (24, 24, 180, 324, 1), (24, 24, 180, 324, 4), (24, 12, 180, 324, 1), (32, 8, 180, 324, 1),
                             (36, 36, 60, 108, 1), (80, 8, 60, 108, 1)]:
    x = torch.randn(B, cin, H, W, device=dev, generator=g)
    w = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.1
    b = torch.zeros(cout, device=dev)
    rec = {"shape": [B, cin, cout, H, W, d]}
    for split in (False, True):
        if not ops.conv2d_tf32_supported(cin, cout, H, W, d, split):
            rec["3xtf32_us" if split else "tf32_us"] = None
            continue
        wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b, split=split)
        rec["3xtf32_us" if split else "tf32_us"] = round(timeit(lambda: ops.conv2d_tf32_nchw_cat([x], wp, bp, cout, d, True, split=split)), 2)
    rec["hbm_floor_us"] = round(4.0 * B * H * W * (cin + cout) / 6538e9 * 1e6, 2)
    out["nchw"].append(rec)
    print(rec, flush=True)

for (cin, cout, h, w_) in [(73, 81, 180, 324), (81, 81, 180, 324), (217, 81, 60, 108), (81, 81, 60, 108), (649, 81, 20, 36),
                           (145, 72, 60, 108), (72, 72, 60, 108), (72, 36, 60, 108)]:
    cp = (cin + 7) // 8 * 8
    x = torch.zeros(B, h + 2, w_ + 2, cp, device=dev)
    x[:, 1:-1, 1:-1, :cin] = torch.randn(B, h, w_, cin, device=dev, generator=g)
    w = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05
    b = torch.zeros(cout, device=dev)
    rec = {"shape": [B, cin, cout, h, w_]}
    for split in (False, True):
        wp, bp, np_ = ops.pack_conv2d_tf32_weights(w, b, cp, split=split)
        xx = x if split else ops.rna_tf32(x)
        rec["3xtf32_us" if split else "tf32_us"] = round(timeit(lambda: ops.conv2d_tf32_nhwc_halo(xx, wp, bp, True, split=split)), 2)
    rec["gflop"] = round(2.0 * B * h * w_ * 9 * cin * cout / 1e9, 2)
    out["nhwc_halo"].append(rec)
    print(rec, flush=True)

Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/r02_precision_cost.json").write_text(json.dumps(out, indent=1))
