"""Experiment: device time of one conv3d layer vs padded output width N (CUDA-graph replay, no host
launch path in the timed region) + host time per eager call."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import torch
from decnet_b200 import conv3d as c3
B, D, H, W, cp = 8, 8, 20, 36, 224
x = torch.randn(B, D, H, W, cp, device="cuda").to(torch.bfloat16)
for np_ in (16, 64, 112, 224, 256):
    w = (torch.randn(27, np_, cp, device="cuda") * 0.01).to(torch.bfloat16)
    bias = torch.zeros(np_, device="cuda")
    out = torch.empty(B, D, H, W, np_, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): c3.conv3d_layer(x, w, bias, np_, True, out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): c3.conv3d_layer(x, w, bias, np_, True, out=out)
    host_us = (time.perf_counter() - t0) / 20 * 1e6
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(10): c3.conv3d_layer(x, w, bias, np_, True, out=out)
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"np={np_:4d}  device {us:8.1f} us/layer  ({us*1e3/324:6.1f} ns/stage)  host issue {host_us:6.1f} us/call  "
          f"TF/s (padded) {2*B*D*H*W*27*cp*np_/us/1e6:8.1f}", flush=True)
