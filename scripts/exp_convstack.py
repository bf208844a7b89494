"""Experiment: cuDNN 2-D conv stacks of the hot path under NCHW vs channels_last, plain vs fused bias+relu."""
import sys, time
from pathlib import Path
import torch
import torch.nn.functional as F
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = True
dev = "cuda"

def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def stack(chain, dil, B, H, W, cl, fused, dtype=torch.float32):
    g = torch.Generator(device=dev).manual_seed(0)
    ws = []
    for (ci, co), d in zip(chain, dil):
        w = torch.randn(co, ci, 3, 3, device=dev, generator=g).to(dtype) * 0.1
        b = torch.randn(co, device=dev, generator=g).to(dtype) * 0.1
        if cl: w = w.contiguous(memory_format=torch.channels_last)
        ws.append((w, b, d))
    x = torch.randn(B, chain[0][0], H, W, device=dev, generator=g).to(dtype)
    if cl: x = x.contiguous(memory_format=torch.channels_last)
    def run():
        y = x
        for i, (w, b, d) in enumerate(ws):
            last = i == len(ws) - 1
            if fused and not last:
                y = torch.cudnn_convolution_relu(y, w, b, (1, 1), (d, d), (d, d), 1)
            else:
                y = F.conv2d(y, w, b, padding=d, dilation=d)
                if not last: y = F.relu_(y)
        return y
    return run

B = 8
cases = {
  "refine_s3 (17->8..->1, 540x972)": ([(17,8),(8,8),(8,8),(8,4),(4,4),(4,4),(4,1)], (3,1,6,1,9,1,1), 540, 972),
  "refine_s2 (49->24..->1, 180x324)": ([(49,24),(24,24),(24,24),(24,12),(12,12),(12,12),(12,1)], (2,1,4,1,6,1,1), 180, 324),
  "attn_s3 (12->8->8->1, 540x972)": ([(12,8),(8,8),(8,1)], (1,1,1), 540, 972),
  "dynup_s3 (73->81->81->81, 180x324)": ([(73,81),(81,81),(81,81)], (1,1,1), 180, 324),
  "dynup_s1 (649->81->81->81, 20x36)": ([(649,81),(81,81),(81,81)], (1,1,1), 20, 36),
}
for name, (chain, dil, H, W) in cases.items():
    for cl in (False, True):
        for fused in (False, True):
            for dtype in (torch.float32, torch.bfloat16):
                try:
                    t = bench(stack(chain, dil, B, H, W, cl, fused, dtype))
                    print(f"{name:38s} channels_last={cl!s:5s} fused={fused!s:5s} {str(dtype)[6:]:9s} {t*1e3:9.1f} us", flush=True)
                except Exception as e:
                    print(f"{name:38s} channels_last={cl!s:5s} fused={fused!s:5s} {str(dtype)[6:]:9s} FAILED {type(e).__name__}: {str(e)[:80]}", flush=True)
