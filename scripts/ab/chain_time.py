"""Launch gaps: a CUDA graph of N chained launches of one kernel against N x its isolated duration.
argv[1] in {thin, halo, conv3d}; DECNET_PDL=0 disables programmatic dependent launch."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import ops, conv3d as c3
which = sys.argv[1] if len(sys.argv) > 1 else "thin"
N = 16
g = torch.Generator(device="cuda").manual_seed(0)
if which == "thin":
    shapes = [(8, 8, 540, 972), (8, 24, 180, 324), (8, 36, 60, 108), (8, 8, 60, 108)]
else:
    shapes = [None]
for shp in shapes:
    if which == "thin":
        B, C, H, W = shp
        x = torch.randn(B, C, H, W, device="cuda", generator=g)
        w = torch.randn(C, C, 3, 3, device="cuda", generator=g) * (1.0 / (9 * C)) ** 0.5
        wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, torch.zeros(C, device="cuda"), split=True)
        step = lambda t: ops.conv2d_tf32_nchw_cat([t], wp, bp, C, 1, True, split=True)
    elif which == "halo":
        B, h, w_, cin = 8, 180, 324, 81
        cp = 88
        x = torch.zeros(B, h + 2, w_ + 2, 96, device="cuda")
        x[:, 1:-1, 1:-1, :cin] = torch.randn(B, h, w_, cin, device="cuda", generator=g)
        wt = torch.randn(81, 96, 3, 3, device="cuda", generator=g) * 0.03
        wp, bp, np_ = ops.pack_conv2d_tf32_weights(wt, torch.zeros(81, device="cuda"), 96, split=True)
        step = lambda t: ops.conv2d_tf32_nhwc_halo(t, wp, bp, True, split=True)
    else:
        B, D, H, W, C = 8, 8, 20, 36, 224
        x = torch.randn(B, D, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
        wt = (torch.randn(27, C, C, device="cuda", generator=g) * 0.01).to(torch.bfloat16)
        bias = torch.zeros(C, device="cuda")
        step = lambda t: c3.conv3d_layer(t, wt, bias, C, True)

    def chain():
        t = x
        for _ in range(N):
            t = step(t)
        return t
    for _ in range(2):
        chain()
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        chain()
        with torch.cuda.graph(gr, stream=s):
            chain()
    torch.cuda.current_stream().wait_stream(s)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 10 / N * 1e3
    print(f"{which} {shp}: {per:7.2f} us per launch inside a graph of {N} chained launches")
