"""Check / time the rows formulation (conv2d_rows_tcgen05.cu) against a CPU fp64 conv and against conv2d_tcgen05."""
import sys
from pathlib import Path
import torch
import torch.nn.functional as F
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from decnet_b200 import ops

def tf32(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)

def run(B, chans, Cout, H, W, dil, relu=True, check=True, iters=0):
    g = torch.Generator(device="cuda").manual_seed(1)
    srcs = [tf32(torch.randn(B, c, H, W, device="cuda", generator=g)) for c in chans]
    cin = sum(chans)
    w = tf32(torch.randn(Cout, cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    wc, b8 = ops.pack_conv2d_tf32_rows_weights(w, b, chans)
    got = ops.conv2d_tf32_rows_nchw_cat(srcs, wc, b8, Cout, dil, relu)
    torch.cuda.synchronize()
    msg = f"B={B} chans={chans} Cout={Cout} H={H} W={W} dil={dil}:"
    if check:
        x = torch.cat(srcs, 1)
        want = F.conv2d(x.cpu().double(), w.cpu().double(), b.cpu().double(), padding=dil, dilation=dil)
        want = (F.relu(want) if relu else want).float()
        diff = (got.cpu() - want).abs()
        msg += f" max err {diff.max().item():.3e} (scale {want.abs().max().item():.2f})"
        if diff.max().item() > 1e-3:
            bad = (diff > 1e-3).nonzero()
            msg += f" BAD {bad.shape[0]}/{diff.numel()} first {bad[:3].tolist()}"
            msg += f"\n   bad cols {sorted(set(bad[:, 3].tolist()))[:30]}\n   bad rows {sorted(set(bad[:, 2].tolist()))[:30]} chans {sorted(set(bad[:,1].tolist()))}"
    if iters:
        def t(fn):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / iters
        us = t(lambda: ops.conv2d_tf32_rows_nchw_cat(srcs, wc, b8, Cout, dil, relu))
        wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b, chans)
        us1 = t(lambda: ops.conv2d_tf32_nchw_cat(srcs, wp, bp, Cout, dil, relu))
        byts = 4.0 * B * H * W * (cin + Cout)
        msg += f"  rows {us:.1f} us ({byts / us / 1e3:.0f} GB/s)   pixels-on-M kernel {us1:.1f} us"
    print(msg, flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "dbg":
        import ctypes
        from decnet_b200 import _lib
        f = ctypes.CDLL(str(Path(_lib.__file__).parent / "libdecnet_b200.so")).decnet_conv2d_tf32_rows_debug
        for flags in (0, 1, 2, 4, 3, 6, 7):
            f(flags); print("dbg flags", flags, "(1 no rounding, 2 no stores, 4 no MMAs)")
            run(8, (8,), 8, 540, 972, 1, check=False, iters=10)
            run(8, (8, 8, 1), 8, 540, 972, 3, check=False, iters=10)
        f(0)
    elif len(sys.argv) > 1 and sys.argv[1] == "time":
        for args in [(8, (8,), 8, 540, 972, 1), (8, (8, 4), 8, 540, 972, 1), (8, (8, 8, 1), 8, 540, 972, 3), (8, (8,), 4, 540, 972, 1),
                     (8, (8,), 3, 540, 972, 1), (8, (8,), 1, 540, 972, 1), (8, (3,), 3, 540, 972, 1), (8, (4,), 4, 540, 972, 1),
                     (8, (24,), 8, 180, 324, 1), (16, (3,), 8, 540, 972, 1), (16, (8, 8), 8, 540, 972, 1)]:
            run(*args, check=False, iters=20)
    else:
        run(1, (8,), 8, 64, 96, 1)
        run(2, (8,), 8, 37, 100, 1)
        run(1, (8, 8, 1), 8, 50, 120, 3)
        run(1, (8, 4), 8, 45, 200, 1)
        run(1, (5,), 3, 33, 60, 2, relu=False)
        run(1, (24,), 8, 60, 108, 1)
