"""Debug / timing harness for the NCHW TF32 tcgen05 Conv2d: one shape, checked against a CPU fp64 conv."""
import sys, time
from pathlib import Path
import torch
import torch.nn.functional as F
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from decnet_b200 import ops

def tf32(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)

def run(B, Cin, Cout, H, W, dil, relu=True, check=True, iters=0):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = tf32(torch.randn(B, Cin, H, W, device="cuda", generator=g))
    w = tf32(torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * Cin)) ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g) * 0.1
    wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b)
    got = ops.conv2d_tf32_nchw(x, wp, bp, Cout, dil, relu)
    torch.cuda.synchronize()
    msg = f"B={B} Cin={Cin} Cout={Cout} H={H} W={W} dil={dil}:"
    if check:
        want = F.conv2d(x.cpu().double(), w.cpu().double(), b.cpu().double(), padding=dil, dilation=dil)
        want = (F.relu(want) if relu else want).float()
        diff = (got.cpu() - want).abs()
        msg += f" max err {diff.max().item():.3e} (scale {want.abs().max().item():.2f})"
        if diff.max().item() > 1e-3:
            bad = (diff > 1e-3).nonzero()
            msg += f" BAD {bad.shape[0]}/{diff.numel()} first {bad[:4].tolist()}"
            cols = sorted(set(bad[:, 3].tolist()))[:40]; rows = sorted(set(bad[:, 2].tolist()))[:40]
            msg += f"\n   bad cols {cols}\n   bad rows {rows} chans {sorted(set(bad[:,1].tolist()))}"
    if iters:
        for _ in range(3):
            ops.conv2d_tf32_nchw(x, wp, bp, Cout, dil, relu)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.conv2d_tf32_nchw(x, wp, bp, Cout, dil, relu)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        byts = 4.0 * B * H * W * (Cin + Cout)
        msg += f"  {us:.1f} us  ({byts / us / 1e3:.0f} GB/s algorithmic)"
        if ops.conv2d_small_supported(Cin, Cout, 3):
            wq = ops.pack_conv2d_weights(w)
            for _ in range(3):
                ops.conv2d_small(x, wq, b, Cout, 3, dil, relu)
            e0.record()
            for _ in range(iters):
                ops.conv2d_small(x, wq, b, Cout, 3, dil, relu)
            e1.record(); torch.cuda.synchronize()
            msg += f"   [direct fp32 kernel {e0.elapsed_time(e1) * 1e3 / iters:.1f} us]"
    print(msg, flush=True)

if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    if mode == "l2":
        pass
    elif mode == "check":
        run(1, 8, 8, 64, 96, 1)
        run(1, 8, 8, 37, 100, 1)
        run(1, 17, 8, 50, 120, 3)
        run(1, 49, 24, 60, 108, 2)
    elif mode == "dbg":
        import ctypes
        from decnet_b200 import _lib
        f = ctypes.CDLL(str(Path(_lib.__file__).parent / "libdecnet_b200.so")).decnet_conv2d_tf32_debug
        prof = torch.zeros(16, dtype=torch.int64, device="cuda")
        f.argtypes = [ctypes.c_int, ctypes.c_void_p]
        for flags in (0, 2):
            for shape in ((8, 8, 8, 540, 972, 1),):
                f(flags, prof.data_ptr()); print("dbg flags", flags)
                run(*shape, check=False, iters=5)
                torch.cuda.synchronize()
                pr = prof.cpu().tolist()
                print("   producer: alive %d wait-empty %d | mma: alive %d wait-tmem %d wait-ready %d | conv: alive %d wait-full %d | epi: alive %d wait-acc %d"
                      % (pr[0], pr[1], pr[4], pr[5], pr[6], pr[8], pr[9], pr[12], pr[13]))
        f(0, None)
    else:
        for (B, Cin, Cout, H, W, d) in [(8, 8, 8, 540, 972, 1), (8, 17, 8, 540, 972, 3), (8, 12, 8, 540, 972, 1),
                                        (8, 8, 4, 540, 972, 1), (8, 4, 4, 540, 972, 6), (8, 4, 4, 540, 972, 9),
                                        (8, 4, 1, 540, 972, 1), (8, 8, 3, 540, 972, 1), (8, 8, 1, 540, 972, 1),
                                        (8, 24, 8, 180, 324, 1), (8, 49, 24, 180, 324, 2), (8, 24, 24, 180, 324, 1),
                                        (8, 24, 12, 180, 324, 4), (8, 28, 24, 180, 324, 1), (8, 72, 8, 60, 108, 1),
                                        (8, 36, 36, 60, 108, 1)]:
            run(B, Cin, Cout, H, W, d, check=False, iters=20)


def l2_resident_test():
    """Is a chain of 8->8 layers faster per pair when the batch chunk keeps activations in the 126 MB L2?"""
    g = torch.Generator(device="cuda").manual_seed(1)
    w = torch.randn(8, 8, 3, 3, device="cuda", generator=g) * 0.1
    b = torch.zeros(8, device="cuda")
    wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b)
    for B in (8, 4, 2, 1):
        x = torch.randn(B, 8, 540, 972, device="cuda", generator=g)
        y = torch.empty_like(x)
        bufs = [x, y]
        def chain(n=6):
            for i in range(n):
                src, dst = bufs[i & 1], bufs[(i + 1) & 1]
                ops._call("decnet_conv2d_tf32_nchw", src, src.data_ptr(), wp.data_ptr(), bp.data_ptr(), dst.data_ptr(),
                          B, 8, 8, 540, 972, 1, 1)
        chain(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            chain()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 60
        print(f"B={B}: {us:.1f} us per layer launch, {us / B:.2f} us per pair-layer ({B * 8 * 540 * 972 * 8 / us / 1e3:.0f} GB/s)")


if len(sys.argv) > 1 and sys.argv[1] == "l2":
    l2_resident_test()
