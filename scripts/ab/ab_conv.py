"""A/B: round-1 library vs current library, same launches (plain TF32 entry points), CUDA events."""
import ctypes as C
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from decnet_b200 import ops  # noqa

# A/B of two builds of the library: put the older build at scripts/ab/libdecnet_old.so first, e.g.
#   git stash; python -m decnet_b200.build; cp decnet_b200/libdecnet_b200.so scripts/ab/libdecnet_old.so; git stash pop; python -m decnet_b200.build
# (git-ignored; it travels to the GPU box with the snapshot).  The packed-weight formats must match the old build's.
if not (ROOT / "scripts/ab/libdecnet_old.so").exists():
    sys.exit("scripts/ab/libdecnet_old.so is missing: build the revision to compare against first (see the comment above)")
old = C.CDLL(str(ROOT / "scripts/ab/libdecnet_old.so"))
new = C.CDLL(str(ROOT / "decnet_b200/libdecnet_b200.so"))
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B, H, W = 8, 540, 972
x = torch.randn(B, 8, H, W, device=dev, generator=g)
w = torch.randn(8, 8, 3, 3, device=dev, generator=g) * 0.1
wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, torch.zeros(8, device=dev))
out = torch.empty(B, 8, H, W, device=dev)
st = torch.cuda.current_stream().cuda_stream
vp = C.c_void_p


def run(lib):
    f = lib.decnet_conv2d_tf32_nchw
    f.argtypes = [vp] * 4 + [C.c_int] * 7 + [vp]
    rc = f(x.data_ptr(), wp.data_ptr(), bp.data_ptr(), out.data_ptr(), B, 8, 8, H, W, 1, 1, st)
    assert rc == 0, rc


def timeit(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for rep in range(3):
    print("conv2d nchw 8->8: old %.1f us  new %.1f us" % (timeit(lambda: run(old)), timeit(lambda: run(new))), flush=True)

# halo kernel 81->81 at 180x324
cp = 88
xh = torch.zeros(B, 182, 326, cp, device=dev)
xh[:, 1:-1, 1:-1, :81] = torch.randn(B, 180, 324, 81, device=dev, generator=g)
xh = ops.rna_tf32(xh)
wh, bh, np_ = ops.pack_conv2d_tf32_weights(torch.randn(81, 81, 3, 3, device=dev, generator=g) * 0.05, torch.zeros(81, device=dev), cp)
oh = torch.empty(B, 182, 326, np_, device=dev)


def runh(lib):
    f = lib.decnet_conv2d_tf32_nhwc_halo
    f.argtypes = [vp] * 4 + [C.c_int] * 7 + [vp]
    rc = f(xh.data_ptr(), wh.data_ptr(), bh.data_ptr(), oh.data_ptr(), B, 180, 324, cp, np_, 1, 1, st)
    assert rc == 0, rc


for rep in range(3):
    print("halo 81->81: old %.1f us  new %.1f us" % (timeit(lambda: runh(old)), timeit(lambda: runh(new))), flush=True)
