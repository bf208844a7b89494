"""TEST INFRASTRUCTURE (lives under tests/ because it links the checker: oracle/_ref).
Microbench of the sparse ops (BASELINE.json config 5): ours vs the reference kernels
recompiled for sm_100a (oracle/_ref), per level / density.  CUDA-event timing, L2 flushed
between iterations.  Usage: python tests/bench_sparse_vs_reference.py [--B 8] [--iters 20]"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from decnet_b200 import ops, _lib  # noqa: E402
from helpers import make_feats, make_masks  # noqa: E402


def time_fn(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--rhos", type=float, nargs="+", default=[0.01, 0.03, 0.1, 0.3])
    ap.add_argument("--ref", type=int, default=1)
    ap.add_argument("--combos", type=lambda t: tuple(int(x) for x in t.split(",")), nargs="+",
                    default=[(2, 1), (2, 2), (0, 3), (0, 4)], help="path,variant pairs")
    ap.add_argument("--levels", nargs="+", default=["s1", "s2", "s3"])
    args = ap.parse_args()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    from oracle import ref_cuda
    have_ref = args.ref and ref_cuda.available()
    # the model's three levels, then BASELINE.json configs[4]'s other channel counts at 1/3 and full resolution
    levels = [("s1", 72, 60, 108, 24), ("s2", 24, 180, 324, 72), ("s3", 8, 540, 972, 216),
              ("c32_third", 32, 180, 324, 72), ("c64_third", 64, 180, 324, 72),
              ("c32_full", 32, 540, 972, 216), ("c64_full", 64, 540, 972, 216)]
    rows = []
    for name, C, H, W, D in levels:
        if name not in args.levels:
            continue
        L, R = make_feats(args.B, C, H, W, device="cuda")
        for rho in args.rhos:
            ml, mr = make_masks(args.B, H, W, rho, rho, device="cuda")
            abytes = 4 * args.B * H * W * (2 * C + 2 + 4)
            rec = {"level": name, "C": C, "H": H, "W": W, "D": D, "B": args.B, "rho": rho,
                   "alg_MB_fused": abytes / 1e6}
            # (path, variant): 1/2 = staged rows by cp.async / TMA, variant 1 one row per CTA, 2 persistent with
            # next-row prefetch; path 0 + variant 3 = sector-gather kernel, 4 = its software-pipelined form
            for path, variant in args.combos:
                key = f"fused_p{path}v{variant}"
                _lib.lib().decnet_set_sparse_path(path)
                _lib.lib().decnet_set_sparse_variant(variant)
                try:
                    ms = time_fn(lambda: ops.spamat_spavar_forward(L, R, ml, mr, D), args.iters, flush)
                    rec[f"{key}_us"] = round(ms * 1e3, 2)
                    rec[f"{key}_frac"] = round(abytes / (ms * 1e-3) / 1e9 / peak, 4)
                except _lib.DecnetError as e:
                    rec[f"{key}_us"] = None
                finally:
                    _lib.lib().decnet_set_sparse_path(0)
                    _lib.lib().decnet_set_sparse_variant(0)
            if have_ref:
                def ref_both():
                    o, _, _ = ref_cuda.spamat_forward(L, R, ml, mr, D, sync=False)
                    ref_cuda.spavar_forward(L, R, ml, mr, o, D, sync=False)
                torch.cuda.synchronize()
                ms = time_fn(ref_both, max(3, args.iters // 4), flush)
                rec["ref_cuda_mat+var_us"] = ms * 1e3
            rows.append(rec)
            print(json.dumps(rec), flush=True)
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "bench_sparse.json").write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
