#!/usr/bin/env python
"""bench.py -- decomposed-matching hot path throughput (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # reference arm: CPU port on host cores

One "step" = one pass of the hot path (SparseDenseNetRefinementMask.forward after feature
extraction: cost volume -> 3-D aggregation -> soft-argmin, then per level detail masks ->
dynamic up-sampling -> SpaMat/SpaVar -> soft attention + blend -> refinement) over one batch of
synthetic SceneFlow-shaped feature pyramids (540x960 -> 540x972, max_disp "192" -> 216), random-init
weights.  N=1 workload = BASELINE.json configs[1] (batch 8 on one B200); N>1 shards by stereo pair,
no collective on the data path (weak scaling).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "stereo pairs/s @540x960 (decomposed-matching hot path, SceneFlow shape, max_disp 216)"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sceneflow")
    ap.add_argument("--batch", type=int, default=8, help="stereo pairs per GPU per step")
    ap.add_argument("--rho", type=float, default=0.10, help="calibrated lost-detail mask density")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32"],
                    help="arithmetic mode of the 2-D tensor-core convs: fp32 = error-compensated 3xTF32 (parity-gated default), "
                         "tf32 = plain TF32 (what the reference gets from cuDNN on a GPU)")
    ap.add_argument("--mode", default="pairs", choices=["pairs", "bands"],
                    help="pairs: shard by stereo pair (weak scaling); bands: ONE pair split into row bands "
                         "with halo exchange over NCCL (strong scaling, BASELINE.json configs[3])")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-overlap", action="store_true",
                    help="masks + sparse ops on the main stream instead of a forked second stream")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-from-images", action="store_true",
                    help="skip the extra leg that starts from images (feature extractor + hot path, SURVEY.md 8f rank 2)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="host time of the bounded CPU sample")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------
# CPU port (oracle) -- used ONLY as the reported cpu_baseline and as the --impl reference arm
# ------------------------------------------------------------------------------------------
def _calibrate_cpu_params(P, left, right, rho, thold):
    """Same one-scalar-per-level calibration as decnet_b200.synthetic.calibrate_mask_density, on the oracle's
    parameter dict: shift the detectors' last BN bias so the learned masks have density `rho`."""
    import math
    import torch
    from oracle import glue as og
    logit_t = math.log(thold / (1.0 - thold))
    pre = left["stage0"]
    for l in range(3):
        cur = left[f"stage{l + 1}"]
        with torch.no_grad():
            lg = og.detail_logits(cur, pre, P, f"detail_detection.{l}")
            q = torch.quantile(lg.flatten()[:: max(1, lg.numel() // 2_000_000)], 1.0 - rho)
        P[f"detail_detection.{l}.conv.1.bn.bias"] = P[f"detail_detection.{l}.conv.1.bn.bias"] + (logit_t - q)
        pre = cur


def cpu_port_pairs_per_s(workload, rho, min_seconds=10.0, max_pairs=64, warmup=1):
    """The oracle pipeline (torch-CPU dense/glue + C/OpenMP SpaMat/SpaVar) on all host cores, the same
    configuration as the CUDA arm (learned detectors calibrated to `rho`), one pair per call, repeated
    until `min_seconds` of CPU work.  Returns (pairs/s, cores, seconds, pairs)."""
    import torch
    from decnet_b200.params import make_features, make_hotpath_state
    from decnet_b200.synthetic import WORKLOADS
    from oracle import pipeline as opipe
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    H, W, max_disp, skip = WORKLOADS[workload]
    P = make_hotpath_state(17)
    left, right = make_features(1, H, W, seed=17)
    _calibrate_cpu_params(P, left, right, rho, 0.9)

    def one_pair():
        with torch.no_grad():
            opipe.forward(P, left, right, max_disp, use_detail=True, thold=0.9, skip_stage_id=skip)
    for _ in range(warmup):
        one_pair()
    n, t0 = 0, time.perf_counter()
    while True:
        one_pair()
        n += 1
        el = time.perf_counter() - t0
        if el >= min_seconds or n >= max_pairs:
            break
    return n / el, cores, el, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # K steps, each a bounded sample (pairs for ~1.5 s of host time, at most the arm's batch)
    v1, cores, el1, n1 = cpu_port_pairs_per_s(args.workload, args.rho, min_seconds=1.5, max_pairs=args.batch,
                                              warmup=max(1, min(args.warmup, 2)))
    steps = max(1, min(args.steps, 40))
    tot_pairs, tot_s = n1, el1
    for _ in range(steps - 1):
        v, _, el, n = cpu_port_pairs_per_s(args.workload, args.rho, min_seconds=1.5, max_pairs=args.batch, warmup=0)
        tot_pairs += n; tot_s += el
    v = tot_pairs / tot_s
    sample = (f"{steps} steps x {n1} pair(s) of the {args.workload} workload (learned detectors calibrated to rho={args.rho}), "
              f"torch-CPU dense/glue + C/OpenMP SpaMat/SpaVar oracle, {tot_s:.1f} s")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": max(1, min(args.warmup, 2)), "ms_per_step": tot_s / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload} {n1} pair(s)/step on the host cores; the reference's own ops for this "
                                   "path are CUDA-only, so its CPU execution is the oracle port (oracle/)",
                       "max_disp": 216},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_from_images(args, info, dev, timed):
    """demo.py's device work from the decoded images on: uint8 RGB pairs -> pad / scale / normalise -> feature extractor
    (both views) -> hot path -> 16-bit disparity image, one CUDA graph; returns the `from_images` object."""
    import torch
    from decnet_b200 import ops
    from decnet_b200.features import FeatExtNetChannelPlus
    from decnet_b200.model import DecompMatching
    from decnet_b200.params import make_featext_state, make_hotpath_state
    from decnet_b200.synthetic import calibrate_mask_density
    B, H, W = args.batch, info["H"], info["W"]
    oh, ow = (540, 960) if args.workload == "sceneflow" else (H, W)          # unpadded image size (demo.py pads top/left)
    fe = FeatExtNetChannelPlus(8, precision=args.precision)
    fe.load_state_dict(make_featext_state(17))
    fe = fe.to(dev)
    model = DecompMatching(max_disp=info["max_disp"], skip_stage_id=info["skip_stage_id"], use_detail=True, thold=0.9,
                           precision=args.precision)
    model.load_state_dict(make_hotpath_state(17))
    model = model.to(dev)
    model.overlap = not args.no_overlap
    g = torch.Generator(device=dev).manual_seed(99)
    sets = []
    for _ in range(2):
        sets.append({"l": torch.randint(0, 256, (B, oh, ow, 3), device=dev, generator=g, dtype=torch.uint8),
                     "r": torch.randint(0, 256, (B, oh, ow, 3), device=dev, generator=g, dtype=torch.uint8)})

    def feats(u8):
        return fe(ops.image_prepare_u8(u8, want01=False)[1])

    dens = calibrate_mask_density(model, feats(sets[0]["l"]), feats(sets[0]["r"]), args.rho)

    from decnet_b200.features import extract_pair

    def step_eager(st):
        if args.no_overlap:
            fl, fr = feats(st["l"]), feats(st["r"])
        else:                                            # right view on a forked second stream
            fl, fr = extract_pair(fe, st["l"], st["r"], prepare=lambda u8: ops.image_prepare_u8(u8, want01=False)[1])
        return ops.disp_to_u16(model(fl, fr)[0], oh, ow)

    for st in sets:
        step_eager(st)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step_eager(st)
            with torch.cuda.graph(graph, stream=side):
                st["out"] = step_eager(st)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        st["graph"] = graph
    for _ in range(3):
        sets[0]["graph"].replay()
    ms = timed(lambda: sets[0]["graph"].replay(), args.steps)
    value = B * args.steps / (ms * 1e-3)
    # end to end: pinned uint8 images -> device (copy stream, double-buffered), graph, uint16 disparity -> pinned host
    host = {k: sets[0][k].cpu().pin_memory() for k in ("l", "r")}
    host_out = torch.empty((B, oh, ow), dtype=torch.uint16).pin_memory()
    copy_stream, main = torch.cuda.Stream(), torch.cuda.current_stream()

    def upload(buf, after):
        with torch.cuda.stream(copy_stream):
            if after is not None:
                copy_stream.wait_event(after)
            for k in ("l", "r"):
                sets[buf][k].copy_(host[k], non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_stream)
        return ev

    def run(n):
        computed = [None, None]
        copied = upload(0, None)
        for i in range(n):
            buf = i & 1
            main.wait_event(copied)
            sets[buf]["graph"].replay()
            host_out.copy_(sets[buf]["out"], non_blocking=True)
            ev = torch.cuda.Event(); ev.record(main); computed[buf] = ev
            if i + 1 < n:
                copied = upload(1 - buf, computed[1 - buf])
        main.synchronize()

    run(3)
    n = max(4, args.steps)
    ms_e2e = timed(lambda: run(n), 1)
    return {"what": "uint8 RGB pairs -> pad/scale/normalise -> feature extractor on both views (FeatExtNetChannelPlus drop-in) "
                    "-> hot path -> uint16 disparity image (x256, cropped), one CUDA graph",
            "value": value, "unit": UNIT, "ms_per_step": ms / args.steps,
            "e2e": {"value": B * n / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / n,
                    "h2d_bytes_per_step": 2 * B * 3 * oh * ow, "d2h_bytes_per_step": B * oh * ow * 2},
            "left_mask_density": [round(d, 4) for d in dens]}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from decnet_b200 import _lib, ops
    from decnet_b200.synthetic import build_workload

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True          # only the feature extractor's library layers (from_images leg) use cuDNN
    lib = _lib.lib()

    bands_mode = args.mode == "bands"
    model, left, right, info = build_workload(args.workload, args.batch, seed=17 + (0 if bands_mode else rank),
                                              device=dev, rho=args.rho, precision=args.precision)
    model.overlap = not args.no_overlap
    B = args.batch
    if bands_mode:
        from decnet_b200 import bands as _bands
        transport = _bands.DistTransport() if world > 1 else _bands.LocalTransport(1)

        def step():
            return _bands.forward_bands(model, left, right, transport)[rank if world > 1 else 0]
    else:
        def step():
            return model(left, right)[0]

    # pinned host copies for the end-to-end leg
    host_l = {k: v.cpu().pin_memory() for k, v in left.items()}
    host_r = {k: v.cpu().pin_memory() for k, v in right.items()}
    dev_l = {k: torch.empty_like(v) for k, v in left.items()}
    dev_r = {k: torch.empty_like(v) for k, v in right.items()}
    host_out = torch.empty((B, info["H"], info["W"]), dtype=torch.float32).pin_memory()
    h2d = sum(v.numel() * 4 for v in host_l.values()) * 2
    d2h = host_out.numel() * 4

    def step_e2e():
        # with the graph, the static input buffers ARE the graph's inputs: copy into them, replay
        tl, tr_ = (left, right) if use_graph else (dev_l, dev_r)
        for k in host_l:
            tl[k].copy_(host_l[k], non_blocking=True)
            tr_[k].copy_(host_r[k], non_blocking=True)
        if use_graph:
            out = step()
        else:
            out = (_bands.forward_bands(model, dev_l, dev_r, transport)[rank if world > 1 else 0] if bands_mode
                   else model(dev_l, dev_r)[0])
        host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    lib.decnet_reset_launch_count()
    step(); torch.cuda.synchronize()
    launches_per_step = int(lib.decnet_launch_count())

    # The step is ~120 launches (ours + cuDNN): captured once into a CUDA graph (static input / output
    # buffers) and replayed, so the device is not waiting for the Python launch path.
    use_graph = not args.no_graph and not bands_mode
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            model(left, right)
            with torch.cuda.graph(graph, stream=side):
                graph_out = model(left, right)[0]
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        eager_step = step

        def step():
            graph.replay()
            return graph_out
        # replay must reproduce the eager result
        ref_out = eager_step()
        assert torch.allclose(step(), ref_out, atol=1e-4, rtol=1e-4), "graph replay differs from eager execution"
        for _ in range(3):
            step()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    prof_range = os.environ.get("DECNET_PROFILE_RANGE") == "1"   # ncu --profile-from-start off
    if prof_range:
        torch.cuda.profiler.start()
    ms = timed(step, args.steps)
    if prof_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if rank == 0 else None
    n_units = B if bands_mode else world * B          # bands: all ranks work on the SAME B pairs
    value = n_units * args.steps / (ms * 1e-3)

    if use_graph:
        # End-to-end leg, software-pipelined: while the graph of step i runs on buffer set i%2, the copy
        # stream uploads the pinned host pyramids of step i+1 into the other set; every step still does
        # its full host->device upload and its device->host read of the disparity.
        left2 = {k: torch.empty_like(v) for k, v in left.items()}
        right2 = {k: torch.empty_like(v) for k, v in right.items()}
        for k in left:
            left2[k].copy_(left[k]); right2[k].copy_(right[k])
        graph2 = torch.cuda.CUDAGraph()
        side2 = torch.cuda.Stream()
        side2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side2):
            model(left2, right2)
            with torch.cuda.graph(graph2, stream=side2):
                graph_out2 = model(left2, right2)[0]
        torch.cuda.current_stream().wait_stream(side2)
        torch.cuda.synchronize()
        sets = [(left, right, graph, graph_out), (left2, right2, graph2, graph_out2)]
        copy_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()

        def upload(buf, after_event):
            with torch.cuda.stream(copy_stream):
                if after_event is not None:
                    copy_stream.wait_event(after_event)       # the set's previous graph replay has finished
                tl, tr_ = sets[buf][0], sets[buf][1]
                for k in host_l:
                    tl[k].copy_(host_l[k], non_blocking=True)
                    tr_[k].copy_(host_r[k], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return ev

        def run_e2e(nsteps):
            computed = [None, None]
            copied = upload(0, None)
            for i in range(nsteps):
                buf = i & 1
                main.wait_event(copied)
                sets[buf][2].replay()
                host_out.copy_(sets[buf][3], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(main); computed[buf] = ev
                if i + 1 < nsteps:
                    copied = upload(1 - buf, computed[1 - buf])
            main.synchronize()

        run_e2e(3)
        e2e_steps = max(4, args.steps)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(e2e_steps)
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)                                                  # device clock
        if world > 1:
            t = torch.tensor([ms_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
        e2e_note = ("pinned host feature pyramids -> device (copy stream, double-buffered against the previous "
                    "step's compute), hot path (CUDA graph), disparity -> pinned host, every step")
    else:
        for _ in range(2):
            step_e2e()
        e2e_steps = max(3, args.steps // 2)
        ms_e2e = timed(step_e2e, e2e_steps)
        e2e_note = "pinned host feature pyramids -> device, hot path, disparity -> pinned host, every step"
    e2e_value = n_units * e2e_steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant sparse kernel (fused SpaMat+SpaVar at the finest level), measured
    # live with CUDA events on the launching stream; inputs (370 MB at B=8) exceed the 126 MB L2.
    roof = None
    roof_tensor = None
    roof_conv2d = None
    if rank == 0:
        pk = peaks()
        s = 3 if info["skip_stage_id"] > 3 else 2
        Lf, Rf = left[f"stage{s}"], right[f"stage{s}"]
        Bc, Cc, Hc, Wc = Lf.shape
        Dc = info["max_disp"] // 3 ** (3 - s)
        g = torch.Generator(device=dev).manual_seed(5)
        pl = torch.rand(Bc, Hc, Wc, device=dev, generator=g)
        pr = torch.rand(Bc, Hc, Wc, device=dev, generator=g)
        ml, mr = ops.mask_threshold(pl, pr, 1.0 - args.rho)
        for _ in range(3):
            ops.spamat_spavar_forward(Lf, Rf, ml, mr, Dc)
        iters = 20
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.spamat_spavar_forward(Lf, Rf, ml, mr, Dc)
        e1.record(); torch.cuda.synchronize()
        t_k = e0.elapsed_time(e1) * 1e-3 / iters
        alg = 4.0 * Bc * Hc * Wc * (2 * Cc + 2 + 4)
        ach = alg / t_k / 1e9
        roof = {"bound": "hbm", "kernel": "sparse_row_gather_kernel<FUSED> (SpaMat+SpaVar, finest level)",
                "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": None, "peak_source": pk["source"] + " (burst copy)", "algorithmic_bytes": alg,
                "us_per_launch": t_k * 1e6, "mask_density": args.rho}
        tr = ROOT / "profiles" / "traffic.json"
        if tr.exists():
            try:
                roof["traffic"] = json.loads(tr.read_text()).get("sparse_row_gather_kernel_bytes_per_launch")
            except Exception:
                pass
        # the same launch at other mask densities, and the staged (TMA) row kernel beside the default sector-gather
        # kernel: us per launch, same timing method (back-to-back launches, inputs larger than L2)
        def time_sparse(mlx, mrx, n=10):
            for _ in range(2):
                ops.spamat_spavar_forward(Lf, Rf, mlx, mrx, Dc)
            e0.record()
            for _ in range(n):
                ops.spamat_spavar_forward(Lf, Rf, mlx, mrx, Dc)
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / n
        sweep = {}
        for rho_s in (0.01, 0.03, args.rho):
            mls, mrs = ops.mask_threshold(pl, pr, 1.0 - rho_s)
            rec = {"gather_us": round(time_sparse(mls, mrs), 2)}
            lib.decnet_set_sparse_variant(1)
            try:
                rec["staged_us"] = round(time_sparse(mls, mrs), 2)
            except Exception:
                rec["staged_us"] = None
            finally:
                lib.decnet_set_sparse_variant(0)
            rec["gather_frac"] = round(alg / (rec["gather_us"] * 1e-6) / 1e9 / pk["hbm_gbs"], 4)
            sweep[str(rho_s)] = rec
        roof["density_sweep"] = sweep
        # the thin 3x3 Conv2d layers (conv2d_tcgen05_kernel, 31 % of the step): the 8->8 layer at the finest level,
        # algorithmic bytes = input + output once
        try:
            Bq, Cq, Hq, Wq = left["stage3"].shape
            if ops.conv2d_tf32_supported(Cq, 8, Hq, Wq, 1):
                gq = torch.Generator(device=dev).manual_seed(7)
                wq = torch.randn(8, Cq, 3, 3, device=dev, generator=gq) * 0.1
                split_q = args.precision == "fp32"
                wpk, bpk = ops.pack_conv2d_tf32_nchw_weights(wq, torch.zeros(8, device=dev), split=split_q)
                xq = left["stage3"]
                for _ in range(3):
                    ops.conv2d_tf32_nchw_cat([xq], wpk, bpk, 8, 1, True, split=split_q)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    ops.conv2d_tf32_nchw_cat([xq], wpk, bpk, 8, 1, True, split=split_q)
                e1.record(); torch.cuda.synchronize()
                t_c = e0.elapsed_time(e1) * 1e-3 / 20
                alg_c = 4.0 * Bq * Hq * Wq * (Cq + 8)
                roof_conv2d = {"bound": "hbm", "kernel": "conv2d_tcgen05_kernel (3x3, 8->8 channels, finest level, " + ("3xTF32" if split_q else "TF32") + ")",
                               "achieved": alg_c / t_c / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                               "frac": alg_c / t_c / 1e9 / pk["hbm_gbs"], "traffic": None, "algorithmic_bytes": alg_c,
                               "us_per_launch": t_c * 1e6}
                if tr.exists():
                    try:
                        roof_conv2d["traffic"] = json.loads(tr.read_text()).get("conv2d_tcgen05_kernel_bytes_per_launch")
                    except Exception:
                        pass
        except Exception as e:
            roof_conv2d = {"bound": "hbm", "note": f"{type(e).__name__}: {e}"}
        # tensor roofline of the coarse 3-D aggregation (a3), timed inside the step
        try:
            from decnet_b200 import conv3d as c3
            roof_tensor = c3.measure_roofline(model, left["stage0"], right["stage0"], info["max_disp"] // 27, pk)
        except Exception as e:
            roof_tensor = {"bound": "tensor", "note": f"{type(e).__name__}: {e}"}

    # ---- extra leg: the same step fed from IMAGES -- feature extractor on both
    # views + hot path in one CUDA graph; its end-to-end form uploads 2 x B images (a quarter of the bytes of
    # the feature pyramids) from pinned host memory and reads the disparity back, every step.
    from_images = None
    if rank == 0 and world == 1 and use_graph and not args.no_from_images:
        try:
            from_images = run_from_images(args, info, dev, timed)
        except Exception as e:      # the headline line must not depend on the extra leg
            from_images = {"error": f"{type(e).__name__}: {e}"}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        v, cores, sec, npairs = cpu_port_pairs_per_s(args.workload, args.rho, min_seconds=args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{npairs} pairs of the same workload (one pair per call, learned detectors calibrated to "
                         f"rho={args.rho}), torch-CPU dense/glue + C/OpenMP SpaMat/SpaVar oracle, {sec:.1f} s"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if bands_mode else "weak", "vs_baseline": None,
                "dtype": "f32 (sparse/glue), bf16 in / f32 acc (3-D aggregation)",
                "data": "synthetic",
                "config": {"workload": f"{args.workload} {info['H']}x{info['W']} padded, batch {B}/GPU, max_disp {info['max_disp']}, "
                                       "full decomposition pyramid" + (" (BASELINE.json configs[1])" if args.workload == "sceneflow" and B == 8 else ""),
                           "levels": " | ".join(f"1/{27 // 3 ** i} C{c} {info['H'] * 3 ** i // 27}x{info['W'] * 3 ** i // 27} "
                                                f"D{info['max_disp'] * 3 ** i // 27}" for i, c in enumerate((216, 72, 24, 8))),
                           "left_mask_density": info["left_mask_density"], "precision": args.precision,
                           "launch": "CUDA graph replay of the whole step" if use_graph else "eager launches",
                           "streams": "masks + sparse ops on a forked second stream (two graph branches)" if not args.no_overlap else "one stream",
                           "l2": "inputs (400 MB of feature pyramids per step) exceed the 126 MB L2; no flush",
                           "parallelism": (f"row bands of one batch over {world} rank(s): per-layer halo send/recv in the "
                                           "3-D aggregation, all-gather of the per-level disparity") if bands_mode
                           else f"by stereo pair, {world} rank(s), no collective"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / e2e_steps,
                        "note": e2e_note},
                "gpu_launches": launches_per_step * args.steps,
                "gpu_launches_per_step": launches_per_step,
                "clocks": clocks, "roofline": roof, "roofline_tensor": roof_tensor, "roofline_conv2d": roof_conv2d,
                "cpu_baseline": cpu,
                "from_images": from_images}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
