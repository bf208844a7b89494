"""N-rank check of the peer-memory band transport: banded Middlebury-like pair vs the single-device result, eager and graph."""
import os
import sys
from pathlib import Path
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import bands
from decnet_b200.synthetic import build_workload

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
wl = sys.argv[1] if len(sys.argv) > 1 else "middlebury"
model, left, right, info = build_workload(wl, 1, seed=17, device=dev, rho=0.1)
want = model(left, right)[0]
tr = bands.PeerTransport(left["stage0"].shape[2])
got = bands.forward_bands(model, left, right, tr)[rank].clone()
torch.cuda.synchronize()
scale = float(want.abs().max())
d = (got - want).abs()
print(rank, "eager: max", float(d.max()), "mean", float(d.mean()), "scale", scale, flush=True)
dist.barrier()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    bands.forward_bands(model, left, right, tr)
    with torch.cuda.graph(g, stream=s):
        out = bands.forward_bands(model, left, right, tr)[rank]
torch.cuda.current_stream().wait_stream(s)
dist.barrier()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
d = (out - want).abs()
print(rank, "graph: max", float(d.max()), "mean", float(d.mean()), flush=True)
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g.replay()
e1.record(); torch.cuda.synchronize()
print(rank, "graph replay ms/pair", e0.elapsed_time(e1) / 10, flush=True)
dist.barrier()
dist.destroy_process_group()
