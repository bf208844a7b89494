"""2-rank probe: torch symmetric memory (peer-mapped buffers + raw pointers) under eager and CUDA-graph execution."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok; ptrs", [hex(p) for p in hdl.buffer_ptrs], flush=True)
    peer = (rank + 1) % world
    pb = hdl.get_buffer(peer, (1 << 20,), torch.float32)
    t.fill_(-1.0)
    hdl.barrier(0)
    pb[:1024].copy_(torch.full((1024,), float(rank + 10), device=dev))      # write into the peer's buffer
    hdl.barrier(0)
    torch.cuda.synchronize()
    print(rank, "eager: my buffer now holds", t[:4].tolist(), "expected", float((rank - 1) % world + 10), flush=True)
    # the same under a CUDA graph
    src = torch.full((1024,), float(rank + 100), device=dev)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            pb[:1024].copy_(src)
            hdl.barrier(1)
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    print(rank, "graph: my buffer now holds", t[:4].tolist(), "expected", float((rank - 1) % world + 100), flush=True)
except Exception as e:
    import traceback
    traceback.print_exc()
    print(rank, "FAILED", type(e).__name__, e, flush=True)
dist.barrier()
dist.destroy_process_group()
