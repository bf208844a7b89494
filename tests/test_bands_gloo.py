"""CPU, two processes over gloo: bands.DistTransport -- the neighbour send/recv of the halo rows and the all-gather of the
per-level disparity bands (the host-side logic of the torch.distributed band transport; the NVLink peer-memory transport is
covered on GPUs by tests/test_bands_gpu.py and bench.py's `bands` object)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from decnet_b200 import bands, shard
        tr = bands.DistTransport()
        assert tr.world == world and tr.rank == rank and tr.ranks_here == [rank]
        h0 = 7
        cb = shard.coarse_bands(h0, world)
        rows = cb[rank][1] - cb[rank][0]
        B, D, W, C = 1, 2, 5, 4
        # band tensor with halo slots: owned rows carry (global row index + 1), halo slots start as garbage
        x = torch.full((B, D, rows + 2, W, C), -99.0)
        for i in range(rows):
            x[:, :, 1 + i] = float(cb[rank][0] + i + 1)
        tr.exchange_halo({rank: x})
        top = 0.0 if rank == 0 else float(cb[rank][0])              # the neighbour's last owned row, zero at the image edge
        bot = 0.0 if rank == world - 1 else float(cb[rank][1] + 1)
        ok = bool((x[:, :, 0] == top).all()) and bool((x[:, :, -1] == bot).all())
        # gather of a x3 level
        tr.h_coarse = h0
        mine = torch.arange(cb[rank][0] * 3, cb[rank][1] * 3, dtype=torch.float32).view(1, -1, 1).expand(1, -1, 6).contiguous()
        full = tr.all_gather_rows({rank: mine})[rank]
        ok = ok and full.shape == (1, 3 * h0, 6) and bool((full[0, :, 0] == torch.arange(3 * h0, dtype=torch.float32)).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_dist_transport_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
