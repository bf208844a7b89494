"""CPU, world_size 2 over gloo: the N>1 host logic (pair sharding, metric reduction, row bands)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from decnet_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_pairs(n_pairs, world, rank)
    # every pair has a deterministic "EPE" so the reduced mean can be checked against the serial answer
    g = torch.Generator().manual_seed(0)
    epe = torch.rand(n_pairs, generator=g, dtype=torch.float64)
    npx = torch.arange(1, n_pairs + 1, dtype=torch.float64) * 100
    s = float((epe[mine] * npx[mine]).sum()); n = float(npx[mine].sum())
    mean, total = shard.reduce_metrics(s, n)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, mine, mean, total, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_pair_sharding_and_metric_reduction_world2():
    world, n_pairs = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    g = torch.Generator().manual_seed(0)
    epe = torch.rand(n_pairs, generator=g, dtype=torch.float64)
    npx = torch.arange(1, n_pairs + 1, dtype=torch.float64) * 100
    want = float((epe * npx).sum() / npx.sum())
    seen = []
    for rank, mine, mean, total, gathered in res:
        assert mine == list(range(rank, n_pairs, world))
        assert mean == pytest.approx(want, rel=1e-12) and total == float(npx.sum())
        seen = sorted(sum(gathered, []))
    assert seen == list(range(n_pairs))                     # disjoint cover, no pair twice


def test_shard_pairs_edge_cases():
    assert shard.shard_pairs(0, 4, 1) == []
    assert shard.shard_pairs(3, 8, 5) == []
    assert shard.shard_pairs(64, 8, 7) == list(range(7, 64, 8))
    with pytest.raises(ValueError):
        shard.shard_pairs(4, 2, 2)


@pytest.mark.parametrize("h_coarse,world", [(75, 8), (20, 8), (14, 4), (5, 8)])
def test_row_bands_cover_and_extend(h_coarse, world):
    bands = shard.coarse_bands(h_coarse, world)
    assert bands[0][0] == 0 and bands[-1][1] == h_coarse
    assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
    assert max(b - a for a, b in bands) - min(b - a for a, b in bands) <= 1
    for stage in (1, 2, 3):
        ext = shard.stage_extension(stage)
        prev = None
        for r in range(world):
            b = shard.level_band(h_coarse, world, r, stage, ext)
            assert 0 <= b.e0 <= b.r0 <= b.r1 <= b.e1 <= h_coarse * 3 ** stage
            assert b.rows == (bands[r][1] - bands[r][0]) * 3 ** stage
            if prev is not None:
                assert prev.r1 == b.r0
            prev = b
        assert shard.pred_halo_coarse(stage) * 3 >= ext
