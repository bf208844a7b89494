"""GPU: row-band execution (all ranks simulated in one process, LocalTransport) against the
single-device pipeline on the same pair.  The multi-process NCCL transport is exercised by
`bench.py --mode bands` under torchrun."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(H, W, max_disp, skip, use_detail, seed=31, B=1):
    from decnet_b200.model import DecompMatching
    from decnet_b200.params import make_features, make_hotpath_state
    m = DecompMatching(max_disp=max_disp, skip_stage_id=skip, use_detail=use_detail, thold=0.6)
    m.load_state_dict(make_hotpath_state(seed))
    m = m.cuda()
    left, right = make_features(B, H, W, seed=seed, device="cuda")
    g = torch.Generator().manual_seed(seed + 1)
    lm = [(torch.rand(B, H // f, W // f, generator=g) < 0.2).float().cuda() for f in (9, 3, 1)]
    rm = [(torch.rand(B, H // f, W // f, generator=g) < 0.2).float().cuda() for f in (9, 3, 1)]
    return m, left, right, lm, rm


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("H,W,max_disp,skip,use_detail", [(162, 108, 216, 4, False), (162, 135, 243, 3, True),
                                                           (243, 108, 216, 4, True)])
def test_bands_match_single_device(world, H, W, max_disp, skip, use_detail):
    from decnet_b200 import bands
    m, left, right, lm, rm = _build(H, W, max_disp, skip, use_detail)
    want, taps = m(left, right, lm, rm, is_check=True)
    tr = bands.LocalTransport(world)
    full = bands.forward_bands(m, left, right, tr, lm, rm)
    for r in range(world):
        got = full[r]
        assert got.shape == want.shape
        scale = float(want.abs().max())
        err = float((got - want).abs().max())
        # band-sized tiles change the accumulation grouping; ~20 random-init layers amplify that noise
        assert err <= 1e-3 + 5e-3 * scale, (r, err, scale)
        assert float((got - want).abs().mean()) <= 1e-3 + 1e-4 * scale


def test_dense_stage_bands_bit_exact():
    """The banded 3-D aggregation (halo rows exchanged after every layer) reproduces the full-volume
    result exactly: same kernel, same K order, halo values == the neighbour's owned rows."""
    from decnet_b200 import bands
    m, left, right, lm, rm = _build(243, 108, 216, 4, False)
    want, _ = m.dense_stage(left["stage0"], right["stage0"], 8)
    for world in (2, 3, 4):
        tr = bands.LocalTransport(world)
        pred = bands._dense_stage_bands(m, left["stage0"], right["stage0"], 8, tr)
        got = torch.cat([pred[r] for r in range(world)], dim=1)
        assert torch.equal(got, want), (world, (got - want).abs().max())


def test_conv3d_band_mode_leaves_the_halo_rows_alone():
    """decnet_conv3d_bf16_band: same values as the plain layer on the owned rows, and no store at all into rows 0 / H-1
    (they belong to the neighbouring ranks, which fill them over peer memory)."""
    from decnet_b200 import conv3d as c3
    g = torch.Generator(device="cuda").manual_seed(3)
    B, D, H, W, C = 1, 8, 7, 36, 224
    x = torch.randn(B, D, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(27, C, C, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
    bias = torch.randn(C, device="cuda", generator=g) * 0.1
    want = c3.conv3d_layer(x, w, bias, C, True)
    out = torch.full((B, D, H, W, C), 7.0, device="cuda", dtype=torch.bfloat16)
    c3.conv3d_layer(x, w, bias, C, True, out=out, band=True)
    assert torch.equal(out[:, :, 1:-1], want[:, :, 1:-1])
    assert float((out[:, :, 0].float() - 7.0).abs().max()) == 0 and float((out[:, :, -1].float() - 7.0).abs().max()) == 0


def test_peer_transport_one_rank_equals_single_device():
    """bands.PeerTransport (symmetric-memory buffers, band-mode aggregation layers, device-side signals) with a process
    group of ONE rank: the code path of the multi-GPU band mode on a single GPU; the N > 1 runs are bench.py's `bands` object
    and scripts/probe/bands_probe.py under torchrun."""
    import os
    import socket
    import torch.distributed as dist
    from decnet_b200 import bands
    if dist.is_initialized():
        pytest.skip("a process group already exists in this process")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        m, left, right, lm, rm = _build(162, 135, 243, 3, True)
        want = m(left, right, lm, rm)[0]
        tr = bands.PeerTransport(left["stage0"].shape[2])
        got = bands.forward_bands(m, left, right, tr, lm, rm)[0]
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 1e-3 + 5e-3 * scale
        assert float((got - want).abs().mean()) <= 1e-3 + 1e-4 * scale
    finally:
        dist.destroy_process_group()
