"""CPU: the weight layouts the C ABI documents (include/decnet_b200.h) -- unpack what ops.pack_* produced with the
index formulas of the header and compare with the original weights (no GPU, no kernel call)."""
import torch

from decnet_b200 import ops


def _tf32(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def test_nchw_tf32_pack_layout_matches_header_formula():
    torch.manual_seed(0)
    chans = (5, 1, 12)
    cout, cin = 6, sum(chans)
    w = torch.randn(cout, cin, 3, 3)
    b = torch.randn(cout)
    wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b, chans)
    cpad = ops.padded_cat_channels(chans)
    assert cpad == 8 + 8 + 16
    nck, cp = cpad // 8, 8
    natoms = (3 * cp + 31) // 32
    rows = wp.reshape(-1, 32)
    assert rows.shape[0] == 3 * nck * natoms * 8
    # padded channel index of every original channel
    pos, o = [], 0
    for c in chans:
        pos += list(range(o, o + c)); o += (c + 7) // 8 * 8
    wr = _tf32(w)
    for kh in range(3):
        for kw in range(3):
            for co in range(cout):
                for ci, pc in enumerate(pos):
                    chunk, k = divmod(pc, 8)
                    col = kw * cp + co
                    atom, n = divmod(col, 32)
                    row = ((kh * nck + chunk) * natoms + atom) * 8 + k
                    assert rows[row, n] == wr[co, ci, kh, kw]
    assert rows.abs().sum() == wr.abs().sum()            # everything else is zero padding
    assert torch.equal(bp[:cout], b) and bp[cout:].abs().sum() == 0


def test_rows_pack_layout():
    torch.manual_seed(1)
    w, b = torch.randn(3, 11, 3, 3), torch.randn(3)
    wc, b8 = ops.pack_conv2d_tf32_rows_weights(w, b, (3, 8))
    assert tuple(wc.shape) == (3, 3, 2, 8, 8) and tuple(b8.shape) == (8,)
    wr = _tf32(w)
    assert torch.equal(wc[:, :, 0, :3, :3], wr[:, :3].permute(2, 3, 0, 1))       # source 0: channels 0..2 of chunk 0
    assert torch.equal(wc[:, :, 1, :3, :], wr[:, 3:].permute(2, 3, 0, 1))        # source 1: chunk 1
    assert wc.abs().sum() == wr.abs().sum()
