"""Multi-GPU partitioning of the hot path.  One process per GPU.

* Throughput mode: shard BY STEREO PAIR (pair i -> rank i mod G).  Every op is per-pair and BN is
  eval-mode, so there is no data-path collective -- mirrors the reference's DataParallel-by-batch
  (eval.py:145-146).  Only metrics are reduced (2 scalars).
* Single-huge-pair mode: contiguous ROW BANDS of the 1/27 grid; every finer level's band is 3x the
  rows.  Sparse ops / threshold / blend are row-local (no halo); the stencil ops need halos.
"""
from __future__ import annotations

from dataclasses import dataclass


def shard_pairs(n_pairs: int, world: int, rank: int) -> list[int]:
    """Indices of the stereo pairs rank `rank` processes (round-robin: pair i -> rank i mod world)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_pairs, world))


def reduce_metrics(epe_sum: float, n_pixels: float, group=None):
    """All-reduce (SUM) of the two scalars behind the mean EPE; returns (mean_epe, total_pixels).
    Works on any initialised backend (nccl on GPU ranks, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return epe_sum / max(n_pixels, 1.0), n_pixels
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([epe_sum, n_pixels], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t[0] / t[1].clamp(min=1.0)), float(t[1])


# receptive-field halos, in rows of the level they apply to
REFINE_HALO = {1: 7, 2: 16, 3: 22}      # sum of the dilations of the 7 convs (submodule.py:687-716)
ATTN_HALO = 3                            # 3 convs 3x3
DETAIL_HALO = 3                          # conv_sub (2) + conv.0 (1); deconv k3 s3 adds none
DYNUP_HALO_COARSE = 1 + 3                # 3x3 gather on the coarse map + 3 convs 3x3 on the coarse grid
CONV3D_HALO_PER_LAYER = 1                # 3x3x3
WARP_HALO = 1                            # vertical bilinear taps (y' = h*H/(H-1) - 0.5)


@dataclass(frozen=True)
class Band:
    """Rows [r0, r1) of a level owned by a rank, and the extended range [e0, e1) it must compute on
    (clipped to the image) so that the owned rows are exact."""
    r0: int
    r1: int
    e0: int
    e1: int

    @property
    def rows(self):
        return self.r1 - self.r0


def coarse_bands(h_coarse: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced split of the 1/27 rows (first `h % world` ranks get one row more)."""
    base, extra = divmod(h_coarse, world)
    out, r = [], 0
    for k in range(world):
        n = base + (1 if k < extra else 0)
        out.append((r, r + n))
        r += n
    return out


def level_band(h_coarse: int, world: int, rank: int, stage: int, extend: int) -> Band:
    """Band of `stage` (0 = 1/27 ... 3 = full res): the coarse band scaled by 3**stage, extended by
    `extend` rows on both sides and clipped to the image."""
    r0c, r1c = coarse_bands(h_coarse, world)[rank]
    f = 3 ** stage
    H = h_coarse * f
    r0, r1 = r0c * f, r1c * f
    return Band(r0, r1, max(0, r0 - extend), min(H, r1 + extend))


def stage_extension(stage: int) -> int:
    """Rows of `stage` by which a band must be extended so that the refinement output on the owned
    rows is exact when every conv stack runs on the extended band with zero padding at its edges."""
    return REFINE_HALO[stage] + ATTN_HALO + DETAIL_HALO + WARP_HALO


def pred_halo_coarse(stage: int) -> int:
    """Rows of the PREVIOUS level's disparity a rank needs beyond its own (previous-level) band to run
    `stage` on its extended band: ceil(extension / 3) + the dynamic-upsampling halo."""
    return -(-stage_extension(stage) // 3) + DYNUP_HALO_COARSE
