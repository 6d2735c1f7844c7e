"""Tile-band partition of the framebuffer across ranks and the end-of-frame band all-gather.

The partition mirrors `compute_frame_view` in csrc/fdc_context.cu: bands are whole rows of 16-px tiles,
`ceil(tile_rows / world)` tile rows per rank.  The framebuffer a rank renders into is padded to `world` equal bands so
the gather is one in-place `all_gather_into_tensor` (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

TILE_H = 16


def band_layout(height: int, world: int, tile_h: int = TILE_H) -> Tuple[int, List[Tuple[int, int]]]:
    """Returns (rows per padded band, [(y0, y1) pixel rows owned by each rank])."""
    tiles_y = (height + tile_h - 1) // tile_h
    per = (tiles_y + world - 1) // world
    bands = []
    for r in range(world):
        t0 = min(r * per, tiles_y)
        t1 = min(t0 + per, tiles_y)
        bands.append((t0 * tile_h, min(t1 * tile_h, height)))
    return per * tile_h, bands


def padded_rows(height: int, world: int) -> int:
    return band_layout(height, world)[0] * world


def allgather_bands(fb, rank: int, world: int, group=None) -> None:
    """`fb`: uint8 tensor [padded_rows, W, 4]; every rank has written its own band.  In place."""
    if world == 1:
        return
    import torch.distributed as dist

    rows = fb.shape[0] // world
    band = fb[rank * rows:(rank + 1) * rows]
    try:
        dist.all_gather_into_tensor(fb, band, group=group)
    except (RuntimeError, NotImplementedError):
        parts = [fb[r * rows:(r + 1) * rows] for r in range(world)]
        tmp = [p.clone() for p in parts]
        dist.all_gather(tmp, band.clone(), group=group)
        for p, t in zip(parts, tmp):
            p.copy_(t)
