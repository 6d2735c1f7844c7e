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
    dist.all_gather_into_tensor(fb, band, group=group)  # in place; supported by both NCCL and gloo -- no silent fallback


def resolve_across_ranks(ctx, world: int, dist=None, group=None, max_rounds: int = 4) -> int:
    """Wait for the frame in flight on every rank and agree on its fate.  A rank whose bin lists overflowed regrows them
    and reports FDC_ERR_RETRY (6) instead of re-running privately (peers gathered / read rows of the aborted frame): the
    statuses are max-reduced and, if any rank said retry, EVERY rank calls fdc_retry_frame -- same barrier sequence, same
    pixels.  `ctx` needs syncStatus() and retryFrame() (CudaContext); returns the number of re-runs."""
    rounds = 0
    while True:
        rc = int(ctx.syncStatus())
        if world > 1:
            import torch

            if dist is None:
                import torch.distributed as dist  # noqa: PLW0642
            backend = dist.get_backend(group)
            dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
            t = torch.tensor([rc], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            rc = int(t.item())
        if rc != 6:
            return rounds
        if rounds >= max_rounds:
            raise RuntimeError("bin lists still overflow after %d cross-rank retries" % rounds)
        ctx.retryFrame()
        rounds += 1
