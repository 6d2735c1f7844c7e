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


def balance_rows(costs, world: int) -> List[int]:
    """Contiguous partition of the tile rows into `world` bands that minimises the largest band cost.  Returns world + 1
    boundaries (tile rows).  Every band gets at least one row while there are rows to give."""
    c = [float(v) for v in costs]
    n = len(c)
    if world <= 1 or n == 0:
        return [0] + [n] * max(world, 1)

    def bands_needed(limit: float) -> int:
        used, run = 1, 0.0
        for v in c:
            if run + v > limit and run > 0.0:
                used, run = used + 1, 0.0
            run += v
        return used

    lo, hi = max(c), sum(c)
    for _ in range(48):  # smallest limit that `world` bands can meet
        mid = 0.5 * (lo + hi)
        if bands_needed(mid) <= world:
            hi = mid
        else:
            lo = mid
    bounds, run = [0], 0.0
    for i, v in enumerate(c):
        rows_left, bands_left = n - i, world - (len(bounds) - 1)
        # start a new band when this row would break the limit -- or when the rows left are only just enough to give
        # every remaining band one
        if len(bounds) <= world - 1 and i > bounds[-1] and (run + v > hi * (1.0 + 1e-9) or rows_left < bands_left):
            bounds.append(i)
            run = 0.0
        run += v
    while len(bounds) < world:  # fewer rows than bands, or the limit left bands over: give the tail single rows / nothing
        bounds.append(min(n, bounds[-1] + 1) if bounds[-1] < n and n - bounds[-1] > world - len(bounds) else n)
    bounds.append(n)
    return bounds


def rebalance_across_ranks(ctx, world: int, tiles_x: int, dist=None, group=None, tile_cost: float = 3.0) -> List[int]:
    """Choose bands from the tile-entry profile of the frame every rank has just rendered: sum the ranks' per-row entry
    counts (each rank knows its own rows), add a constant per tile (clearing, copying a tile out costs about as much as
    shading `tile_cost` entries), split into `world` contiguous bands of equal cost and hand the boundaries to the
    context (fdc_set_band_tile_rows).  Every rank computes the same boundaries.  Returns them."""
    import numpy as np

    costs = np.asarray(ctx.tileRowCosts(), dtype=np.int64)
    if world > 1:
        import torch

        if dist is None:
            import torch.distributed as dist  # noqa: PLW0642
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        t = torch.from_numpy(costs).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        costs = t.cpu().numpy()
    bounds = balance_rows(costs.astype(np.float64) + tile_cost * float(tiles_x), world)
    ctx.setBandTileRows(bounds)
    return bounds
