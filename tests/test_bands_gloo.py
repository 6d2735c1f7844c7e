"""N>1 host logic on CPU: band layout + in-place band all-gather over gloo (world_size 2 and 3).

Each rank produces its band of a frame (with the CPU oracle, rows restricted to the band) into the padded
framebuffer layout bench.py uses, gathers, and the result must equal the single-process frame bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from figdraw_b200 import bands, scenes_synth as ss


def test_band_layout_matches_backend_rule():
    rows, b = bands.band_layout(2160, 8)
    assert rows == 17 * 16 and b[0] == (0, 272) and b[-1] == (1904, 2160)
    rows, b = bands.band_layout(100, 3)  # 7 tile rows -> 3,3,1
    assert b == [(0, 48), (48, 96), (96, 100)] and rows == 48
    rows, b = bands.band_layout(16, 4)   # more ranks than tile rows: empty bands
    assert b == [(0, 16), (16, 16), (16, 16), (16, 16)]
    assert bands.padded_rows(2160, 8) >= 2160


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle

    tr = ss.config_trace(5, 320, 200, n_rects=400, n_glyphs=80)
    rows, layout = bands.band_layout(tr.height, world)
    fb = torch.zeros((rows * world, tr.width, 4), dtype=torch.uint8)
    y0, y1 = layout[rank]
    if y1 > y0:
        img = oracle.render_trace(tr, n_threads=1, rows=(y0, y1))
        fb[y0:y1] = torch.from_numpy(img[y0:y1])
    bands.allgather_bands(fb, rank, world)
    if rank == 0:
        np.save(os.path.join(out_dir, f"gathered_{world}.npy"), fb[: tr.height].numpy())

    class FakeCtx:
        """Stands in for CudaContext: the last rank's bin lists 'overflow' on the first attempt (FDC_ERR_RETRY = 6)."""

        def __init__(self):
            self.syncs, self.retries = 0, 0

        def syncStatus(self):
            self.syncs += 1
            return 6 if (rank == world - 1 and self.retries == 0) else 0

        def retryFrame(self):
            self.retries += 1

    fake = FakeCtx()
    rounds = bands.resolve_across_ranks(fake, world, dist)
    assert rounds == 1 and fake.retries == 1 and fake.syncs == 2  # EVERY rank re-runs the frame, once

    class FakeBandCtx:
        """Each rank knows the tile-entry counts of its own rows only (equal bands to start with)."""

        def __init__(self):
            tiles_y = 13
            per = (tiles_y + world - 1) // world
            whole = np.array([1, 1, 1, 1, 40, 40, 40, 1, 1, 1, 1, 1, 1], dtype=np.uint32)
            self.costs = np.zeros(tiles_y, dtype=np.uint32)
            self.costs[rank * per:min((rank + 1) * per, tiles_y)] = whole[rank * per:min((rank + 1) * per, tiles_y)]
            self.bounds = None

        def tileRowCosts(self):
            return self.costs

        def setBandTileRows(self, b):
            self.bounds = list(b)

    fb_ctx = FakeBandCtx()
    got = bands.rebalance_across_ranks(fb_ctx, world, tiles_x=1, dist=dist, tile_cost=0.0)
    gathered = [None] * world
    dist.all_gather_object(gathered, got)
    assert all(g == gathered[0] for g in gathered) and fb_ctx.bounds == got  # every rank chose the same boundaries
    assert got == bands.balance_rows([1, 1, 1, 1, 40, 40, 40, 1, 1, 1, 1, 1, 1], world)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_allgather_reassembles_frame(tmp_path, world):
    from oracle import oracle

    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / f"gathered_{world}.npy")
    want = oracle.render_trace(ss.config_trace(5, 320, 200, n_rects=400, n_glyphs=80), n_threads=2)
    assert np.array_equal(got, want)


def test_balance_rows_minimises_the_heaviest_band():
    costs = [1, 1, 1, 1, 40, 40, 40, 1, 1, 1, 1, 1, 1]
    b = bands.balance_rows(costs, 3)
    assert b[0] == 0 and b[-1] == len(costs) and len(b) == 4
    assert max(sum(costs[b[i]:b[i + 1]]) for i in range(3)) == 46  # {1,1,1,1,40} {40} {40,1,1,1,1,1,1}: nothing beats 46
    assert bands.balance_rows([5] * 8, 8) == list(range(9))        # one row each
    assert bands.balance_rows([1, 2, 3], 5) == [0, 1, 2, 3, 3, 3]   # more bands than rows: the tail is empty
    assert bands.balance_rows([], 2) == [0, 0, 0]
    rng = np.random.default_rng(3)
    for _ in range(300):
        n, w = int(rng.integers(1, 80)), int(rng.integers(1, 9))
        c = rng.choice([0, 1, 7, 60], size=n).tolist()
        b = bands.balance_rows(c, w)
        assert len(b) == w + 1 and b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))
        if n >= w:
            assert all(b[i] < b[i + 1] for i in range(w))
        # optimal against brute force on small cases
        if n <= 9 and w <= 3 and n >= w:
            import itertools
            best = min(max(sum(c[x:y]) for x, y in zip((0,) + cut, cut + (n,))) for cut in itertools.combinations(range(1, n), w - 1))
            assert max(sum(c[b[i]:b[i + 1]]) for i in range(w)) == best
