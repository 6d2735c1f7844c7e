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
