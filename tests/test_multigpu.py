"""Two or more B200s: the band all-gather fused into the shade kernel's copy-out through the NVSwitch multicast
mapping of a torch symmetric-memory framebuffer.  Needs >= 2 GPUs (skipped on the single-GPU test box); run with
    gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu -q
The test spawns one process per GPU itself."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm

    from figdraw_b200 import bands, scenes_synth as ss
    from figdraw_b200.cuda_context import CudaContext, prepare_calls

    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    results = {}
    for name, tr in (("cfg5", ss.config_trace(5, 1920, 1080, n_rects=8000, n_glyphs=1500)), ("cfg2", ss.config_trace(2)),
                     ("cfg4", ss.config_trace(4, 1920, 1080, rows=60, cols=8))):
        rows = bands.padded_rows(tr.height, world)
        nbytes = ((tr.width * rows * 4 + 255) & ~255) + 4096 + 64 * (len(tr.calls) + 64)  # pixels + flags + record exchange area
        t = symm.empty(nbytes, dtype=torch.uint8, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD)
        t.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        ctx = CudaContext(atlasSize=tr.atlas_size, device=rank, rank=rank, nRanks=world)
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        ctx.bindSharedFramebuffer(t.data_ptr(), nbytes, [int(p) for p in hdl.buffer_ptrs], mc, tr.width, rows)
        for _i, key, img in tr.images:
            ctx.putImage(key, img)
        # compact records: long runs are uploaded 1/world per rank and pushed to all ranks over NVLink (sharded upload)
        prepared = prepare_calls(tr.calls, compact=True, min_compact_run=64)
        for k in range(3):
            ctx.beginFrame((tr.width, tr.height), clearMain=tr.clear is not None, clearMainColor=tr.clear or (1, 1, 1, 1))
            if k == 1:
                ctx.submitCalls(tr.calls)  # plain 128-byte records: every rank uploads everything
            else:
                ctx.submitPrepared(prepared)
            ctx.endFrame()
            bands.resolve_across_ranks(ctx, world, dist)
        torch.cuda.synchronize()
        dist.barrier()
        results[name] = t[: tr.height * tr.width * 4].view(tr.height, tr.width, 4).cpu().numpy().copy()
        results[name + "_mc"] = np.array([mc != 0])
        dist.barrier()
        # bands of equal cost instead of equal height (every rank adopts the same boundaries), blur halos included
        bounds = bands.rebalance_across_ranks(ctx, world, (tr.width + 15) // 16, dist)
        for _k in range(2):
            ctx.beginFrame((tr.width, tr.height), clearMain=tr.clear is not None, clearMainColor=tr.clear or (1, 1, 1, 1))
            ctx.submitPrepared(prepared)
            ctx.endFrame()
            bands.resolve_across_ranks(ctx, world, dist)
        torch.cuda.synchronize()
        dist.barrier()
        results[name + "_balanced"] = t[: tr.height * tr.width * 4].view(tr.height, tr.width, 4).cpu().numpy().copy()
        results[name + "_bounds"] = np.array(bounds)
        dist.barrier()
        ctx.close()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **results)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_multicast_gather_gives_every_rank_the_whole_frame(tmp_path):
    import torch.multiprocessing as mp

    from figdraw_b200 import scenes_synth as ss
    from figdraw_b200.cuda_context import render_trace

    world = min(torch.cuda.device_count(), 8)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    want = {"cfg5": render_trace(ss.config_trace(5, 1920, 1080, n_rects=8000, n_glyphs=1500)), "cfg2": render_trace(ss.config_trace(2)),
            "cfg4": render_trace(ss.config_trace(4, 1920, 1080, rows=60, cols=8))}
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        for name, img in want.items():
            assert np.array_equal(got[name], img), f"rank {r}: {name} differs from the single-GPU frame"
            assert np.array_equal(got[name + "_balanced"], img), f"rank {r}: {name} differs under balanced bands {got[name + '_bounds']}"
        print("rank", r, "multicast mapping:", bool(got["cfg5_mc"][0]))
