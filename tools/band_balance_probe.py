"""Equal bands vs bands balanced by the tile-entry profile: every band of an n-way partition rendered alone on one GPU
(no peers), slowest band reported.  usage: band_balance_probe.py <cfg 3|5> [n_ranks]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from figdraw_b200 import scenes_synth as ss  # noqa: E402
from figdraw_b200.bands import balance_rows  # noqa: E402
from figdraw_b200.cuda_context import CudaContext, render_trace  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
tr = ss.config_trace(cfg)
one = CudaContext(atlasSize=tr.atlas_size)
render_trace(tr, one)
profile = one.tileRowCosts().astype(np.int64)
one.close()
tiles_x = (tr.width + 15) // 16
balanced = balance_rows(profile + 3 * tiles_x, n)


def run(bounds):
    ms = []
    for r in range(n):
        ctx = CudaContext(atlasSize=tr.atlas_size, rank=r, nRanks=n)
        if bounds is not None:
            ctx.setBandTileRows(bounds)
        render_trace(tr, ctx)
        g = []
        for _ in range(10):
            ctx.replayFrame()
            ctx.sync()
            g.append(ctx.frameStats().gpu_ms)
        ms.append(float(np.median(g[3:])))
        ctx.close()
    return ms


eq, bal = run(None), run(balanced)
print(f"cfg{cfg} {tr.width}x{tr.height}, {n} bands: equal height: slowest {max(eq):.4f} ms (mean {np.mean(eq):.4f}); "
      f"balanced {balanced}: slowest {max(bal):.4f} ms (mean {np.mean(bal):.4f})")
