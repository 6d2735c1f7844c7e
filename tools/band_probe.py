"""One band of an n-way partition rendered alone on one GPU (no peers, no barriers): what a rank's kernels cost by
themselves.  usage: band_probe.py [n_ranks] [width height scale]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from figdraw_b200 import scenes_synth as ss  # noqa: E402
from figdraw_b200.cuda_context import CudaContext, render_trace  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
W, H, scale = (int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else (3840, 2160, 1.0)
tr = ss.config_trace(5, W, H, scale=scale)
for r in sorted(set([0, n // 2, n - 1])):
    ctx = CudaContext(atlasSize=tr.atlas_size, rank=r, nRanks=n)
    render_trace(tr, ctx)
    ctx.setReplayGraph(False)
    sh, bn, tot = [], [], []
    for _ in range(12):
        ctx.replayFrame()
        ctx.sync()
        st = ctx.frameStats()
        sh.append(st.shade_ms); bn.append(st.bin_ms); tot.append(st.gpu_ms)
    ctx.setReplayGraph(True)
    g = []
    for _ in range(12):
        ctx.replayFrame()
        ctx.sync()
        g.append(ctx.frameStats().gpu_ms)
    print(f"rank {r}/{n} {W}x{H}: shade {np.median(sh[3:]):.4f} bin {np.median(bn[3:]):.4f} frame {np.median(tot[3:]):.4f} graph {np.median(g[3:]):.4f} ms, "
          f"tile entries {ctx.frameStats().n_tile_entries}")
    ctx.close()
