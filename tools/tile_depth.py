"""Tile-list depth distribution of cfg5 (what bounds the shade kernel's critical path).  usage: tile_depth.py [W H scale]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from figdraw_b200 import scenes_synth as ss  # noqa: E402
from figdraw_b200.cuda_context import CudaContext, render_trace  # noqa: E402

W, H, scale = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160, 1.0)
tr = ss.config_trace(5, W, H, scale=scale)
ctx = CudaContext(atlasSize=tr.atlas_size)
render_trace(tr, ctx)
off, ent = ctx.debugBins(0)
cnt = np.diff(off.astype(np.int64))
tx = (W + 15) // 16
cnt2 = cnt.reshape(-1, tx)
print("tiles", cnt.size, "entries", int(cnt.sum()), "mean %.1f" % cnt.mean(), "p50", int(np.percentile(cnt, 50)), "p90", int(np.percentile(cnt, 90)),
      "p99", int(np.percentile(cnt, 99)), "max", int(cnt.max()))
print("per tile row: mean of row maxima %.1f" % cnt2.max(axis=1).mean(), "rows with max > 2x mean:", int((cnt2.max(axis=1) > 2 * cnt.mean()).sum()))
hist, edges = np.histogram(cnt, bins=[0, 1, 32, 64, 96, 128, 192, 256, 384, 512, 100000])
for h, e0, e1 in zip(hist, edges[:-1], edges[1:]):
    print("  [%d, %d): %d tiles" % (e0, e1, h))
ctx.close()
