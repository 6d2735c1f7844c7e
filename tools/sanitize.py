"""A small frame through every kernel of the coarse path (and one band of a 2-way partition) for compute-sanitizer:
    compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from figdraw_b200 import scenes_synth as ss  # noqa: E402
from figdraw_b200.cuda_context import CudaContext, render_trace  # noqa: E402

tr = ss.config_trace(5, 1280, 720, n_rects=3000, n_glyphs=300)
ctx = CudaContext(atlasSize=tr.atlas_size)
a = render_trace(tr, ctx)
ctx.replayFrame()
b = ctx.readPixels()
assert (a == b).all()
print("tile entries", ctx.frameStats().n_tile_entries, "row costs", int(ctx.tileRowCosts().sum()))
ctx.close()
band = CudaContext(atlasSize=tr.atlas_size, rank=1, nRanks=2)
band.setBandTileRows([0, 10, 45])
c = render_trace(tr, band)
y0, y1 = band.bandRows()
assert (c[y0:y1] == a[y0:y1]).all()
band.close()
print("sanitize scene ok")
