"""Renders one BASELINE config and replays it a few times launch by launch (for ncu).  usage: run_cfg.py <2|3|4|5> [variant]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from figdraw_b200 import scenes_synth as ss
from figdraw_b200.cuda_context import CudaContext, render_trace

cfg = int(sys.argv[1])
kw = {}
if len(sys.argv) > 2 and sys.argv[2] == "rectmask":
    kw["rect_mask"] = True
if len(sys.argv) > 2 and sys.argv[2] == "msdf":
    kw["msdf_glyphs"] = 20000
tr = ss.config_trace(cfg, **kw)
ctx = CudaContext(atlasSize=tr.atlas_size)
render_trace(tr, ctx)
ctx.setReplayGraph(False)
for _ in range(3):
    ctx.replayFrame()
    ctx.sync()
