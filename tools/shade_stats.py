import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from figdraw_b200 import scenes_synth as ss
from figdraw_b200.cuda_context import CudaContext, render_trace
tr = ss.config_trace(5)
ctx = CudaContext(atlasSize=tr.atlas_size)
render_trace(tr, ctx)
st = ctx.frameStats()
print(json.dumps({"entries": int(st.n_tile_entries), **{k: int(v) for k, v in ctx.shadeStats().items()}}))
