"""Diagnostic run for the GPU box: renders every scene through the C ABI, compares with the oracle and writes a
report (and diff images for failures) under gpurun_out/."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
from figdraw_b200 import scenes, scenes_synth as ss  # noqa: E402
from figdraw_b200.cuda_context import CudaContext, render_trace  # noqa: E402
from oracle import oracle  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
rows = []


def run(name, tr, save=True):
    t0 = time.time()
    ctx = CudaContext(atlasSize=tr.atlas_size)
    try:
        got = render_trace(tr, ctx)
        st = ctx.frameStats()
        ctx.replayFrame()
        st2 = ctx.frameStats()
    except Exception as e:  # noqa: BLE001
        rows.append({"scene": name, "error": repr(e)})
        print(name, "ERROR", e, flush=True)
        return
    t1 = time.time()
    want = oracle.render_trace(tr)
    t2 = time.time()
    d = np.abs(got.astype(np.int16) - want.astype(np.int16)).max(axis=2)
    row = {"scene": name, "size": [tr.width, tr.height], "draws": tr.n_draws, "max_diff": int(d.max()),
           "n_diff": int((d > 0).sum()), "n_gt1": int((d > 1).sum()), "n_gt2": int((d > 2).sum()),
           "gpu_ms": round(st2.gpu_ms, 4), "bin_ms": round(st2.bin_ms, 4), "shade_ms": round(st2.shade_ms, 4),
           "blur_ms": round(st2.blur_ms, 4), "entries": int(st2.n_tile_entries), "launches": st2.n_launches,
           "first_gpu_ms": round(st.gpu_ms, 4), "oracle_s": round(t2 - t1, 2), "cuda_wall_s": round(t1 - t0, 2)}
    rows.append(row)
    print(json.dumps(row), flush=True)
    if save and d.max() > 1:
        from PIL import Image

        Image.fromarray(got).save(os.path.join(OUT, f"{name}_cuda.png"))
        Image.fromarray(want).save(os.path.join(OUT, f"{name}_oracle.png"))
        Image.fromarray((np.minimum(d, 8) * 31).astype(np.uint8)).save(os.path.join(OUT, f"{name}_diff.png"))
        ys, xs = np.nonzero(d > 1)
        with open(os.path.join(OUT, f"{name}_diff.txt"), "w") as fh:
            for x, y in list(zip(xs, ys))[:200]:
                fh.write(f"{x} {y} cuda={got[y, x].tolist()} oracle={want[y, x].tolist()}\n")
    ctx.close()


for name in scenes.GOLDEN_SCENES:
    run(name, scenes.golden_trace(name))
run("rect_mask", scenes.trace_scene(scenes.layers_rect_mask, 800, 375))
run("cfg2", ss.config_trace(2))
run("cfg3_small", ss.config_trace(3, 1920, 1080, n_glyphs=6000, msdf_glyphs=500))
run("cfg4_small", ss.config_trace(4, 1920, 1080, rows=60, cols=8))
run("cfg4_small_rm", ss.config_trace(4, 1920, 1080, rows=60, cols=8, rect_mask=True))
run("cfg5_small", ss.config_trace(5, 1280, 720, n_rects=12000, n_glyphs=2400))
if "--full" in sys.argv:
    run("cfg5_4k", ss.config_trace(5), save=False)
    run("cfg4_4k", ss.config_trace(4), save=False)
    run("cfg3_4k", ss.config_trace(3), save=False)
with open(os.path.join(OUT, "report.json"), "w") as fh:
    json.dump(rows, fh, indent=1)
