"""Times every BASELINE config through the C ABI on one B200 and checks parity against the oracle.
Writes gpurun_out/configs.json."""
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

rows = []
cases = [("cfg1 rgb_boxes_sdf 800x600", lambda: scenes.golden_trace("rgb_boxes_sdf")),
         ("cfg2 renderlist_100 1920x1080", lambda: ss.config_trace(2)),
         ("cfg3 text+msdf 3840x2160", lambda: ss.config_trace(3)),
         ("cfg3 +20k msdf quads 3840x2160", lambda: ss.config_trace(3, msdf_glyphs=20000)),
         ("cfg4 clip table sub-clip 3840x2160", lambda: ss.config_trace(4)),
         ("cfg4 clip table rect-mask 3840x2160", lambda: ss.config_trace(4, rect_mask=True)),
         ("cfg5 100k rects + 20k glyphs 3840x2160", lambda: ss.config_trace(5)),
         ("cfg5 x2 7680x4320 (single GPU)", lambda: ss.config_trace(5, 7680, 4320, scale=2.0))]
for name, build in cases:
    tr = build()
    ctx = CudaContext(atlasSize=tr.atlas_size)
    got = render_trace(tr, ctx)
    times, graph_ms = [], []
    ctx.setReplayGraph(False)  # launch by launch: per-phase times
    for _ in range(int(os.environ.get("FDC_REPLAYS", "12"))):
        ctx.replayFrame()
        st = ctx.frameStats()
        times.append((st.gpu_ms, st.bin_ms, st.shade_ms, st.blur_ms))
    ctx.setReplayGraph(True)   # one CUDA-graph launch per frame
    for _ in range(int(os.environ.get("FDC_REPLAYS", "12")) + 1):
        ctx.replayFrame()
        graph_ms.append(ctx.frameStats().gpu_ms)
    t = np.median(np.array(times), axis=0)
    g = float(np.median(np.array(graph_ms[1:])))
    t0 = time.time()
    want = got if os.environ.get("FDC_SKIP_ORACLE") else oracle.render_trace(tr)
    cpu_s = time.time() - t0
    d = np.abs(got.astype(np.int16) - want.astype(np.int16)).max(axis=2)
    mpx = tr.width * tr.height / 1e6
    # algorithmic roofline of the shade kernel: fragments the reference's GL path would shade x flops per mode
    import bench as B  # noqa: E402  (repo root is on sys.path)

    counts = oracle.count_fragments(tr)
    flops = B.algorithmic_flops(counts)
    n_blur = int((tr.calls["op"] == 13).sum())
    peak = 148 * 128 * 2 * 1965e6 / 1e12
    row = {"config": name, "draws": tr.n_draws, "segments": int(st.n_segments), "launches": int(st.n_launches),
           "gpu_ms": round(float(t[0]), 4), "graph_ms": round(g, 4), "bin_ms": round(float(t[1]), 4), "shade_ms": round(float(t[2]), 4),
           "blur_ms": round(float(t[3]), 4), "mpix_per_s": round(mpx / (t[0] * 1e-3), 1), "fps": round(1e3 / t[0], 1),
           "fragments": int(counts.sum()), "algorithmic_gflop": round(flops / 1e9, 3),
           "shade_algorithmic_tflops": round(flops / (t[2] * 1e-3) / 1e12, 2) if t[2] > 0 else None,
           "shade_frac_of_fp32_peak_at_1965MHz": round(flops / (t[2] * 1e-3) / 1e12 / peak, 3) if t[2] > 0 else None,
           "blur_nodes": n_blur,
           "cpu_oracle_s": round(cpu_s, 3), "cpu_threads": oracle.max_threads(), "max_diff_lsb": int(d.max()),
           "pixels_differing": int((d > 0).sum())}
    rows.append(row)
    print(json.dumps(row), flush=True)
    ctx.close()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
