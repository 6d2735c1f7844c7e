"""For each fuzz seed whose CUDA frame differs from the oracle by more than 2 LSB, find the first call that introduces
the difference (prefixes of the stream are re-balanced: open masks / rect masks / transforms are closed)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from figdraw_b200 import scenes_fuzz  # noqa: E402
from figdraw_b200.abi import CALL_DTYPE, Op  # noqa: E402
from figdraw_b200.cuda_context import render_trace  # noqa: E402
from figdraw_b200.figbackend import Trace  # noqa: E402
from oracle import oracle  # noqa: E402


def balanced_prefix(tr, n):
    calls = tr.calls[:n]
    masks, rms, xf, begun = 0, [], 0, False
    for c in calls:
        op = int(c["op"])
        if op == Op.BEGIN_MASK:
            masks += 1; begun = True
        elif op == Op.END_MASK:
            begun = False
        elif op == Op.POP_MASK:
            masks -= 1
        elif op == Op.BEGIN_RECT_MASK:
            rms.append(1)
        elif op == Op.POP_RECT_MASK:
            rms.pop()
        elif op == Op.SAVE_TRANSFORM:
            xf += 1
        elif op == Op.RESTORE_TRANSFORM:
            xf -= 1
    tail = []
    def rec(op):
        r = np.zeros(1, dtype=CALL_DTYPE); r["op"] = int(op); tail.append(r)
    if begun:
        rec(Op.END_MASK)
    for _ in rms:
        rec(Op.POP_RECT_MASK)
    for _ in range(masks):
        rec(Op.POP_MASK)
    for _ in range(xf):
        rec(Op.RESTORE_TRANSFORM)
    t = Trace(tr.width, tr.height, tr.clear)
    t.calls = np.concatenate([calls] + tail) if tail else calls.copy()
    t.images = tr.images
    t.atlas_size = tr.atlas_size
    return t


def maxdiff(t):
    a = render_trace(t)
    b = oracle.render_trace(t)
    d = np.abs(a.astype(np.int16) - b.astype(np.int16)).max(axis=2)
    return int(d.max()), d, a, b


out = []
seeds = [int(s) for s in sys.argv[1:]] or list(range(12))
for seed in seeds:
    tr = scenes_fuzz.random_trace(seed)
    mx, d, a, b = maxdiff(tr)
    if mx <= 2:
        continue
    lo, hi = 0, len(tr.calls)  # invariant: prefix(lo) ok, prefix(hi) bad
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if maxdiff(balanced_prefix(tr, mid))[0] > 2:
            hi = mid
        else:
            lo = mid
    c = tr.calls[hi - 1]
    mx2, d2, a2, b2 = maxdiff(balanced_prefix(tr, hi))
    ys, xs = np.nonzero(d2 > 2)
    y, x = int(ys[0]), int(xs[0])
    # state ops in effect
    ctx_ops = [(int(k), Op(int(q["op"])).name) for k, q in enumerate(tr.calls[:hi - 1]) if int(q["op"]) < 32][-12:]
    row = {"seed": seed, "max_diff": mx, "first_bad_call": hi - 1, "op": Op(int(c["op"])).name, "u": [int(v) for v in c["u"]],
           "f": [round(float(v), 3) for v in c["f"][:17]], "n_bad_px": int((d2 > 2).sum()), "bbox": [int(xs.min()), int(ys.min()), int(xs.max()), int(ys.max())],
           "sample": {"xy": [x, y], "cuda": a2[y, x].tolist(), "oracle": b2[y, x].tolist()}, "recent_state_ops": ctx_ops}
    out.append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fuzz_bisect.json"), "w"), indent=1)
