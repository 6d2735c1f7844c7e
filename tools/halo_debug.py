"""Debug aid: banded blur (peer halo reads) vs single-context frame on fuzz traces; prints where they differ."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from figdraw_b200.abi import Op
from figdraw_b200.cuda_context import CudaContext, render_trace
from figdraw_b200.scenes_fuzz import random_trace
import test_gpu_parity as T

for seed in range(14):
    tr = random_trace(seed)
    nb = int((tr.calls["op"] == Op.BACKDROP_BLUR).sum())
    if not nb:
        continue
    full = T._render_banded_with_peers(tr, 1)[0]
    for n in (2, 3):
        imgs = T._render_banded_with_peers(tr, n)
        for r, img in enumerate(imgs):
            d = np.abs(img.astype(int) - full.astype(int)).max(axis=2)
            if d.max() == 0:
                continue
            ys, xs = np.nonzero(d)
            print(f"seed {seed} blurs {nb} n {n} rank {r}: {len(ys)} px, max {d.max()}, rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()}",
                  "rows hist", np.bincount(ys // 16, minlength=24).tolist())
    blur = tr.calls[tr.calls["op"] == Op.BACKDROP_BLUR]
    print(f"seed {seed}: blur rects", [(c["f"][:4].round(1).tolist(), float(c["f"][12])) for c in blur][:8])
