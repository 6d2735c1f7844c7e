"""Synthetic glyph outlines for the GPU glyph rasteriser's tests and smoke run (no font file is read anywhere on this
path: font parsing and shaping stay upstream in the reference, `common/typefaces.nim` / pixie / harfbuzz).

Outlines are closed contours of lines and quadratic Beziers in the glyph bitmap's pixel space (x right, y down), outer
contours and holes wound in opposite senses -- the form TrueType `glyf` outlines take after scaling."""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

from . import abi

Seg = Tuple[float, float, float, float, float, float, int]


def _line(a, b) -> Seg:
    return (a[0], a[1], b[0], b[1], 0.0, 0.0, 0)


def _quad(a, c, b) -> Seg:
    return (a[0], a[1], b[0], b[1], c[0], c[1], 1)


def polygon(points, reverse: bool = False) -> List[Seg]:
    pts = list(points)[::-1] if reverse else list(points)
    return [_line(pts[i], pts[(i + 1) % len(pts)]) for i in range(len(pts))]


def ellipse(cx, cy, rx, ry, n_arcs: int = 8, reverse: bool = False) -> List[Seg]:
    """Closed ellipse from `n_arcs` quadratic arcs (control points on the tangent intersections)."""
    segs = []
    for k in range(n_arcs):
        a0, a1 = 2 * math.pi * k / n_arcs, 2 * math.pi * (k + 1) / n_arcs
        if reverse:
            a0, a1 = -a0, -a1
        am = 0.5 * (a0 + a1)
        r = 1.0 / math.cos(0.5 * (a1 - a0))
        p0 = (cx + rx * math.cos(a0), cy + ry * math.sin(a0))
        p1 = (cx + rx * math.cos(a1), cy + ry * math.sin(a1))
        c = (cx + rx * r * math.cos(am), cy + ry * r * math.sin(am))
        segs.append(_quad(p0, c, p1))
    return segs


def to_array(segs: List[Seg]) -> np.ndarray:
    arr = np.zeros(len(segs), dtype=abi.OUTLINE_SEG_DTYPE)
    for i, s in enumerate(segs):
        arr[i] = (s[0], s[1], s[2], s[3], s[4], s[5], s[6], 0)
    return arr


def sample_glyphs(seed: int = 1, count: int = 12, size: Tuple[int, int] = (28, 36)):
    """`count` letter-like shapes: rings, boxes with holes, stems with bowls, stars.  Returns [(width, height, segs)]."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(count):
        w, h = int(size[0] + rng.integers(-6, 7)), int(size[1] + rng.integers(-6, 7))
        kind = k % 4
        m = 2.0 + rng.uniform(0, 1.5)
        if kind == 0:  # 'O': ring
            segs = ellipse(w / 2 + rng.uniform(-0.4, 0.4), h / 2, w / 2 - m, h / 2 - m)
            segs += ellipse(w / 2, h / 2 + rng.uniform(-0.4, 0.4), w / 2 - m - 3.3, h / 2 - m - 4.1, reverse=True)
        elif kind == 1:  # box with a rectangular counter
            segs = polygon([(m, m), (w - m, m + 0.6), (w - m - 0.3, h - m), (m + 0.2, h - m - 0.4)])
            segs += polygon([(m + 4.2, m + 5.1), (w - m - 4.4, m + 5.3), (w - m - 4.1, h - m - 6.2), (m + 4.5, h - m - 6.0)], reverse=True)
        elif kind == 2:  # 'P': stem + bowl
            segs = polygon([(m, m), (m + 4.6, m), (m + 4.6, h - m), (m, h - m)])
            segs += ellipse(m + 4.6 + (w - 2 * m - 4.6) / 2 - 1.0, m + h * 0.27, (w - 2 * m - 4.6) / 2, h * 0.22)
            segs += ellipse(m + 4.6 + (w - 2 * m - 4.6) / 2 - 1.0, m + h * 0.27, (w - 2 * m - 4.6) / 2 - 2.8, h * 0.22 - 2.9, reverse=True)
        else:  # star polygon
            n = 5 + k % 3
            pts = []
            for i in range(2 * n):
                r = (min(w, h) / 2 - m) * (1.0 if i % 2 == 0 else 0.45)
                a = math.pi * i / n + 0.3
                pts.append((w / 2 + r * math.cos(a), h / 2 + r * math.sin(a)))
            segs = polygon(pts)
        out.append((w, h, segs))
    return out


def jobs_for(glyphs, first_key: int = 70000):
    """(jobs array, segment array, keys) for CudaContext.rasterizeGlyphs."""
    jobs = np.zeros(len(glyphs), dtype=abi.GLYPH_JOB_DTYPE)
    all_segs: List[Seg] = []
    keys = []
    for i, (w, h, segs) in enumerate(glyphs):
        jobs[i] = (first_key + i, len(all_segs), len(segs), w, h)
        keys.append(first_key + i)
        all_segs += segs
    return jobs, to_array(all_segs), keys
