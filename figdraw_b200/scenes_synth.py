"""Synthetic scenes of BASELINE.json's configs 2..5 (shapes: SURVEY.md section 8d).

All randomness comes from one counter-based generator (SplitMix64 -> uniform float32), so any host language
can regenerate identical bytes from (seed, counter).  cfg2-cfg4 build `Renders` trees and go through the
front-end (figrender.py); cfg5 (245k primitives) emits the same backend calls the front-end would, vectorised
with numpy, because 100k Python `Fig` objects would take minutes to walk.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

from .abi import CALL_DTYPE, FillKindAbi, Op, SdfMode
from .figbackend import Trace, TraceBackend
from .fignodes import (BackdropBlurStyle, Fig, FigFlags, FigKind, FillGradientAxis, Glyph, MsdfImageStyle, RenderList,
                       RenderShadow, RenderStroke, Renders, ShadowStyle, fill, linear, rect, rgba)
from .figrender import renderFrame, setFigUiScale

f32 = np.float32
MASK64 = (1 << 64) - 1


# ----------------------------------------------------------------------------- RNG
def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x.copy()
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


class Rng:
    """uniform(i) is a pure function of (seed, stream, i)."""

    def __init__(self, seed: int):
        self.seed = np.uint64(seed)
        self.stream = 0

    def uniform(self, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
        self.stream += 1
        with np.errstate(over="ignore"):
            ctr = np.arange(n, dtype=np.uint64) + np.uint64(self.stream) * np.uint64(0x1_0000_0000) + self.seed * np.uint64(0x9E3779B1)
            bits = splitmix64(ctr)
        u = (bits >> np.uint64(40)).astype(np.float32) * f32(1.0 / (1 << 24))
        return (f32(lo) + u * f32(hi - lo)).astype(np.float32)

    def integers(self, n: int, lo: int, hi: int) -> np.ndarray:  # [lo, hi)
        return np.minimum((self.uniform(n) * f32(hi - lo)).astype(np.int64) + lo, hi - 1)


def pack_rgba(r, g, b, a) -> np.ndarray:
    r, g, b, a = (np.asarray(v).astype(np.uint32) & np.uint32(255) for v in (r, g, b, a))
    return r | (g << np.uint32(8)) | (b << np.uint32(16)) | (a << np.uint32(24))


def _nim_round(x: np.ndarray) -> np.ndarray:
    return (np.sign(x) * np.floor(np.abs(x) + f32(0.5))).astype(np.float32)


# ----------------------------------------------------------------------------- synthetic atlas content
GLYPH_W, GLYPH_H, N_GLYPHS = 12, 24, 95
GLYPH_KEY0 = 0x67_6C_79_00_00  # glyph keys are GLYPH_KEY0 + index


def make_glyph_bitmaps(seed: int = 3) -> List[np.ndarray]:
    """95 procedural 12x24 coverage bitmaps (white, alpha = coverage): stand-ins for pixie's glyph rasters
    (common/textrasters/pixie_raster.nim:45-95, out of scope).  Strokes of random segments, anti-aliased."""
    rng = Rng(seed * 7919 + 11)
    out = []
    yy, xx = np.mgrid[0:GLYPH_H, 0:GLYPH_W].astype(np.float32)
    xx += f32(0.5)
    yy += f32(0.5)
    for g in range(N_GLYPHS):
        p = rng.uniform(12)
        cov = np.zeros((GLYPH_H, GLYPH_W), dtype=np.float32)
        for s in range(3):
            ax, ay = 1.5 + p[4 * s] * 9.0, 3.0 + p[4 * s + 1] * 18.0
            bx, by = 1.5 + p[4 * s + 2] * 9.0, 3.0 + p[4 * s + 3] * 18.0
            dx, dy = bx - ax, by - ay
            den = max(dx * dx + dy * dy, 1e-3)
            t = np.clip(((xx - ax) * dx + (yy - ay) * dy) / den, 0.0, 1.0)
            d = np.sqrt((xx - (ax + t * dx)) ** 2 + (yy - (ay + t * dy)) ** 2)
            cov = np.maximum(cov, np.clip(1.3 - d, 0.0, 1.0))
        img = np.full((GLYPH_H, GLYPH_W, 4), 255, dtype=np.uint8)
        img[..., 3] = np.round(cov * 255.0).astype(np.uint8)
        out.append(img)
    return out


STAR_MSDF_KEY, STAR_MTSDF_KEY = 0x5354_4152_0001, 0x5354_4152_0002


def make_star_field(size: int = 32, px_range: float = 4.0, mtsdf: bool = False) -> np.ndarray:
    """A 5-point star as a multi-channel signed distance field (shape: examples/windy_msdf_star.nim:89-267,
    whose field comes from `sdfy`, not vendored).  Edges are coloured in turn; channel c holds the signed
    distance to the nearest edge carrying colour c, alpha (MTSDF) the true distance."""
    c = size / 2.0
    ro, ri = size * 0.40, size * 0.17
    pts = []
    for k in range(10):
        ang = -math.pi / 2 + k * math.pi / 5
        r = ro if k % 2 == 0 else ri
        pts.append((c + r * math.cos(ang), c + r * math.sin(ang)))
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float64)
    xx += 0.5
    yy += 0.5
    inside = np.zeros((size, size), dtype=bool)
    dist_c = [np.full((size, size), 1e9) for _ in range(3)]
    dist_all = np.full((size, size), 1e9)
    colours = [(0, 1), (1, 2), (0, 2)]  # channel pairs per edge, cycling
    for k in range(10):
        (ax, ay), (bx, by) = pts[k], pts[(k + 1) % 10]
        crosses = ((ay > yy) != (by > yy)) & (xx < (bx - ax) * (yy - ay) / (by - ay + 1e-30) + ax)
        inside ^= crosses
        dx, dy = bx - ax, by - ay
        t = np.clip(((xx - ax) * dx + (yy - ay) * dy) / (dx * dx + dy * dy), 0.0, 1.0)
        d = np.hypot(xx - (ax + t * dx), yy - (ay + t * dy))
        dist_all = np.minimum(dist_all, d)
        for ch in colours[k % 3]:
            dist_c[ch] = np.minimum(dist_c[ch], d)
    sign = np.where(inside, 1.0, -1.0)
    img = np.zeros((size, size, 4), dtype=np.uint8)
    for ch in range(3):
        img[..., ch] = np.clip(np.round((0.5 + sign * dist_c[ch] / px_range) * 255.0), 0, 255).astype(np.uint8)
    img[..., 3] = np.clip(np.round((0.5 + sign * dist_all / px_range) * 255.0), 0, 255).astype(np.uint8) if mtsdf else 255
    return img


# ----------------------------------------------------------------------------- cfg2: renderlist_100 shape
def renderlist_100(w: float, h: float, copies: int = 100, seed: int = 12345, with_blur: bool = True) -> Renders:
    """Shape of examples/renderlist_100_common.nim:11-251 at frame 0 (t = 0); positions from SplitMix64 instead of
    Nim's std/random stream."""
    w, h = f32(w), f32(h)
    lst = RenderList()
    lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(0, 0, w, h), fill=rgba(255, 255, 255, 155)))
    maxW, maxH = f32(260.0), f32(180.0)
    maxX = max(f32(0), w - (f32(320.0) + maxW))
    maxY = max(f32(0), h - (f32(300.0) + maxH))
    rng = Rng(seed)
    bx, by = rng.uniform(copies, 0.0, float(maxX)), rng.uniform(copies, 0.0, float(maxY))
    sin, cos = (lambda v: f32(math.sin(float(v)))), (lambda v: f32(math.cos(float(v))))
    for i in range(copies):
        fi = f32(i)
        ox = min(max(bx[i] + sin(fi * f32(0.15)) * 20, f32(0)), maxX)
        oy = min(max(by[i] + cos(fi * f32(0.2)) * 20, f32(0)), maxY)
        pw = f32(0.5) + f32(0.5) * sin(fi * f32(0.07))
        ph = f32(0.5) + f32(0.5) * cos(fi * f32(0.09))
        redW, redH = f32(160) + f32(100) * pw, f32(110) + f32(70) * ph
        greenW, greenH = f32(160) + f32(100) * ph, f32(110) + f32(70) * pw
        blueW, blueH = f32(160) + f32(100) * (f32(1) - pw), f32(110) + f32(70) * (f32(1) - ph)
        cp = f32(0.5) + f32(0.5) * sin(fi * f32(0.11))
        c0, c1 = f32(4) + f32(26) * cp, f32(6) + f32(22) * (f32(1) - cp)
        c2 = f32(8) + f32(18) * (f32(0.5) + f32(0.5) * sin(fi * f32(0.05)))
        c3 = f32(10) + f32(16) * (f32(0.5) + f32(0.5) * cos(fi * f32(0.06)))
        gp = f32(0.5) + f32(0.5) * cos(fi * f32(0.08))
        g0, g1 = f32(6) + f32(22) * gp, f32(8) + f32(18) * (f32(1) - gp)
        g2 = f32(10) + f32(16) * (f32(0.5) + f32(0.5) * cos(fi * f32(0.04)))
        g3 = f32(12) + f32(14) * (f32(0.5) + f32(0.5) * sin(fi * f32(0.05)))
        sp = f32(0.5) + f32(0.5) * sin(fi * f32(0.05))
        sBlur, sSpread = max(f32(0), f32(6) + f32(18) * sp), max(f32(0), f32(4) + f32(20) * (f32(1) - sp))
        sX, sY = f32(6) + f32(10) * sin(fi * f32(0.03)), f32(6) + f32(10) * cos(fi * f32(0.03))
        ip = f32(0.5) + f32(0.5) * sin(fi * f32(0.06))
        iBlur, iSpread = max(f32(0), f32(8) + f32(10) * ip), max(f32(0), f32(2) + f32(10) * (f32(1) - ip))
        iX, iY = f32(6) * sin(fi * f32(0.04)), f32(6) * cos(fi * f32(0.04))
        greenGrad, blueGrad = (i % 2) == 0, (i % 3) == 0
        lst.addRoot(Fig(kind=FigKind.nkRectangle, corners=(int(c0), int(c1), int(c2), int(c3)),
                        cornerRadiiY=(int(c0), int(c1 * 2), int(c2), int(c3 * 2)), flags=FigFlags.NfEllipticalCorners,
                        screenBox=rect(f32(60) + ox, f32(60) + oy, redW, redH), fill=rgba(220, 40, 40, 155),
                        stroke=RenderStroke(weight=5.0, fill=rgba(0, 0, 0, 155))))
        lst.addRoot(Fig(
            kind=FigKind.nkRectangle, screenBox=rect(f32(320) + ox, f32(120) + oy, greenW, greenH),
            corners=(int(g0), int(g1), int(g2), int(g3)),
            fill=(linear(rgba(18, 112, 64, 255), rgba(40, 180, 90, 255), rgba(78, 224, 188, 255),
                         axis=FillGradientAxis.fgaX if (i % 4) < 2 else FillGradientAxis.fgaDiagTLBR, midPos=128)
                  if greenGrad else rgba(40, 180, 90, 155)),
            shadows=[RenderShadow(style=ShadowStyle.DropShadow, blur=sBlur, spread=sSpread, x=sX, y=sY,
                                  fill=rgba(0, 0, 0, 155))]))
        lst.addRoot(Fig(
            kind=FigKind.nkRectangle, screenBox=rect(f32(180) + ox, f32(300) + oy, blueW, blueH),
            fill=(linear(rgba(44, 72, 186, 255), rgba(60, 90, 220, 255), rgba(118, 168, 255, 255),
                         axis=FillGradientAxis.fgaY if (i % 2) == 0 else FillGradientAxis.fgaDiagBLTR, midPos=132)
                  if blueGrad else rgba(60, 90, 220, 155)),
            stroke=RenderStroke(weight=4.0, fill=rgba(255, 255, 255, 210)),
            shadows=[RenderShadow(style=ShadowStyle.InnerShadow, blur=iBlur, spread=iSpread, x=iX, y=iY,
                                  fill=(linear(rgba(25, 25, 40, 100), rgba(65, 65, 95, 180),
                                               axis=FillGradientAxis.fgaDiagBLTR) if blueGrad
                                        else rgba(40, 40, 60, 150)))]))
    lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(max(f32(20), w - f32(200)), 20, 180, 100),
                    fill=rgba(238, 140, 30, 220), corners=(90, 90, 90, 90), cornerRadiiY=(50, 50, 50, 50),
                    flags=FigFlags.NfEllipticalCorners, stroke=RenderStroke(weight=4.0, fill=rgba(90, 45, 0, 220))))
    yW, yH, yM = f32(360), f32(240), f32(20)
    yX = yM + max(f32(0), w - yW - yM * 2) * f32(0.5)
    yY = yM + max(f32(0), h - yH - yM * 2) * f32(1.0)
    if with_blur:
        lst.addRoot(Fig(kind=FigKind.nkBackdropBlur, corners=(20, 20, 20, 20), screenBox=rect(yX, yY, yW, yH),
                        fill=rgba(0, 0, 0, 0), backdropBlur=BackdropBlurStyle(blur=18.0)))
    lst.addRoot(Fig(kind=FigKind.nkRectangle, corners=(20, 20, 20, 20), screenBox=rect(yX, yY, yW, yH),
                    fill=rgba(255, 225, 55, 120), stroke=RenderStroke(weight=6.0, fill=rgba(95, 72, 0, 185))))
    r = Renders()
    r.setLayer(0, lst)
    return r


# ----------------------------------------------------------------------------- cfg3: text page + MSDF star
def text_page(w: float, h: float, n_glyphs: int = 20000, seed: int = 3, msdf_glyphs: int = 0) -> Renders:
    cols = max(1, int(w) // GLYPH_W)
    rng = Rng(seed)
    which = rng.integers(n_glyphs, 0, N_GLYPHS)
    fills = [fill(rgba(235, 235, 235, 255)), fill(rgba(120, 200, 255, 255)), fill(rgba(255, 180, 90, 230)),
             fill(rgba(160, 255, 160, 255)), fill(rgba(255, 255, 255, 140)), fill(rgba(250, 120, 160, 255)),
             linear(rgba(255, 90, 90, 255), rgba(90, 120, 255, 255), axis=FillGradientAxis.fgaX),
             linear(rgba(255, 240, 120, 255), rgba(120, 255, 200, 200), axis=FillGradientAxis.fgaY)]
    span = rng.integers(n_glyphs, 0, len(fills))
    lst = RenderList()
    root = lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(0, 0, w, h), fill=rgba(24, 26, 32, 255)))
    per_node = 4096
    for n0 in range(0, n_glyphs, per_node):
        glyphs = []
        for i in range(n0, min(n0 + per_node, n_glyphs)):
            glyphs.append(Glyph(key=GLYPH_KEY0 + int(which[i]), pos=(float((i % cols) * GLYPH_W), float((i // cols) * GLYPH_H)),
                                fill=fills[int(span[i])]))
        lst.addChild(root, Fig(kind=FigKind.nkText, screenBox=rect(0, 0, w, h), glyphs=glyphs))
    s = f32(min(w, h) * 0.55)
    sx, sy = f32(w) * f32(0.5) - s * f32(0.5), f32(h) * f32(0.5) - s * f32(0.5)
    lst.addChild(root, Fig(kind=FigKind.nkMtsdfImage, screenBox=rect(sx + 14, sy + 18, s, s),
                           mtsdfImage=MsdfImageStyle(id=STAR_MTSDF_KEY, fill=rgba(0, 0, 0, 110), pxRange=4.0, sdThreshold=0.42)))
    lst.addChild(root, Fig(kind=FigKind.nkMsdfImage, screenBox=rect(sx, sy, s, s),
                           msdfImage=MsdfImageStyle(id=STAR_MSDF_KEY, fill=rgba(255, 212, 48, 235), pxRange=4.0)))
    lst.addChild(root, Fig(kind=FigKind.nkMtsdfImage, screenBox=rect(sx, sy, s, s),
                           mtsdfImage=MsdfImageStyle(id=STAR_MTSDF_KEY, fill=rgba(120, 70, 0, 255), pxRange=4.0,
                                                     strokeWeight=6.0)))
    if msdf_glyphs:
        px, py = rng.uniform(msdf_glyphs, 0.0, float(w) - 16.0), rng.uniform(msdf_glyphs, 0.0, float(h) - 16.0)
        for i in range(msdf_glyphs):
            lst.addChild(root, Fig(kind=FigKind.nkMsdfImage, screenBox=rect(px[i], py[i], 16, 16),
                                   msdfImage=MsdfImageStyle(id=STAR_MSDF_KEY, fill=rgba(200, 220, 255, 255), pxRange=4.0)))
    r = Renders()
    r.setLayer(0, lst)
    return r


def text_page_images():
    imgs = [(GLYPH_KEY0 + i, g) for i, g in enumerate(make_glyph_bitmaps())]
    imgs.append((STAR_MSDF_KEY, make_star_field(32, 4.0, mtsdf=False)))
    imgs.append((STAR_MTSDF_KEY, make_star_field(32, 4.0, mtsdf=True)))
    return imgs


# ----------------------------------------------------------------------------- cfg4: clip-mask table
def clip_mask_table(w: float, h: float, rows: int = 180, cols: int = 12, rect_mask: bool = False, blurs: bool = True) -> Renders:
    """Shape of examples/windy_clip_mask_benchmark.nim:147-184 (table of clipped cells with overflowing
    children), plus two backdrop-blur panels and an overlay layer, with 3-stop gradient cells."""
    w, h = f32(w), f32(h)
    margin, gap, cellH, scrollY = f32(22), f32(4), f32(22), f32(37)
    vx, vy, vw, vh = margin, margin, w - margin * 2, h - margin * 2
    cellW = (vw - gap * f32(cols + 1)) / f32(cols)
    bg = RenderList()
    bg.addRoot(Fig(kind=FigKind.nkRectangle, zlevel=-20, screenBox=rect(0, 0, w, h), fill=rgba(248, 249, 251, 255)))
    lst = RenderList()
    vp = lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(vx, vy, vw, vh), fill=rgba(232, 235, 240, 255),
                         flags=FigFlags.NfClipContent, corners=(10, 10, 10, 10)))
    cflag = FigFlags.NfRectMaskContent if rect_mask else FigFlags.NfClipContent
    for row in range(rows):
        y = vy + gap + f32(row) * (cellH + gap) - scrollY
        for col in range(cols):
            x = vx + gap + f32(col) * (cellW + gap)
            base = rgba(255, 255, 255, 255) if (row + col) % 2 == 0 else rgba(242, 246, 250, 255)
            cfill = linear(base, rgba(226, 236, 250, 255), rgba(200, 216, 244, 255),
                           axis=FillGradientAxis(col % 4), midPos=96 + (row * 5) % 64)
            cell = lst.addChild(vp, Fig(kind=FigKind.nkRectangle, screenBox=rect(x, y, cellW, cellH), fill=cfill,
                                        flags=cflag, corners=(4, 4, 4, 4)))
            tone = 42 + (row * 7 + col * 17) % 72
            accent = rgba(36, 120 + (row * 5) % 80, 235, 255)
            spill = rgba(tone, 170 - (col * 11) % 70, 220, 255)
            muted = rgba(190 + (row + col) % 30, 210, 220, 255)
            lst.addChild(cell, Fig(kind=FigKind.nkRectangle, screenBox=rect(x - 12, y + 4, cellW + 24, 5), fill=accent,
                                   corners=(2, 2, 2, 2)))
            lst.addChild(cell, Fig(kind=FigKind.nkRectangle,
                                   screenBox=rect(x + cellW * f32(0.38), y - 5, cellW * f32(0.74), cellH + 10),
                                   fill=spill, corners=(3, 3, 3, 3)))
            lst.addChild(cell, Fig(kind=FigKind.nkRectangle, screenBox=rect(x + 7, y + cellH - 7, cellW - 14, 8),
                                   fill=muted, corners=(2, 2, 2, 2)))
    if blurs:
        for k, radius in enumerate((18.0, 40.0)):
            px, py = w * f32(0.18 + 0.42 * k), h * f32(0.30 + 0.25 * k)
            lst.addRoot(Fig(kind=FigKind.nkBackdropBlur, corners=(24, 24, 24, 24), screenBox=rect(px, py, 600, 400),
                            fill=rgba(255, 255, 255, 40), backdropBlur=BackdropBlurStyle(blur=radius)))
    top = RenderList()
    for k in range(64):
        bx, by = f32(60 + (k % 16) * 230), f32(80 + (k // 16) * 520)
        top.addRoot(Fig(kind=FigKind.nkRectangle, zlevel=20, screenBox=rect(bx, by, 180, 44), corners=(12, 12, 12, 12),
                        fill=rgba(43, 159, 234, 235), stroke=RenderStroke(weight=2.0, fill=rgba(20, 90, 160, 255))))
    r = Renders()
    r.setLayer(-20, bg)
    r.setLayer(0, lst)
    r.setLayer(20, top)
    return r


# ----------------------------------------------------------------------------- cfg5: 100k shadowed rects + 20k glyphs
def _rounded_rect_calls(n: int) -> np.ndarray:
    c = np.zeros(n, dtype=CALL_DTYPE)
    c["op"] = int(Op.ROUNDED_RECT)
    return c


def _cfg5_params(width: int, height: int, n_rects: int, seed: int, scale: float):
    """The random node parameters of cfg5, shared by the call-level generator and the node-level one."""
    rng = Rng(seed)
    S = f32(scale)
    W, H = f32(width), f32(height)
    w = rng.uniform(n_rects, 16.0, 80.0) * S
    h = rng.uniform(n_rects, 12.0, 52.0) * S
    x = rng.uniform(n_rects, 0.0, 1.0) * (W - w)
    y = rng.uniform(n_rects, 0.0, 1.0) * (H - h)
    half = np.minimum(w, h) * f32(0.5)
    radii = np.stack([np.floor(rng.uniform(n_rects) * half) for _ in range(4)], axis=1).astype(np.float32)  # u16 corners
    cr, cg, cb = (rng.integers(n_rects, 20, 256) for _ in range(3))
    alpha = np.where(rng.uniform(n_rects) < 0.5, 155, 255)
    grad = rng.uniform(n_rects) < 0.3
    axis = rng.integers(n_rects, 0, 4)
    blur = rng.uniform(n_rects, 2.0, 10.0) * S
    spread = rng.uniform(n_rects, 0.0, 6.0) * S
    sx, sy = rng.uniform(n_rects, -6.0, 6.0) * S, rng.uniform(n_rects, -6.0, 6.0) * S
    stroke = rng.uniform(n_rects) < 0.25

    base_col = pack_rgba(cr, cg, cb, alpha)
    mid_col = pack_rgba((cr + 40) % 256, (cg + 90) % 256, cb, 255)
    stop_col = pack_rgba(cb, cr, (cg + 128) % 256, 255)

    return dict(rng=rng, S=S, w=w, h=h, x=x, y=y, radii=radii, cr=cr, cg=cg, cb=cb, alpha=alpha, grad=grad, axis=axis, blur=blur,
                spread=spread, sx=sx, sy=sy, stroke=stroke, base_col=base_col, mid_col=mid_col, stop_col=stop_col)


def rects_and_glyphs(width: int, height: int, n_rects: int = 100_000, n_glyphs: int = 20_000, seed: int = 5,
                     scale: float = 1.0, layers: int = 4) -> Trace:
    """cfg5: n_rects rounded rects, each with a drop shadow, 25 % with a 2 px stroke, 30 % with a
    3-stop gradient, plus n_glyphs atlas glyph quads, spread over `layers` z-levels in emission order.
    Emits exactly the calls figrender.nim makes per node: drop shadow (:654-689) -> fill -> stroke (:806-873);
    text: save/translate, one drawImage per glyph, restore (:417-497)."""
    P = _cfg5_params(width, height, n_rects, seed, scale)
    rng, S, w, h, x, y, radii = P["rng"], P["S"], P["w"], P["h"], P["x"], P["y"], P["radii"]
    cr, cg, cb, grad, axis, blur, spread = P["cr"], P["cg"], P["cb"], P["grad"], P["axis"], P["blur"], P["spread"]
    sx, sy, stroke, base_col, mid_col, stop_col = P["sx"], P["sy"], P["stroke"], P["base_col"], P["mid_col"], P["stop_col"]

    # drop shadow quads
    sh = _rounded_rect_calls(n_rects)
    pad = np.maximum(_nim_round(spread) + _nim_round(f32(1.5) * blur), f32(0))
    sh["f"][:, 0] = (x + sx) - pad
    sh["f"][:, 1] = (y + sy) - pad
    sh["f"][:, 2] = w + f32(2) * pad
    sh["f"][:, 3] = h + f32(2) * pad
    sh["f"][:, 4:8] = radii
    sh["f"][:, 8:12] = radii
    sh["f"][:, 12] = blur
    sh["f"][:, 13] = spread
    sh["f"][:, 14] = w
    sh["f"][:, 15] = h
    sh["f"][:, 16] = 0.5
    sh["u"][:, 0] = int(SdfMode.sdfModeDropShadow)
    sh["u"][:, 1] = int(FillKindAbi.COLOR)
    sh["u"][:, 3] = pack_rgba(0, 0, 0, 90)

    fl = _rounded_rect_calls(n_rects)
    fl["f"][:, 0], fl["f"][:, 1], fl["f"][:, 2], fl["f"][:, 3] = x, y, w, h
    fl["f"][:, 4:8] = radii
    fl["f"][:, 8:12] = radii
    fl["f"][:, 12] = 4.0
    fl["u"][:, 0] = int(SdfMode.sdfModeClipAA)
    fl["u"][:, 1] = np.where(grad, int(FillKindAbi.LINEAR3), int(FillKindAbi.COLOR))
    fl["u"][:, 2] = np.where(grad, axis, 0)
    fl["u"][:, 3] = np.where(grad, pack_rgba(cr, cg, cb, 255), base_col)
    fl["u"][:, 4] = np.where(grad, mid_col, 0)
    fl["u"][:, 5] = np.where(grad, stop_col, 0)
    fl["f"][:, 16] = np.where(grad, f32(128.0 / 255.0), f32(0.5))

    st = _rounded_rect_calls(n_rects)
    st["f"][:, 0:12] = fl["f"][:, 0:12]
    st["f"][:, 12] = f32(2.0) * S
    st["f"][:, 16] = 0.5
    st["u"][:, 0] = int(SdfMode.sdfModeAnnularAA)
    st["u"][:, 1] = int(FillKindAbi.COLOR)
    st["u"][:, 3] = pack_rgba(cb // 3, cr // 3, cg // 3, 255)

    # interleave per node: shadow, fill, [stroke]
    per = np.stack([sh, fl, st], axis=1)  # [n, 3]
    keep = np.ones((n_rects, 3), dtype=bool)
    keep[:, 2] = stroke

    # glyph quads (1:1 texels)
    gw, gh = GLYPH_W, GLYPH_H  # atlas bitmaps are not scaled: glyph rasters are produced at the final pixel size
    cols = max(1, int(width) // gw)
    which = rng.integers(n_glyphs, 0, N_GLYPHS)
    gc = _rounded_rect_calls(n_glyphs)
    gc["op"] = int(Op.IMAGE)
    keys = (GLYPH_KEY0 + which).astype(np.uint64)
    gc["u"][:, 0] = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    gc["u"][:, 1] = (keys >> np.uint64(32)).astype(np.uint32)
    gi = np.arange(n_glyphs)
    gc["f"][:, 0] = (gi % cols) * gw
    gc["f"][:, 1] = (gi // cols) * gh
    gcol = pack_rgba(rng.integers(n_glyphs, 0, 256), rng.integers(n_glyphs, 0, 256), rng.integers(n_glyphs, 0, 256), 255)
    for k in range(4):
        gc["u"][:, 3 + k] = gcol

    tb = TraceBackend(atlasSize=2048)
    for key, img in [(GLYPH_KEY0 + i, g) for i, g in enumerate(make_glyph_bitmaps())]:
        tb.putImage(key, img)
    tb.beginFrame((width, height), clearMain=True)
    tb.saveTransform()
    tb.scale(1.0)
    from .figbackend import solid

    tb.drawRoundedRectSdf((0.0, 0.0, float(width), float(height)), solid(rgba(250, 250, 252, 255)),
                          ((0, 0, 0, 0), (0, 0, 0, 0)))
    for L in range(layers):
        r0, r1 = n_rects * L // layers, n_rects * (L + 1) // layers
        tb.extend(per[r0:r1][keep[r0:r1]])
        g0, g1 = n_glyphs * L // layers, n_glyphs * (L + 1) // layers
        if g1 > g0:
            tb.saveTransform()
            tb.translate((0.0, 0.0))
            tb.extend(gc[g0:g1])
            tb.restoreTransform()
    tb.restoreTransform()
    tb.endFrame()
    return tb.trace()


def rects_and_glyphs_scene(width: int, height: int, n_rects: int = 100_000, n_glyphs: int = 20_000, seed: int = 5,
                           scale: float = 1.0, layers: int = 4):
    """cfg5 as a SCENE: the `Fig` records (one nkRectangle per rect with its drop shadow and stroke, one nkText per layer)
    whose flattening is exactly the call stream of `rects_and_glyphs`.  Built straight into the POD arrays of the native
    front-end (abi.FIG_DTYPE) -- 100k Python `Fig` objects would take longer to build than a thousand frames to render."""
    import ctypes

    from . import abi
    from .native_scene import PackedScene

    P = _cfg5_params(width, height, n_rects, seed, scale)
    rng, S = P["rng"], P["S"]
    nodes = np.zeros(n_rects, dtype=abi.FIG_DTYPE)
    nodes["kind"] = 2  # nkRectangle
    nodes["parent"] = -1
    nodes["screen_box"] = np.stack([P["x"], P["y"], P["w"], P["h"]], axis=1)
    nodes["corners"] = P["radii"]
    nodes["corner_radii_y"] = P["radii"]
    g = P["grad"]
    nodes["fill"]["kind"] = np.where(g, 2, 0)
    nodes["fill"]["axis"] = np.where(g, P["axis"], 0)
    nodes["fill"]["mid_pos"] = 128
    nodes["fill"]["c"][:, 0] = np.where(g, pack_rgba(P["cr"], P["cg"], P["cb"], 255), P["base_col"])
    nodes["fill"]["c"][:, 1] = np.where(g, P["mid_col"], 0)
    nodes["fill"]["c"][:, 2] = np.where(g, P["stop_col"], 0)
    pay = nodes["payload"].view(abi.FIG_RECT_DTYPE).reshape(n_rects)
    sh = pay["shadows"][:, 0]
    sh["style"] = 1  # DropShadow
    sh["fill"]["mid_pos"] = 128
    sh["fill"]["c"][:, 0] = pack_rgba(0, 0, 0, 90)
    # the call-level generator applies `scale` to the already-scaled values; uiScale stays 1
    sh["blur"], sh["spread"], sh["x"], sh["y"] = P["blur"], P["spread"], P["sx"], P["sy"]
    pay["stroke"]["weight"] = np.where(P["stroke"], f32(2.0) * S, f32(0.0))
    pay["stroke"]["fill"]["mid_pos"] = 128
    pay["stroke"]["fill"]["c"][:, 0] = pack_rgba(P["cb"] // 3, P["cr"] // 3, P["cg"] // 3, 255)

    gw, gh = GLYPH_W, GLYPH_H
    cols = max(1, int(width) // gw)
    which = rng.integers(n_glyphs, 0, N_GLYPHS)
    glyphs = np.zeros(max(n_glyphs, 1), dtype=abi.GLYPH_DTYPE)
    gi = np.arange(n_glyphs)
    glyphs["key"][:n_glyphs] = (GLYPH_KEY0 + which).astype(np.uint64)
    glyphs["pos"][:n_glyphs, 0] = (gi % cols) * gw
    glyphs["pos"][:n_glyphs, 1] = (gi // cols) * gh
    glyphs["fill"]["mid_pos"] = 128
    glyphs["fill"]["c"][:n_glyphs, 0] = pack_rgba(rng.integers(n_glyphs, 0, 256), rng.integers(n_glyphs, 0, 256),
                                                  rng.integers(n_glyphs, 0, 256), 255)

    node_arrays, root_arrays = [], []
    for L in range(layers):
        r0, r1 = n_rects * L // layers, n_rects * (L + 1) // layers
        g0, g1 = n_glyphs * L // layers, n_glyphs * (L + 1) // layers
        extra = (1 if L == 0 else 0) + (1 if g1 > g0 else 0)
        arr = np.zeros(r1 - r0 + extra, dtype=abi.FIG_DTYPE)
        k = 0
        if L == 0:  # background
            arr[0]["kind"], arr[0]["parent"] = 2, -1
            arr[0]["screen_box"] = (0.0, 0.0, float(width), float(height))
            arr[0]["fill"]["mid_pos"] = 128
            arr[0]["fill"]["c"][0] = rgba(250, 250, 252, 255)
            k = 1
        arr[k : k + r1 - r0] = nodes[r0:r1]
        if g1 > g0:
            t = arr[-1]
            t["kind"], t["parent"] = 1, -1  # nkText
            t["fill"]["mid_pos"] = 128
            tv = t["payload"][: abi.FIG_TEXT_DTYPE.itemsize].view(abi.FIG_TEXT_DTYPE)[0]
            tv["first_glyph"], tv["n_glyphs"] = g0, g1 - g0
        arr["zlevel"] = L
        node_arrays.append(arr)
        root_arrays.append(np.arange(len(arr), dtype=np.int32))
    lists = (abi.FdcRenderList * layers)()
    for i, (na, ra) in enumerate(zip(node_arrays, root_arrays)):
        lists[i].nodes, lists[i].n_nodes = na.ctypes.data, len(na)
        lists[i].root_ids, lists[i].n_roots = ra.ctypes.data, len(ra)
    ops = np.zeros(1, dtype=abi.DRAW_OP_DTYPE)
    return PackedScene(node_arrays, root_arrays, glyphs, ops, lists)


def glyph_image_keys():
    return [GLYPH_KEY0 + i for i in range(N_GLYPHS)]


# ----------------------------------------------------------------------------- trace helpers
def trace_renders(renders: Renders, width: int, height: int, images=None, atlasSize: int = 2048,
                  clearColor=(1.0, 1.0, 1.0, 1.0)) -> Trace:
    setFigUiScale(1.0)
    tb = TraceBackend(atlasSize=atlasSize)
    for key, img in (images or []):
        tb.putImage(key, img)
    renderFrame(tb, renders, (float(width), float(height)), clearColor=clearColor)
    return tb.trace()


def config_trace(cfg: int, width: int = 0, height: int = 0, **kw) -> Trace:
    """BASELINE.json configs[1..4] -> recorded frame.  Default sizes are the config's own."""
    if cfg == 2:
        width, height = width or 1920, height or 1080
        return trace_renders(renderlist_100(width, height, **kw), width, height)
    if cfg == 3:
        width, height = width or 3840, height or 2160
        return trace_renders(text_page(width, height, **kw), width, height, images=text_page_images(),
                             clearColor=(0.0, 0.0, 0.0, 1.0))
    if cfg == 4:
        width, height = width or 3840, height or 2160
        return trace_renders(clip_mask_table(width, height, **kw), width, height)
    if cfg == 5:
        width, height = width or 3840, height or 2160
        return rects_and_glyphs(width, height, **kw)
    raise ValueError("cfg must be 2..5")
