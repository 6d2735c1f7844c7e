"""Seeded random backend-call streams that exercise every op and mode of the path (parity fuzzing).

Unlike scenes_synth.py these do not mimic an application: they drive the `BackendContext` interface directly with
random transforms (incl. rotation and mirroring), nested clip masks, rect masks, elliptical corners, every SdfMode,
Beziers, filled quads, scaled / minified / flipped images, MSDF/MTSDF strokes, AA-factor changes and backdrop blurs."""
from __future__ import annotations

import numpy as np

from .abi import SdfMode
from .figbackend import BackendFill, Trace, TraceBackend, colors4, solid
from .scenes_synth import Rng, make_glyph_bitmaps, make_star_field


def _col(rng, n=1, alpha=None):
    v = rng.integers(4 * n, 0, 256).reshape(n, 4)
    if alpha is not None:
        v[:, 3] = alpha
    return [int(r) | (int(g) << 8) | (int(b) << 16) | (int(a) << 24) for r, g, b, a in v]


def random_trace(seed: int, width: int = 512, height: int = 384, n_ops: int = 160, blur: bool = True) -> Trace:
    rng = Rng(seed * 2654435761 % (1 << 31) + 17)
    u = lambda lo=0.0, hi=1.0: float(rng.uniform(1, lo, hi)[0])
    ri = lambda lo, hi: int(rng.integers(1, lo, hi)[0])
    tb = TraceBackend(atlasSize=512)
    glyphs = make_glyph_bitmaps()[:12]
    for i, g in enumerate(glyphs):
        tb.putImage(100 + i, g)
    tb.putImage(200, make_star_field(32, 4.0, mtsdf=False))
    tb.putImage(201, make_star_field(32, 4.0, mtsdf=True))
    photo = (rng.integers(96 * 64 * 4, 0, 256).reshape(64, 96, 4)).astype(np.uint8)
    tb.putImage(300, photo)
    tb.beginFrame((width, height), clearMain=True, clearMainColor=(u(), u(), u(), 1.0))
    tb.saveTransform()
    tb.scale(1.0)

    def fill():
        k = ri(0, 4)
        if k == 0:
            return solid(_col(rng, 1, alpha=ri(0, 2) * 100 + 155)[0])
        if k == 1:
            c = _col(rng, 2)
            return BackendFill(kind=2, axis=ri(0, 4), c=(c[0], c[1], 0, 0))
        if k == 2:
            c = _col(rng, 3)
            return BackendFill(kind=3, axis=ri(0, 4), c=(c[0], c[1], c[2], 0), midPos=min(max(u(), 0.01), 0.99))
        return colors4(_col(rng, 4))

    def radii(w, h):
        if ri(0, 3) == 0:
            rx = [u(0, w * 0.6) for _ in range(4)]
            ry = [u(0, h * 0.6) for _ in range(4)]
            return (tuple(rx), tuple(ry))
        r = tuple(u(-2, min(w, h) * 0.7) for _ in range(4))
        return (r, r)

    def rect():
        w, h = u(4, 180), u(4, 140)
        return (u(-40, width - 20), u(-30, height - 20), w, h)

    depth_masks, depth_rm, depth_xf = 0, 0, 0
    for _ in range(n_ops):
        op = ri(0, 24)
        r = rect()
        if op < 7:
            modes = [SdfMode.sdfModeClipAA, SdfMode.sdfModeAnnularAA, SdfMode.sdfModeDropShadow, SdfMode.sdfModeInsetShadow,
                     SdfMode.sdfModeAnnular, SdfMode.sdfModeDropShadowAA, SdfMode.sdfModeClipAA]
            m = modes[op]
            ss = (r[2] * u(0.5, 1.0), r[3] * u(0.5, 1.0)) if m in (SdfMode.sdfModeDropShadow, SdfMode.sdfModeDropShadowAA) else (0.0, 0.0)
            if m == SdfMode.sdfModeInsetShadow:
                ss = (u(-8, 8), u(-8, 8))
            tb.drawRoundedRectSdf(r, fill(), radii(r[2], r[3]), mode=m, factor=u(0.5, 14), spread=u(0, 8), shapeSize=ss)
        elif op == 7:
            k = 100 + ri(0, 12)
            size = (0.0, 0.0) if ri(0, 2) else (u(3, 60), u(5, 90))
            tb.drawImage(k, (r[0], r[1]), _col(rng, 4), size, bool(ri(0, 2)))
        elif op == 8:
            tb.drawImage(300, (r[0], r[1]), _col(rng, 1, alpha=255) * 4, (u(10, 220), u(8, 160)), bool(ri(0, 2)))
        elif op == 9:
            fn = tb.drawMtsdfImage if ri(0, 2) else tb.drawMsdfImage
            fn(201 if fn == tb.drawMtsdfImage else 200, (r[0], r[1]), _col(rng, 1)[0], (u(12, 200), u(12, 200)), 4.0,
               u(0.35, 0.65), u(0, 5) if ri(0, 2) else 0.0, bool(ri(0, 2)))
        elif op == 10:
            hw, hh = r[2] / 2, r[3] / 2
            tb.drawQuadraticBezierSdf(r, fill(), (u(-hw, hw), u(-hh, hh)), (u(-hw, hw), u(-hh, hh)), (u(-hw, hw), u(-hh, hh)),
                                      u(1, 12), ri(0, 4))
        elif op == 11:
            c = (r[0] + r[2] / 2, r[1] + r[3] / 2)
            verts = [(c[0] + u(-60, 60), c[1] + u(-60, 60)) for _ in range(4)]
            tb.drawFilledQuad(verts, _col(rng, 4))
        elif op == 12:
            tb.drawRect(r, _col(rng, 1)[0])
        elif op == 13 and depth_masks < 3:
            tb.beginMask(r, radii(r[2], r[3]))
            if ri(0, 4) == 0:  # a second shape drawn into the same mask level
                r2 = rect()
                tb.drawRoundedRectSdf(r2, solid(_col(rng, 1, alpha=255)[0]), radii(r2[2], r2[3]))
            tb.endMask()
            depth_masks += 1
        elif op == 14 and depth_masks > 0 and depth_rm == 0:
            tb.popMask()
            depth_masks -= 1
        elif op == 15 and depth_rm == 0 and depth_masks == 0:
            tb.beginRectMask(r, radii(r[2], r[3]))
            depth_rm = 1
        elif op == 16 and depth_rm == 1 and depth_masks == 0:
            tb.popRectMask()
            depth_rm = 0
        elif op == 17 and depth_xf < 4:
            tb.saveTransform()
            tb.translate((u(-30, 30), u(-30, 30)))
            k = ri(0, 4)
            if k == 0:
                tb.translate((width / 2, height / 2))
                tb.rotate(u(-3.2, 3.2))
                tb.translate((-width / 2, -height / 2))
            elif k == 1:
                tb.scale((u(0.5, 1.6), u(0.5, 1.6)))
            elif k == 2:
                tb.translate((0.0, height * 0.8))
                tb.scale((1.0, -1.0))
            depth_xf += 1
        elif op == 18 and depth_xf > 0:
            tb.restoreTransform()
            depth_xf -= 1
        elif op == 19:
            tb.setSdfAaFactor(u(0.4, 3.0) if ri(0, 2) else 1.2)
        elif op == 20 and blur and depth_masks == 0:
            tb.drawBackdropBlur(r, radii(r[2], r[3]), u(0.2, 70))
        elif op == 21:
            tb.setTextSubpixelPositioningEnabled(bool(ri(0, 2)))
            tb.setTextSubpixelShift(u(0, 1.2))
        else:
            tb.drawRoundedRectSdf(r, fill(), radii(r[2], r[3]))
    while depth_masks > 0:
        tb.popMask()
        depth_masks -= 1
    if depth_rm:
        tb.popRectMask()
    while depth_xf > 0:
        tb.restoreTransform()
        depth_xf -= 1
    tb.restoreTransform()
    tb.endFrame()
    return tb.trace()
