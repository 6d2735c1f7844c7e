"""Front-end: host-side restatement of the `figdraw/figrender` frame entry.

The reference's front-end (src/figdraw/figrender.nim) stays Nim and is reused unchanged on top of the
new backend; it cannot be compiled in this image (no Nim), so this module restates the part of it
that decides WHICH backend calls are made, in WHAT order, with WHAT parameters:

  renderFrame   figrender.nim:1960-2002      renderRoot   :1946-1958
  render        :1756-1839 (stage order)     renderDropShadows :654-689   renderInnerShadows :716-744
  renderRoundedShapeScaledCorners :806-873   renderText :417-497 (glyph loop only)
  renderImage/renderMsdfImage/renderMtsdfImage/renderBackdropBlur :1673-1754
  renderDrawableLine :946-995, Circle :1122-1136, Rect :1138-1142, Ellipse :1617-1635,
  renderDrawableQuadraticBezierSdf :1330-1370 (3-control Beziers only)

It drives any `BackendContext` (figbackend.py): the trace recorder, the CUDA context, or a test fake.
All arithmetic the reference does in float32 is float32 here.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

from .abi import SdfMode
from .figbackend import BackendContext, BackendFill, ZeroRadii, colors4, solid, toBackendFill
from .fignodes import (DrawableKind, DrawableOp, Fig, FigFlags, FigKind, Fill, FillGradientAxis, FillKind, Rect,
                       RenderList, Renders, RenderStroke, ShadowStyle, StrokeCap, StrokeJoin, f32, rgba, rgba_tuple)

_uiScale = f32(1.0)


def setFigUiScale(scale: float) -> None:
    """common/shared.nim:67-71."""
    global _uiScale
    _uiScale = f32(scale)


def figUiScale() -> float:
    return float(_uiScale)


def _round(x) -> np.float32:
    """Nim `round`: half away from zero."""
    x = float(x)
    return f32(math.copysign(math.floor(abs(x) + 0.5), x))


def scaled(v):
    if isinstance(v, Rect):
        return v.scaled(_uiScale)
    if isinstance(v, (tuple, list)):
        return tuple(f32(c) * _uiScale for c in v)
    return f32(v) * _uiScale


# ----------------------------------------------------------------------------- fill helpers (figrender.nim:580-640)
def _lerpColor(a: int, b: int, t) -> int:
    t = min(max(f32(t), f32(0)), f32(1))
    inv = f32(1) - t
    ca, cb = rgba_tuple(a), rgba_tuple(b)
    return rgba(*[int(_round(f32(x) * inv + f32(y) * t)) for x, y in zip(ca, cb)])


def fillAlphaMax(fill: Fill) -> int:
    if fill.kind == FillKind.flColor:
        return (fill.color >> 24) & 255
    if fill.kind == FillKind.flLinear2:
        return max((fill.start >> 24) & 255, (fill.stop >> 24) & 255)
    return max((fill.start >> 24) & 255, (fill.mid >> 24) & 255, (fill.stop >> 24) & 255)


def sampleGradientColor(fill: Fill, t) -> int:
    if fill.kind == FillKind.flColor:
        return fill.color
    if fill.kind == FillKind.flLinear2:
        return _lerpColor(fill.start, fill.stop, t)
    ct = min(max(f32(t), f32(0)), f32(1))
    mid = min(max(f32(fill.midPos) / f32(255.0), f32(0.01)), f32(0.99))
    if ct <= mid:
        return _lerpColor(fill.start, fill.mid, ct / mid)
    return _lerpColor(fill.mid, fill.stop, (ct - mid) / (f32(1) - mid))


def fillCenterColor(fill: Fill) -> int:
    return sampleGradientColor(fill, 0.5)


def gradientColors(fill: Fill):
    """figrender.nim:623-647; vertex order BL, BR, TR, TL."""
    ax = fill.axis if fill.kind != FillKind.flColor else FillGradientAxis.fgaX
    ts = {
        FillGradientAxis.fgaX: (0.0, 1.0, 1.0, 0.0),
        FillGradientAxis.fgaY: (1.0, 1.0, 0.0, 0.0),
        FillGradientAxis.fgaDiagTLBR: (0.5, 1.0, 0.5, 0.0),
        FillGradientAxis.fgaDiagBLTR: (0.0, 0.5, 1.0, 0.5),
    }[ax]
    return [sampleGradientColor(fill, t) for t in ts]


# ----------------------------------------------------------------------------- corners (figrender.nim:549-571)
def _scaledCorners(x: Sequence[float], y: Sequence[float]):
    return (tuple(float(f32(v) * _uiScale) for v in x), tuple(float(f32(v) * _uiScale) for v in y))


def nodeScaledCorners(node: Fig):
    y = node.cornerRadiiY if (node.flags & FigFlags.NfEllipticalCorners) else node.corners
    return _scaledCorners(node.corners, y)


def _radiusCorner(radius) -> int:
    if radius <= 0.0:
        return 0
    if radius >= 65535.0:
        return 65535
    return int(_round(radius))


# ----------------------------------------------------------------------------- shadows
def renderDropShadows(ctx: BackendContext, node: Fig) -> None:
    for shadow in node.shadows:
        if shadow.style != ShadowStyle.DropShadow:
            continue
        if shadow.blur <= 0.0 and shadow.spread <= 0.0:
            continue
        if fillAlphaMax(shadow.fill) == 0:
            continue
        box = scaled(node.screenBox)
        sx, sy = scaled(shadow.x), scaled(shadow.y)
        blur, spread = scaled(shadow.blur), scaled(shadow.spread)
        blurPad = _round(f32(1.5) * blur)
        pad = max(_round(spread) + blurPad, f32(0))
        srx, sry, srw, srh = box.x + sx, box.y + sy, box.w + f32(0), box.h + f32(0)
        quad = (srx - pad, sry - pad, srw + f32(2) * pad, srh + f32(2) * pad)
        ctx.drawRoundedRectSdf(
            rect=tuple(float(v) for v in quad),
            fill=toBackendFill(shadow.fill),
            radii=nodeScaledCorners(node),
            mode=SdfMode.sdfModeDropShadow,
            factor=float(blur),
            spread=float(spread),
            shapeSize=(float(srw), float(srh)),
        )


def hasActiveInnerShadow(node: Fig) -> bool:
    for shadow in node.shadows:
        if shadow.style != ShadowStyle.InnerShadow:
            continue
        if shadow.blur <= 0.0 and shadow.spread <= 0.0:
            continue
        if fillAlphaMax(shadow.fill) == 0:
            continue
        return True
    return False


def renderInnerShadows(ctx: BackendContext, node: Fig) -> None:
    for shadow in node.shadows:
        if shadow.style != ShadowStyle.InnerShadow:
            continue
        if shadow.blur <= 0.0 and shadow.spread <= 0.0:
            continue
        if fillAlphaMax(shadow.fill) == 0:
            continue
        ctx.drawRoundedRectSdf(
            rect=scaled(node.screenBox).tuple(),
            fill=toBackendFill(shadow.fill),
            radii=nodeScaledCorners(node),
            mode=SdfMode.sdfModeInsetShadow,
            factor=float(scaled(shadow.blur)),
            spread=float(scaled(shadow.spread)),
            shapeSize=(float(scaled(shadow.x)), float(scaled(shadow.y))),
        )


# ----------------------------------------------------------------------------- boxes
def renderRoundedShapeScaledCorners(ctx, shapeBox: Rect, shapeFill: Fill, shapeStroke: RenderStroke, corners) -> None:
    """figrender.nim:806-873 (SDF branch)."""
    box = scaled(shapeBox).tuple()
    hasGradient = shapeFill.kind in (FillKind.flLinear2, FillKind.flLinear3) and fillAlphaMax(shapeFill) > 0
    if hasGradient:
        ctx.drawRoundedRectSdf(rect=box, fill=toBackendFill(shapeFill), radii=corners, mode=SdfMode.sdfModeClipAA,
                               factor=4.0, spread=0.0, shapeSize=(0.0, 0.0))
    elif fillAlphaMax(shapeFill) > 0:
        ctx.drawRoundedRectSdf(rect=box, fill=solid(fillCenterColor(shapeFill)), radii=corners,
                               mode=SdfMode.sdfModeClipAA, factor=4.0, spread=0.0, shapeSize=(0.0, 0.0))
    if fillAlphaMax(shapeStroke.fill) > 0 and shapeStroke.weight > 0:
        ctx.drawRoundedRectSdf(rect=box, fill=toBackendFill(shapeStroke.fill), radii=corners,
                               mode=SdfMode.sdfModeAnnularAA, factor=float(scaled(shapeStroke.weight)), spread=0.0,
                               shapeSize=(0.0, 0.0))


def renderRoundedShape(ctx, shapeBox, shapeFill, shapeStroke, cornersX, cornersY=None) -> None:
    cornersY = cornersX if cornersY is None else cornersY
    renderRoundedShapeScaledCorners(ctx, shapeBox, shapeFill, shapeStroke, _scaledCorners(cornersX, cornersY))


def renderBoxes(ctx, node: Fig) -> None:
    y = node.cornerRadiiY if (node.flags & FigFlags.NfEllipticalCorners) else node.corners
    renderRoundedShape(ctx, node.screenBox, node.fill, node.stroke, node.corners, y)


# ----------------------------------------------------------------------------- drawables
def _vlen(x, y) -> np.float32:
    return f32(math.sqrt(float(f32(x) * f32(x) + f32(y) * f32(y))))


def renderDrawableStrokeCap(ctx, center, radius, fill: Fill) -> None:
    radius = f32(radius)
    if radius <= 0.0 or fillAlphaMax(fill) == 0:
        return
    d = radius * f32(2)
    box = Rect(f32(center[0]) - radius, f32(center[1]) - radius, d, d)
    rc = _radiusCorner(radius)
    renderRoundedShape(ctx, box, fill, RenderStroke(), (rc, rc, rc, rc))


def renderDrawableLine(ctx, origin, op: DrawableOp, stroke: RenderStroke) -> None:
    """figrender.nim:946-995: a rotated zero-radius box (+ round caps)."""
    weight = max(f32(0), f32(stroke.weight))
    if weight <= 0.0 or fillAlphaMax(stroke.fill) == 0:
        return
    ax, ay = f32(origin[0]) + f32(op.a[0]), f32(origin[1]) + f32(op.a[1])
    bx, by = f32(origin[0]) + f32(op.b[0]), f32(origin[1]) + f32(op.b[1])
    dx, dy = bx - ax, by - ay
    length = _vlen(dx, dy)
    if length <= 0.0:
        return
    cap = StrokeCap.scButt if stroke.cap == StrokeCap.scAuto else stroke.cap
    capRadius = weight * f32(0.5)
    dirx, diry = dx / length, dy / length
    dax, day, dbx, dby, drawLength = ax, ay, bx, by, length
    if cap == StrokeCap.scSquare:
        dax, day = ax - dirx * capRadius, ay - diry * capRadius
        dbx, dby = bx + dirx * capRadius, by + diry * capRadius
        drawLength = length + weight
    cx, cy = (dax + dbx) / f32(2), (day + dby) / f32(2)
    box = Rect(cx - drawLength / f32(2), cy - weight / f32(2), drawLength, weight)
    sb = scaled(box)
    pivot = (sb.x + sb.w / f32(2), sb.y + sb.h / f32(2))
    angle = f32(math.atan2(float(dy), float(dx)))
    ctx.saveTransform()
    try:
        ctx.translate((float(pivot[0]), float(pivot[1])))
        ctx.rotate(float(angle))
        ctx.translate((float(-pivot[0]), float(-pivot[1])))
        renderRoundedShape(ctx, box, stroke.fill, RenderStroke(), (0, 0, 0, 0))
    finally:
        ctx.restoreTransform()
    if cap == StrokeCap.scRound:
        renderDrawableStrokeCap(ctx, (ax, ay), capRadius, stroke.fill)
        renderDrawableStrokeCap(ctx, (bx, by), capRadius, stroke.fill)


def _quadraticPoint(p0, p1, p2, t):
    t = f32(t)
    inv = f32(1) - t
    return tuple(f32(a) * (inv * inv) + f32(b) * (f32(2) * inv * t) + f32(c) * (t * t) for a, b, c in zip(p0, p1, p2))


def _quadraticBounds(p0, p1, p2, padding):
    mn = [min(f32(p0[0]), f32(p2[0])), min(f32(p0[1]), f32(p2[1]))]
    mx = [max(f32(p0[0]), f32(p2[0])), max(f32(p0[1]), f32(p2[1]))]
    for k in (0, 1):
        denom = f32(p0[k]) - f32(2) * f32(p1[k]) + f32(p2[k])
        if abs(denom) > 0.000001:
            t = (f32(p0[k]) - f32(p1[k])) / denom
            if 0.0 < t < 1.0:
                q = _quadraticPoint(p0, p1, p2, t)
                for j in (0, 1):
                    mn[j] = min(mn[j], q[j])
                    mx[j] = max(mx[j], q[j])
    padding = f32(padding)
    return Rect(mn[0] - padding, mn[1] - padding, mx[0] - mn[0] + padding * f32(2), mx[1] - mn[1] + padding * f32(2))


def renderDrawableQuadraticBezierSdf(ctx, origin, p0, p1, p2, stroke: RenderStroke, cap=StrokeCap.scAuto) -> None:
    """figrender.nim:1330-1370."""
    if cap == StrokeCap.scAuto:
        cap = StrokeCap.scRound if stroke.cap == StrokeCap.scAuto else stroke.cap
    cr = (f32(p1[0]) - f32(p0[0])) * (f32(p2[1]) - f32(p1[1])) - (f32(p1[1]) - f32(p0[1])) * (f32(p2[0]) - f32(p1[0]))
    if abs(cr) <= 0.0001:
        s2 = RenderStroke(weight=stroke.weight, fill=stroke.fill, cap=cap, join=stroke.join)
        renderDrawableLine(ctx, origin, DrawableOp(kind=DrawableKind.dkLine, a=tuple(p0), b=tuple(p2)), s2)
        return
    weight = max(f32(0), f32(stroke.weight))
    padding = weight * f32(0.5) + f32(2.0) / _uiScale
    a = (f32(origin[0]) + f32(p0[0]), f32(origin[1]) + f32(p0[1]))
    b = (f32(origin[0]) + f32(p1[0]), f32(origin[1]) + f32(p1[1]))
    c = (f32(origin[0]) + f32(p2[0]), f32(origin[1]) + f32(p2[1]))
    box = _quadraticBounds(a, b, c, padding)
    if box.w <= 0.0 or box.h <= 0.0:
        return
    cx, cy = box.x + box.w * f32(0.5), box.y + box.h * f32(0.5)
    loc = lambda p: (float((p[0] - cx) * _uiScale), float((p[1] - cy) * _uiScale))
    ctx.drawQuadraticBezierSdf(rect=scaled(box).tuple(), fill=toBackendFill(stroke.fill), p0=loc(a), p1=loc(b),
                               p2=loc(c), strokeWeight=float(scaled(weight)), cap=int(cap))


# ----------------------------------------------------------------------------- curves (figrender.nim:908-1130, :1134-1611)
# Transcendentals (cos, sin, arccos, arctan2) are evaluated in double and rounded to float32 -- the reference calls the
# float32 libm entry points, which may differ in the last ulp; the native flattener does exactly what is done here.
DrawableAdaptiveTolerancePx = f32(0.5)
MaxAdaptiveDrawableSteps = max(48 * 4, 64)   # DefaultDrawableBezierSteps = 48 (fignodes.nim:100)
MaxAdaptiveCurveDepth = 8


def _v(p):
    return (f32(p[0]), f32(p[1]))


def _add(a, b):
    return (a[0] + b[0], a[1] + b[1])


def _sub(a, b):
    return (a[0] - b[0], a[1] - b[1])


def _mul(a, k):
    return (a[0] * k, a[1] * k)


def _normalizedOr(v, fallback):
    ln = _vlen(v[0], v[1])
    if ln <= 0.000001:
        return fallback
    return (v[0] / ln, v[1] / ln)


def _normalLeft(d):
    return (-d[1], d[0])


def _cross2(a, b):
    return a[0] * b[1] - a[1] * b[0]


def _withCap(stroke: RenderStroke, cap) -> RenderStroke:
    return RenderStroke(weight=stroke.weight, fill=stroke.fill, cap=cap, join=stroke.join)


def resolveCurveCap(stroke: RenderStroke):
    return StrokeCap.scRound if stroke.cap == StrokeCap.scAuto else stroke.cap


def resolveCurveJoin(stroke: RenderStroke):
    return StrokeJoin.sjRound if stroke.join == StrokeJoin.sjAuto else stroke.join


def renderDrawableEndpointCap(ctx, origin, point, tangent, radius, stroke: RenderStroke, cap, isStart: bool) -> None:
    """figrender.nim:1010-1039."""
    if radius <= 0.0 or fillAlphaMax(stroke.fill) == 0:
        return
    if cap == StrokeCap.scRound:
        renderDrawableStrokeCap(ctx, _add(origin, point), radius, stroke.fill)
    elif cap == StrokeCap.scSquare:
        d = _normalizedOr(tangent, (f32(1), f32(0)))
        a = _sub(point, _mul(d, radius)) if isStart else point
        b = point if isStart else _add(point, _mul(d, radius))
        renderDrawableLine(ctx, origin, DrawableOp(kind=DrawableKind.dkLine, a=a, b=b), _withCap(stroke, StrokeCap.scButt))


def _lineIntersection(p, r, q, s):
    denom = _cross2(r, s)
    if abs(denom) <= 0.000001:
        return None
    t = _cross2(_sub(q, p), s) / denom
    return _add(p, _mul(r, t))


def renderDrawableFilledQuad(ctx, verts, fill: Fill) -> None:
    """figrender.nim:1049-1057."""
    if fillAlphaMax(fill) == 0:
        return
    color = fillCenterColor(fill)
    ctx.drawFilledQuad([tuple(float(c * _uiScale) for c in v) for v in verts], [color] * 4)


def renderDrawableStrokeJoin(ctx, origin, point, incomingTangent, outgoingTangent, radius, fill: Fill, join) -> None:
    """figrender.nim:1059-1109."""
    if radius <= 0.0 or fillAlphaMax(fill) == 0:
        return
    if join == StrokeJoin.sjRound:
        renderDrawableStrokeCap(ctx, _add(origin, point), radius, fill)
    elif join in (StrokeJoin.sjBevel, StrokeJoin.sjMiter):
        incoming = _normalizedOr(incomingTangent, (f32(1), f32(0)))
        outgoing = _normalizedOr(outgoingTangent, incoming)
        turn = _cross2(incoming, outgoing)
        if abs(turn) <= 0.0001:
            return
        side = f32(-1.0) if turn > 0.0 else f32(1.0)
        incomingOuter = _add(point, _mul(_normalLeft(incoming), radius * side))
        outgoingOuter = _add(point, _mul(_normalLeft(outgoing), radius * side))
        if join == StrokeJoin.sjMiter:
            mp = _lineIntersection(incomingOuter, incoming, outgoingOuter, outgoing)
            if mp is not None:
                dm = _sub(mp, point)
                if _vlen(dm[0], dm[1]) <= radius * f32(4.0):
                    renderDrawableFilledQuad(ctx, [_add(origin, point), _add(origin, incomingOuter), _add(origin, mp),
                                                   _add(origin, outgoingOuter)], fill)
                    return
        renderDrawableFilledQuad(ctx, [_add(origin, point), _add(origin, incomingOuter), _add(origin, outgoingOuter),
                                       _add(origin, outgoingOuter)], fill)


def bezierPoint(controls, t):
    """De Casteljau, figrender.nim:1134-1147."""
    if not controls:
        return (f32(0), f32(0))
    t = f32(t)
    work = [_v(p) for p in controls]
    count = len(work)
    while count > 1:
        for i in range(count - 1):
            work[i] = _add(_mul(work[i], f32(1) - t), _mul(work[i + 1], t))
        count -= 1
    return work[0]


def explicitDrawableStepCount(steps: int, nodeSteps: int) -> int:
    if steps != 0:
        return max(1, int(steps))
    if nodeSteps != 0:
        return max(1, int(nodeSteps))
    return 0


def _spanStartTangent(span):
    p0, p1, p2 = span
    return _normalizedOr(_sub(p1, p0), _normalizedOr(_sub(p2, p0), (f32(1), f32(0))))


def _spanEndTangent(span):
    p0, p1, p2 = span
    return _normalizedOr(_sub(p2, p1), _normalizedOr(_sub(p2, p0), (f32(1), f32(0))))


def _pointDistancePx(a, b):
    d = _mul(_sub(a, b), _uiScale)
    return _vlen(d[0], d[1])


def _distanceToLine(p, a, b):
    ab = _sub(b, a)
    denom = ab[0] * ab[0] + ab[1] * ab[1]
    if denom <= 0.000001:
        d = _sub(p, a)
        return _vlen(d[0], d[1])
    pa = _sub(p, a)
    h = min(max((pa[0] * ab[0] + pa[1] * ab[1]) / denom, f32(0)), f32(1))
    d = _sub(p, _add(a, _mul(ab, h)))
    return _vlen(d[0], d[1])


def bezierQuadraticSpan(controls, t0, t2):
    t0, t2 = f32(t0), f32(t2)
    tm = (t0 + t2) * f32(0.5)
    p0, pm, p2 = bezierPoint(controls, t0), bezierPoint(controls, tm), bezierPoint(controls, t2)
    p1 = _sub(_mul(pm, f32(2)), _mul(_add(p0, p2), f32(0.5)))
    return (p0, p1, p2)


def _quadraticApproxErrorPx(controls, span, t0, t2):
    result = f32(0)
    for localT in (f32(0.25), f32(0.75)):
        t = t0 + (t2 - t0) * localT
        actual = bezierPoint(controls, t)
        approx = _quadraticPoint(span[0], span[1], span[2], localT)
        result = max(result, _pointDistancePx(actual, approx))
    return result


def _appendAdaptiveBezierSpan(controls, t0, t2, depth, spans):
    span = bezierQuadraticSpan(controls, t0, t2)
    error = _quadraticApproxErrorPx(controls, span, f32(t0), f32(t2))
    if error <= DrawableAdaptiveTolerancePx or depth >= MaxAdaptiveCurveDepth or len(spans) >= MaxAdaptiveDrawableSteps - 1:
        spans.append(span)
    else:
        tm = (f32(t0) + f32(t2)) * f32(0.5)
        _appendAdaptiveBezierSpan(controls, t0, tm, depth + 1, spans)
        _appendAdaptiveBezierSpan(controls, tm, t2, depth + 1, spans)


def _bezierSpans(controls, fixedSteps):
    if fixedSteps > 0:
        return [bezierQuadraticSpan(controls, f32(step) / f32(fixedSteps), f32(step + 1) / f32(fixedSteps))
                for step in range(fixedSteps)]
    spans = []
    _appendAdaptiveBezierSpan(controls, f32(0), f32(1), 0, spans)
    return spans


def _renderQuadraticSpans(ctx, origin, spans, stroke: RenderStroke) -> None:
    """Shared body of renderDrawableBezierQuadratics (:1414-1457) and renderDrawableArcQuadratics (:1551-1593)."""
    cap, join = resolveCurveCap(stroke), resolveCurveJoin(stroke)
    simpleRoundSpans = cap == StrokeCap.scRound and join == StrokeJoin.sjRound
    spanCap = StrokeCap.scRound if simpleRoundSpans else StrokeCap.scButt
    capRadius = max(f32(0), f32(stroke.weight)) / f32(2)
    previous = None
    for step, span in enumerate(spans):
        renderDrawableQuadraticBezierSdf(ctx, origin, span[0], span[1], span[2], stroke, spanCap)
        if not simpleRoundSpans:
            if step == 0:
                renderDrawableEndpointCap(ctx, origin, span[0], _spanStartTangent(span), capRadius, stroke, cap, True)
            else:
                renderDrawableStrokeJoin(ctx, origin, span[0], _spanEndTangent(previous), _spanStartTangent(span), capRadius,
                                         stroke.fill, join)
            if step == len(spans) - 1:
                renderDrawableEndpointCap(ctx, origin, span[2], _spanEndTangent(span), capRadius, stroke, cap, False)
        previous = span


def _appendAdaptiveBezierSegmentPoint(controls, t0, t2, depth, points):
    p0, p2 = bezierPoint(controls, t0), bezierPoint(controls, t2)
    tm = (f32(t0) + f32(t2)) * f32(0.5)
    pm = bezierPoint(controls, tm)
    error = _distanceToLine(_mul(pm, _uiScale), _mul(p0, _uiScale), _mul(p2, _uiScale))
    if error <= DrawableAdaptiveTolerancePx or depth >= MaxAdaptiveCurveDepth or len(points) >= MaxAdaptiveDrawableSteps:
        points.append(p2)
    else:
        _appendAdaptiveBezierSegmentPoint(controls, t0, tm, depth + 1, points)
        _appendAdaptiveBezierSegmentPoint(controls, tm, t2, depth + 1, points)


def renderDrawableBezierSegments(ctx, origin, op: DrawableOp, stroke: RenderStroke, nodeSteps: int) -> None:
    """figrender.nim:1368-1412: polyline with endpoint caps and joins (what a 2-control Bezier takes)."""
    fixedSteps = explicitDrawableStepCount(op.steps, nodeSteps)
    points = [bezierPoint(op.controls, f32(0))]
    if fixedSteps > 0:
        for step in range(1, fixedSteps + 1):
            points.append(bezierPoint(op.controls, f32(step) / f32(fixedSteps)))
    else:
        _appendAdaptiveBezierSegmentPoint(op.controls, f32(0), f32(1), 0, points)
    if len(points) < 2:
        return
    cap, join = resolveCurveCap(stroke), resolveCurveJoin(stroke)
    capRadius = max(f32(0), f32(stroke.weight)) / f32(2)
    segmentStroke = _withCap(stroke, StrokeCap.scButt)
    previous, previousTangent = points[0], (f32(1), f32(0))
    for step in range(1, len(points)):
        current = points[step]
        tangent = _sub(current, previous)
        renderDrawableLine(ctx, origin, DrawableOp(kind=DrawableKind.dkLine, a=previous, b=current), segmentStroke)
        if step == 1:
            renderDrawableEndpointCap(ctx, origin, previous, tangent, capRadius, stroke, cap, True)
        else:
            renderDrawableStrokeJoin(ctx, origin, previous, previousTangent, tangent, capRadius, stroke.fill, join)
        if step == len(points) - 1:
            renderDrawableEndpointCap(ctx, origin, current, tangent, capRadius, stroke, cap, False)
        previous, previousTangent = current, tangent


def renderDrawableBezier(ctx, origin, op: DrawableOp, stroke: RenderStroke, nodeSteps: int) -> None:
    """figrender.nim:1459-1486 (SDF build: 3 controls -> one quadratic SDF, more -> quadratic spans, 2 -> segments)."""
    if len(op.controls) < 2:
        return
    if stroke.weight <= 0.0 or fillAlphaMax(stroke.fill) == 0:
        return
    if len(op.controls) == 3:
        renderDrawableQuadraticBezierSdf(ctx, origin, op.controls[0], op.controls[1], op.controls[2], stroke,
                                         resolveCurveCap(stroke))
    elif len(op.controls) > 3:
        _renderQuadraticSpans(ctx, origin, _bezierSpans(op.controls, explicitDrawableStepCount(op.steps, nodeSteps)), stroke)
    else:
        renderDrawableBezierSegments(ctx, origin, op, stroke, nodeSteps)


def _arcPoint(center, radius, angle):
    a = float(angle)
    return (f32(center[0]) + f32(math.cos(a)) * radius, f32(center[1]) + f32(math.sin(a)) * radius)


def adaptiveArcStepCount(radius, sweepAngle) -> int:
    """figrender.nim:1307-1318."""
    radiusPx = max(f32(0), f32(radius) * _uiScale)
    absSweep = abs(f32(sweepAngle))
    if radiusPx <= 0.0 or absSweep <= 0.0:
        return 1
    cosLimit = min(max(f32(1) - DrawableAdaptiveTolerancePx / radiusPx, f32(-1)), f32(1))
    maxAngle = max(f32(0.01), f32(2) * f32(math.acos(float(cosLimit))))
    return min(max(int(math.ceil(float(absSweep / maxAngle))), 1), MaxAdaptiveDrawableSteps)


def renderDrawableArc(ctx, origin, op: DrawableOp, stroke: RenderStroke, nodeSteps: int) -> None:
    """figrender.nim:1595-1611 -> renderDrawableArcQuadratics :1551-1593, arcQuadraticSpan :1535-1549."""
    radius = max(f32(0), f32(op.radius))
    if radius <= 0.0 or f32(op.sweepAngle) == 0.0:
        return
    if stroke.weight <= 0.0 or fillAlphaMax(stroke.fill) == 0:
        return
    steps = explicitDrawableStepCount(op.steps, nodeSteps) or adaptiveArcStepCount(op.radius, op.sweepAngle)
    start, sweep = f32(op.startAngle), f32(op.sweepAngle)
    spans = []
    for step in range(steps):
        t0, t2 = f32(step) / f32(steps), f32(step + 1) / f32(steps)
        tm = (t0 + t2) * f32(0.5)
        p0 = _arcPoint(op.center, radius, start + sweep * t0)
        pm = _arcPoint(op.center, radius, start + sweep * tm)
        p2 = _arcPoint(op.center, radius, start + sweep * t2)
        spans.append((p0, _sub(_mul(pm, f32(2)), _mul(_add(p0, p2), f32(0.5))), p2))
    _renderQuadraticSpans(ctx, origin, spans, stroke)


def renderDrawableOps(ctx, node: Fig) -> None:
    origin = (node.screenBox.x, node.screenBox.y)
    fill, stroke = node.fill, node.drawStroke
    for op in node.drawOps:
        if op.kind == DrawableKind.dkLine:
            renderDrawableLine(ctx, origin, op, stroke)
        elif op.kind == DrawableKind.dkCircle:
            radius = max(f32(0), f32(op.radius))
            if radius <= 0.0:
                continue
            d = radius * f32(2)
            box = Rect(origin[0] + f32(op.center[0]) - radius, origin[1] + f32(op.center[1]) - radius, d, d)
            rc = _radiusCorner(radius)
            renderRoundedShape(ctx, box, fill, stroke, (rc, rc, rc, rc))
        elif op.kind == DrawableKind.dkRectangle:
            box = Rect(origin[0] + op.box.x, origin[1] + op.box.y, op.box.w, op.box.h)
            renderRoundedShape(ctx, box, fill, stroke, op.corners)
        elif op.kind == DrawableKind.dkEllipse:
            rx, ry = max(f32(0), f32(op.ellipseRadii[0])), max(f32(0), f32(op.ellipseRadii[1]))
            if rx <= 0.0 or ry <= 0.0:
                continue
            box = Rect(origin[0] + f32(op.center[0]) - rx, origin[1] + f32(op.center[1]) - ry, rx * f32(2), ry * f32(2))
            renderRoundedShape(ctx, box, fill, stroke, (rx,) * 4, (ry,) * 4)
        elif op.kind == DrawableKind.dkBezier:
            renderDrawableBezier(ctx, origin, op, stroke, node.drawSteps)
        elif op.kind == DrawableKind.dkArc:
            renderDrawableArc(ctx, origin, op, stroke, node.drawSteps)
        else:
            raise NotImplementedError(f"drawable op {op.kind!r}")


def renderDrawable(ctx, node: Fig) -> None:
    """figrender.nim:1653-1667."""
    if node.drawAa <= 0.0:
        renderDrawableOps(ctx, node)
        return
    oldAa = ctx.sdfAaFactor()
    if oldAa == node.drawAa:
        renderDrawableOps(ctx, node)
        return
    ctx.setSdfAaFactor(node.drawAa)
    try:
        renderDrawableOps(ctx, node)
    finally:
        ctx.setSdfAaFactor(oldAa)


# ----------------------------------------------------------------------------- text / images / blur
def renderText(ctx, node: Fig) -> None:
    """figrender.nim:417-497: selection rects, decorations, then one atlas quad per glyph with 4 vertex colours from the
    span fill.  The layout itself (pixie arrangement) is upstream; per-glyph subpixel VARIANTS select a different atlas
    key and are therefore the caller's choice of `Glyph.key`."""
    subpixel = ctx.textSubpixelPositioningEnabled()
    ctx.saveTransform()
    ctx.translate((float(scaled(node.screenBox.x)), float(scaled(node.screenBox.y))))
    if node.flags & FigFlags.NfInvertY:
        ctx.translate((0.0, float(scaled(node.screenBox.h))))
        ctx.scale((1.0, -1.0))
    if (node.flags & FigFlags.NfSelectText) and fillAlphaMax(node.fill) > 0:
        for sel in node.selectionRects:
            if sel.h > 0:
                r = Rect(sel.x, sel.y, max(sel.w, f32(1.0)), sel.h)
                ctx.drawRoundedRectSdf(rect=scaled(r).tuple(), fill=toBackendFill(node.fill), radii=ZeroRadii,
                                       mode=SdfMode.sdfModeClipAA, factor=4.0, spread=0.0, shapeSize=(0.0, 0.0))
    for deco, color in node.decorations:  # drawTextDecoration :355-368
        if deco.w <= 0 or deco.h <= 0:
            continue
        ctx.drawRoundedRectSdf(rect=scaled(deco).tuple(), fill=toBackendFill(color), radii=ZeroRadii,
                               mode=SdfMode.sdfModeClipAA, factor=4.0, spread=0.0, shapeSize=(0.0, 0.0))
    for glyph in node.glyphs:
        gx, shift = f32(glyph.pos[0]), f32(0.0)
        if subpixel:
            snapped = f32(math.floor(float(gx)))
            shift = max(f32(0.0), min(gx - snapped, f32(0.999)))
            gx = snapped
        ctx.setTextSubpixelShift(float(shift))
        if not ctx.hasImage(glyph.key):
            ctx.setTextSubpixelShift(0.0)
            continue
        ctx.drawImage(glyph.key, (float(gx), float(glyph.pos[1])), gradientColors(glyph.fill), (0.0, 0.0), False)
        if subpixel:
            ctx.setTextSubpixelShift(0.0)
    ctx.setTextSubpixelShift(0.0)
    ctx.restoreTransform()


def renderImage(ctx, node: Fig) -> None:
    if node.image.id == 0:
        return
    box = scaled(node.screenBox)
    c = fillCenterColor(node.image.fill)
    ctx.drawImage(node.image.id, (float(box.x), float(box.y)), [c, c, c, c], (float(box.w), float(box.h)),
                  bool(node.flags & FigFlags.NfInvertY))


def _renderSdfImage(ctx, node: Fig, style, mtsdf: bool) -> None:
    if style.id == 0:
        return
    box = scaled(node.screenBox)
    pxRange = style.pxRange if style.pxRange > 0.0 else 4.0
    thr = style.sdThreshold if 0.0 < style.sdThreshold < 1.0 else 0.5
    sw = float(scaled(max(f32(0), f32(style.strokeWeight))))
    fn = ctx.drawMtsdfImage if mtsdf else ctx.drawMsdfImage
    fn(style.id, (float(box.x), float(box.y)), fillCenterColor(style.fill), (float(box.w), float(box.h)), pxRange, thr,
       sw, bool(node.flags & FigFlags.NfInvertY))


def renderBackdropBlur(ctx, node: Fig) -> None:
    """figrender.nim:1734-1754."""
    if node.backdropBlur.blur > 0.0:
        ctx.drawBackdropBlur(scaled(node.screenBox).tuple(), nodeScaledCorners(node),
                             float(scaled(node.backdropBlur.blur)))
    if fillAlphaMax(node.fill) == 0:
        return
    overlay = Fig(kind=FigKind.nkRectangle, screenBox=node.screenBox, fill=node.fill, corners=node.corners,
                  cornerRadiiY=node.cornerRadiiY)
    if node.flags & FigFlags.NfEllipticalCorners:
        overlay.flags |= FigFlags.NfEllipticalCorners
    overlay.stroke = RenderStroke(weight=0.0, fill=rgba(0, 0, 0, 0))
    renderBoxes(ctx, overlay)


# ----------------------------------------------------------------------------- the DFS
def render(ctx: BackendContext, nodes: RenderList, idx: int) -> None:
    """figrender.nim:1756-1839: paint-order contract.  `finally` stages run in reverse at the end."""
    node = nodes.nodes[idx]
    if node.flags & FigFlags.NfDisableRender:
        return
    box = scaled(node.screenBox)
    cleanups = []

    if node.rotation != 0:
        ctx.saveTransform()
        c = (float(box.x + box.w / f32(2)), float(box.y + box.h / f32(2)))
        ctx.translate(c)
        ctx.rotate(float(f32(node.rotation) / f32(180) * f32(math.pi)))
        ctx.translate((-c[0], -c[1]))
        cleanups.append(ctx.restoreTransform)

    if node.kind == FigKind.nkTransform:
        ctx.saveTransform()
        tr = node.transform.translation
        if tr[0] != 0.0 or tr[1] != 0.0:
            ctx.translate(tuple(float(v) for v in scaled(tr)))
        if node.transform.useMatrix:
            ctx.applyTransform(node.transform.matrix)
        cleanups.append(ctx.restoreTransform)

    if node.kind == FigKind.nkRectangle:
        renderDropShadows(ctx, node)

    if node.flags & FigFlags.NfClipContent:
        ctx.beginMask(box.tuple(), nodeScaledCorners(node))
        ctx.endMask()
        cleanups.append(ctx.popMask)

    if node.flags & FigFlags.NfRectMaskContent:
        ctx.beginRectMask(box.tuple(), nodeScaledCorners(node))
        cleanups.append(ctx.popRectMask)

    if node.kind == FigKind.nkText:
        renderText(ctx, node)
    elif node.kind == FigKind.nkDrawable:
        renderDrawable(ctx, node)
    elif node.kind == FigKind.nkRectangle:
        renderBoxes(ctx, node)
    elif node.kind == FigKind.nkImage:
        renderImage(ctx, node)
    elif node.kind == FigKind.nkMsdfImage:
        _renderSdfImage(ctx, node, node.msdfImage, False)
    elif node.kind == FigKind.nkMtsdfImage:
        _renderSdfImage(ctx, node, node.mtsdfImage, True)
    elif node.kind == FigKind.nkBackdropBlur:
        renderBackdropBlur(ctx, node)

    if node.kind == FigKind.nkRectangle and hasActiveInnerShadow(node):
        renderInnerShadows(ctx, node)

    for child in nodes.childIndex(idx):
        render(ctx, nodes, child)

    for fn in reversed(cleanups):
        fn()


def renderRoot(ctx: BackendContext, renders: Renders) -> None:
    """figrender.nim:1946-1958: layers in table order (NOT sorted), roots in rootIds order."""
    for _zlvl, lst in renders.pairs():
        for root in lst.rootIds:
            render(ctx, lst, root)


def renderFrame(ctx: BackendContext, renders: Renders, frameSize, clearMain=True,
                clearColor=(1.0, 1.0, 1.0, 1.0)) -> None:
    """figrender.nim:1960-2002."""
    fs = scaled(tuple(frameSize))
    ctx.beginFrame((float(fs[0]), float(fs[1])), clearMain=clearMain, clearMainColor=clearColor)
    ctx.saveTransform()
    ctx.scale(ctx.pixelScale())
    renderRoot(ctx, renders)
    ctx.restoreTransform()
    ctx.endFrame()
