"""Host logic: the front-end emits the reference's backend-call sequence (cf. tests/ttransform.nim RecordingBackend)."""
import numpy as np
import pytest

from figdraw_b200 import scenes
from figdraw_b200.abi import Op, SdfMode
from figdraw_b200.figbackend import BackendContext, TraceBackend, toBackendFill
from figdraw_b200.fignodes import (Fig, FigFlags, FigKind, FillGradientAxis, RenderList, RenderShadow, RenderStroke,
                                   Renders, ShadowStyle, linear, rect, rgba)
from figdraw_b200.figrender import renderFrame, renderRoot, setFigUiScale


def ops(trace):
    return [Op(int(o)) for o in trace.calls["op"]]


def test_rgb_boxes_call_sequence():
    tr = scenes.golden_trace("rgb_boxes_sdf")
    draws = tr.calls[tr.calls["op"] == Op.ROUNDED_RECT]
    modes = [SdfMode(int(u[0])) for u in draws["u"]]
    # root fill; red fill + stroke; drop shadow + gradient fill; blue fill + two inset shadows
    assert modes == [SdfMode.sdfModeClipAA, SdfMode.sdfModeClipAA, SdfMode.sdfModeAnnularAA, SdfMode.sdfModeDropShadow,
                     SdfMode.sdfModeClipAA, SdfMode.sdfModeClipAA, SdfMode.sdfModeInsetShadow,
                     SdfMode.sdfModeInsetShadow]
    sh = draws[3]
    # pad = round(spread) + round(1.5*blur) = 10 + 15 (figrender.nim:668-676); shapeSize = box size
    assert list(sh["f"][0:4]) == [320 + 10 - 25, 120 + 10 - 25, 220 + 50, 140 + 50]
    assert list(sh["f"][14:16]) == [220, 140] and sh["f"][12] == 10 and sh["f"][13] == 10
    ins = draws[6]
    assert list(ins["f"][14:16]) == [-6, -6]  # inset mode carries the offset in shapeSize (figrender.nim:731-744)
    assert ops(tr)[:2] == [Op.SAVE_TRANSFORM, Op.SCALE] and ops(tr)[-1] == Op.RESTORE_TRANSFORM


def test_clip_node_order_and_layers_not_sorted():
    tr = scenes.golden_trace("layers_clip")
    o = ops(tr)
    i = o.index(Op.BEGIN_MASK)
    # beginMask; endMask; the node's own fill INSIDE its mask; child; popMask (figrender.nim:1795-1820)
    assert o[i:i + 5] == [Op.BEGIN_MASK, Op.END_MASK, Op.ROUNDED_RECT, Op.ROUNDED_RECT, Op.POP_MASK]
    # renderRoot iterates the table in insertion order and never sorts (figrender.nim:1951)
    r = Renders()
    for z in (5, -3, 1):
        lst = RenderList()
        lst.addRoot(Fig(kind=FigKind.nkRectangle, zlevel=z, screenBox=rect(0, 0, 10 + z, 10), fill=rgba(1, 2, 3, 255)))
        r.setLayer(z, lst)
    tb = TraceBackend()
    tb.beginFrame((64, 64))
    renderRoot(tb, r)
    assert [float(c["f"][2]) for c in tb.trace().calls] == [15.0, 7.0, 11.0]


def test_rect_mask_and_rotation_emit_reference_calls():
    tr = scenes.trace_scene(scenes.layers_rect_mask, 800, 375)
    o = ops(tr)
    assert Op.BEGIN_RECT_MASK in o and Op.POP_RECT_MASK in o and Op.BEGIN_MASK not in o
    tr = scenes.golden_trace("line_rect")
    o = ops(tr)
    i = o.index(Op.ROTATE)
    # save; translate(pivot); rotate(atan2); translate(-pivot); box; restore  (figrender.nim:982-990)
    assert o[i - 2:i + 4] == [Op.SAVE_TRANSFORM, Op.TRANSLATE, Op.ROTATE, Op.TRANSLATE, Op.ROUNDED_RECT,
                             Op.RESTORE_TRANSFORM]
    ang = float(tr.calls[i]["f"][0])
    assert abs(ang - np.arctan2(350.0, 620.0)) < 1e-6


def test_disabled_and_transparent_nodes_are_skipped():
    lst = RenderList()
    root = lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(0, 0, 8, 8), fill=rgba(0, 0, 0, 0)))
    lst.addChild(root, Fig(kind=FigKind.nkRectangle, screenBox=rect(0, 0, 8, 8), fill=rgba(9, 9, 9, 255),
                           flags=FigFlags.NfDisableRender))
    lst.addChild(root, Fig(kind=FigKind.nkRectangle, screenBox=rect(0, 0, 8, 8), fill=rgba(9, 9, 9, 255),
                           shadows=[RenderShadow(style=ShadowStyle.DropShadow, blur=0, spread=0, fill=rgba(0, 0, 0, 255)),
                                    RenderShadow(style=ShadowStyle.DropShadow, blur=4, spread=0, fill=rgba(0, 0, 0, 0))]))
    r = Renders()
    r.setLayer(0, lst)
    tb = TraceBackend()
    tb.beginFrame((8, 8))
    renderRoot(tb, r)
    assert ops(tb.trace()) == [Op.ROUNDED_RECT]  # only the third node's fill


def test_ui_scale_multiplies_every_coordinate():
    def build(w, h):
        lst = RenderList()
        lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(10, 20, 30, 40), corners=(4, 4, 4, 4),
                        fill=rgba(1, 2, 3, 255), stroke=RenderStroke(weight=2.0, fill=rgba(0, 0, 0, 255))))
        r = Renders()
        r.setLayer(0, lst)
        return r

    tr = scenes.trace_scene(build, 100, 100, uiScale=2.0)
    setFigUiScale(1.0)
    assert (tr.width, tr.height) == (200, 200)
    d = tr.calls[tr.calls["op"] == Op.ROUNDED_RECT]
    assert list(d[0]["f"][0:8]) == [20, 40, 60, 80, 8, 8, 8, 8]
    assert d[1]["f"][12] == 4.0  # stroke weight scaled


def test_backend_fill_conversion():
    f = toBackendFill(linear(rgba(1, 2, 3, 4), rgba(5, 6, 7, 8), rgba(9, 9, 9, 9), axis=FillGradientAxis.fgaY, midPos=0))
    assert f.kind == 3 and f.axis == 1 and abs(f.midPos - 0.01) < 1e-7  # clamp(u8/255, .01, .99), figbackend.nim:125
    with pytest.raises(ValueError):
        BackendContext().drawRect((0, 0, 1, 1), 0)  # "Backend drawRect unavailable", figbackend.nim:503-505


def test_figidx_capacity():
    lst = RenderList()
    root = lst.addRoot(Fig(kind=FigKind.nkFrame))
    with pytest.raises(OverflowError):
        for _ in range(40000):
            lst.addChild(root, Fig(kind=FigKind.nkFrame))


# ----------------------------------------------------------------------------------------------------------------------
# Curve drawables: the reference's own call-count pins (tests/ttransform.nim:269-525), restated one to one.
def _drawable_draws(ops_, stroke=None, node_steps=0, box=(5.0, 7.0, 30.0, 20.0)):
    from figdraw_b200.fignodes import fill as solid_fill

    node = Fig(kind=FigKind.nkDrawable, screenBox=rect(*box), drawSteps=node_steps)
    node.drawStroke = stroke or RenderStroke(weight=2.0, fill=solid_fill(rgba(255, 0, 0, 255)))
    node.drawOps = list(ops_)
    r = Renders()
    r.addRoot(0, node)
    tb = TraceBackend()
    renderRoot(tb, r)
    calls = tb._buf[: tb._n]
    return calls[calls["op"] >= 32]


def test_curve_drawables_match_the_reference_call_counts():
    from figdraw_b200.fignodes import StrokeCap, StrokeJoin, drawableArc, drawableBezier, drawableLine
    from figdraw_b200.fignodes import fill as solid_fill

    red = solid_fill(rgba(255, 0, 0, 255))
    # :269 quadratic bezier = one sdf op, whatever `steps` says
    assert len(_drawable_draws([drawableBezier([(0, 0), (10, 20), (20, 0)], steps=4)])) == 1
    # :297 round capped line = segment + 2 caps; :316 square capped line = one extended segment
    assert len(_drawable_draws([drawableLine((0, 0), (10, 0))], RenderStroke(weight=2.0, fill=red, cap=StrokeCap.scRound))) == 3
    assert len(_drawable_draws([drawableLine((0, 0), (10, 0))], RenderStroke(weight=2.0, fill=red, cap=StrokeCap.scSquare))) == 1
    # :335 higher order bezier -> quadratic sdf spans
    d = _drawable_draws([drawableBezier([(0, 0), (10, 20), (20, -10), (30, 0)], steps=4)])
    assert len(d) == 4 and all(int(o) == Op.BEZIER for o in d["op"])
    # :364 adaptive decomposition grows with screen size
    small = _drawable_draws([drawableBezier([(0, 0), (4, 20), (8, -20), (12, 0)])])
    large = _drawable_draws([drawableBezier([(0, 0), (40, 200), (80, -200), (120, 0)])])
    assert 0 < len(small) < len(large)
    # :388 arc -> quadratic spans; :411 adaptive arcs
    assert len(_drawable_draws([drawableArc((10, 10), 8.0, 0.0, 1.5707964, steps=4)])) == 4
    small = _drawable_draws([drawableArc((16, 16), 8.0, 0.0, 3.1415927)])
    large = _drawable_draws([drawableArc((90, 90), 80.0, 0.0, 3.1415927)])
    assert 0 < len(small) < len(large)
    # :454 explicit bevel joins: 4 spans + 3 joins (filled quads), butt caps draw nothing
    d = _drawable_draws([drawableArc((10, 10), 8.0, 0.0, 1.5707964, steps=4)],
                        RenderStroke(weight=2.0, fill=red, cap=StrokeCap.scButt, join=StrokeJoin.sjBevel))
    assert len(d) == 7 and sum(int(o) == Op.FILLED_QUAD for o in d["op"]) == 3
    # :479 node steps are the default for curve ops
    d = _drawable_draws([drawableBezier([(0, 0), (10, 20), (20, 0)]), drawableArc((20, 10), 8.0, 0.0, 1.5707964, steps=2)],
                        node_steps=4, box=(5.0, 7.0, 40.0, 30.0))
    assert len(d) == 3


def test_quadratic_sdf_padding_stays_in_physical_pixels():
    """tests/ttransform.nim:510-524: uiScale 2 -> rect 48 x 18."""
    from figdraw_b200.fignodes import drawableBezier

    setFigUiScale(2.0)
    try:
        d = _drawable_draws([drawableBezier((0, 0), (10, 10), (20, 0))])
        assert len(d) == 1
        assert abs(float(d[0]["f"][2]) - 48.0) < 1e-4 and abs(float(d[0]["f"][3]) - 18.0) < 1e-4
    finally:
        setFigUiScale(1.0)


# ----------------------------------------------------------------------------------------------------------------------
# The remaining "nkTransform render behavior" pins of tests/ttransform.nim (:147-267, :421-452, :526-547).
def _records(renders):
    tb = TraceBackend()
    renderRoot(tb, renders)
    return tb._buf[: tb._n].copy()


def _apply_transforms(calls, draw_index):
    """Position of draw `draw_index`'s rect origin under the recorded transform stack (what RecordingBackend reports)."""
    m, stack, k = np.eye(4, dtype=np.float64), [], -1
    for c in calls:
        op = int(c["op"])
        if op == Op.SAVE_TRANSFORM:
            stack.append(m.copy())
        elif op == Op.RESTORE_TRANSFORM:
            m = stack.pop()
        elif op == Op.TRANSLATE:
            t = np.eye(4)
            t[0, 3], t[1, 3] = c["f"][0], c["f"][1]
            m = m @ t
        elif op == Op.SCALE:
            m = m @ np.diag([c["f"][0], c["f"][1], 1.0, 1.0])
        elif op == Op.APPLY_TRANSFORM:
            m = m @ np.asarray(c["f"][:16], dtype=np.float64).reshape(4, 4).T  # vmath is column-major
        elif op >= 32:
            k += 1
            if k == draw_index:
                p = m @ np.array([c["f"][0], c["f"][1], 0.0, 1.0])
                return float(p[0]), float(p[1])
    raise IndexError(draw_index)


def test_corner_axes_reach_the_backend():
    from figdraw_b200.fignodes import BackdropBlurStyle
    from figdraw_b200.fignodes import fill as solid_fill

    r = Renders()
    r.addRoot(0, Fig(kind=FigKind.nkRectangle, screenBox=rect(5, 7, 40, 20), fill=solid_fill(rgba(255, 0, 0, 255)),
                     flags=FigFlags.NfEllipticalCorners, corners=(12, 10, 8, 6), cornerRadiiY=(4, 5, 6, 7)))
    d = _records(r)
    d = d[d["op"] == Op.ROUNDED_RECT]
    assert len(d) == 1 and list(d[0]["f"][4:8]) == [12, 10, 8, 6] and list(d[0]["f"][8:12]) == [4, 5, 6, 7]
    r = Renders()
    r.addRoot(0, Fig(kind=FigKind.nkRectangle, screenBox=rect(5, 7, 40, 20), fill=solid_fill(rgba(255, 0, 0, 255)),
                     corners=(12, 10, 8, 6)))
    d = _records(r)
    d = d[d["op"] == Op.ROUNDED_RECT]
    assert len(d) == 1 and list(d[0]["f"][4:8]) == list(d[0]["f"][8:12])  # circular corners promoted to equal axes
    r = Renders()
    r.addRoot(0, Fig(kind=FigKind.nkBackdropBlur, flags=FigFlags.NfEllipticalCorners, screenBox=rect(5, 7, 40, 20),
                     corners=(12, 10, 8, 6), cornerRadiiY=(4, 5, 6, 7), backdropBlur=BackdropBlurStyle(blur=10.0)))
    d = _records(r)
    d = d[d["op"] == Op.BACKDROP_BLUR]
    assert len(d) == 1 and list(d[0]["f"][4:8]) == [12, 10, 8, 6] and list(d[0]["f"][8:12]) == [4, 5, 6, 7]


def test_transform_nodes_apply_to_children():
    from figdraw_b200.fignodes import TransformStyle, drawableRect
    from figdraw_b200.fignodes import fill as solid_fill

    def child():
        n = Fig(kind=FigKind.nkDrawable, screenBox=rect(0, 0, 1, 1), fill=solid_fill(rgba(255, 0, 0, 255)))
        n.drawOps = [drawableRect(rect(2, 2, 1, 1))]
        return n

    r = Renders()
    root = r.addRoot(0, Fig(kind=FigKind.nkTransform, transform=TransformStyle(translation=(5.0, -4.0))))
    r.addChild(0, root, child())
    calls = _records(r)
    assert (calls["op"] >= 32).sum() == 1
    x, y = _apply_transforms(calls, 0)
    assert abs(x - 7.0) < 1e-4 and abs(y - (-2.0)) < 1e-4
    r = Renders()
    m = np.diag([2.0, 3.0, 1.0, 1.0]).astype(np.float32)  # scale(vec3(2, 3, 1))
    root = r.addRoot(0, Fig(kind=FigKind.nkTransform,
                            transform=TransformStyle(translation=(10.0, 20.0), matrix=m.T.reshape(16).tolist(), useMatrix=True)))
    r.addChild(0, root, child())
    calls = _records(r)
    x, y = _apply_transforms(calls, 0)
    assert abs(x - 14.0) < 1e-4 and abs(y - 26.0) < 1e-4


def test_ellipse_drawables_and_drawable_aa():
    from figdraw_b200.fignodes import drawableEllipse, drawableRect
    from figdraw_b200.fignodes import fill as solid_fill

    n = Fig(kind=FigKind.nkDrawable, screenBox=rect(5, 7, 30, 20), fill=solid_fill(rgba(20, 40, 80, 255)))
    n.drawStroke = RenderStroke(weight=2.0, fill=solid_fill(rgba(255, 0, 0, 255)))
    n.drawOps = [drawableEllipse((10.0, 8.0), (6.25, 3.5))]
    r = Renders()
    r.addRoot(0, n)
    d = _records(r)
    d = d[d["op"] >= 32]
    assert [int(u[0]) for u in d["u"]] == [SdfMode.sdfModeClipAA, SdfMode.sdfModeAnnularAA]
    for c in d:
        assert list(c["f"][4:8]) == [6.25] * 4 and list(c["f"][8:12]) == [3.5] * 4
    assert [float(v) for v in d[0]["f"][0:4]] == [8.75, 11.5, 12.5, 7.0]
    assert len(_drawable_draws([drawableEllipse((10.0, 10.0), (8.0, 0.0))])) == 0
    # drawAa overrides the backend AA factor and restores it (:526-547)
    n = Fig(kind=FigKind.nkDrawable, screenBox=rect(5, 7, 40, 30), fill=solid_fill(rgba(255, 0, 0, 255)), drawAa=0.75)
    n.drawOps = [drawableRect(rect(2, 3, 10, 8))]
    r = Renders()
    r.addRoot(0, n)
    calls = _records(r)
    aa = [float(c["f"][0]) for c in calls if int(c["op"]) == Op.SET_AA]
    assert (calls["op"] >= 32).sum() == 1 and len(aa) == 2
    assert abs(aa[0] - 0.75) < 1e-4 and abs(aa[1] - 1.2) < 1e-4


def test_text_node_call_order_and_subpixel_positioning():
    """renderText (figrender.nim:417-497): selection rects, decorations, then glyphs; with subpixel positioning the glyph x is
    snapped and its fraction goes to setTextSubpixelShift before the draw and back to 0 after it."""
    from figdraw_b200.fignodes import Glyph
    from figdraw_b200.fignodes import fill as solid_fill

    n = Fig(kind=FigKind.nkText, screenBox=rect(10, 20, 200, 40), fill=solid_fill(rgba(0, 120, 255, 90)),
            flags=FigFlags.NfSelectText)
    n.selectionRects = [rect(4, 2, 0.25, 12), rect(4, 16, 30, 0)]            # width widened to 1, zero height skipped
    n.decorations = [(rect(0, 14, 60, 1), solid_fill(rgba(255, 0, 0, 255))), (rect(0, 30, -2, 1), solid_fill(rgba(1, 2, 3, 255)))]
    n.glyphs = [Glyph(key=501, pos=(3.75, 5.0)), Glyph(key=999, pos=(12.5, 5.0)), Glyph(key=502, pos=(20.0, 5.0))]
    r = Renders()
    r.addRoot(0, n)
    tb = TraceBackend()
    tb.putImage(501, np.zeros((4, 4, 4), np.uint8))
    tb.putImage(502, np.zeros((4, 4, 4), np.uint8))
    tb.setTextSubpixelPositioningEnabled(True)
    renderRoot(tb, r)
    calls = tb._buf[: tb._n]
    seq = [Op(int(c["op"])) for c in calls][1:]  # after the SET_SUBPIXEL that enabled positioning
    assert seq == [Op.SAVE_TRANSFORM, Op.TRANSLATE,
                   Op.ROUNDED_RECT,                                    # one selection rect (the zero-height one is skipped)
                   Op.ROUNDED_RECT,                                    # one decoration (the negative-width one is skipped)
                   Op.SET_SUBPIXEL, Op.IMAGE, Op.SET_SUBPIXEL,        # glyph 501: shift .75, draw at x = 3, shift 0
                   Op.SET_SUBPIXEL, Op.SET_SUBPIXEL,                  # glyph 999 is not resident: shift, reset, skip
                   Op.SET_SUBPIXEL, Op.IMAGE, Op.SET_SUBPIXEL,        # glyph 502 on a whole pixel
                   Op.SET_SUBPIXEL, Op.RESTORE_TRANSFORM]
    sel = calls[3]
    assert [float(v) for v in sel["f"][0:4]] == [4.0, 2.0, 1.0, 12.0] and int(sel["u"][3]) == rgba(0, 120, 255, 90)
    shifts = [float(c["f"][0]) for c in calls[1:] if int(c["op"]) == Op.SET_SUBPIXEL]
    assert shifts == [0.75, 0.0, 0.5, 0.0, 0.0, 0.0, 0.0]
    img = [c for c in calls if int(c["op"]) == Op.IMAGE]
    assert [float(c["f"][0]) for c in img] == [3.0, 20.0]
