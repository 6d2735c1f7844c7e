"""Native scene flattening (fdc_flatten_renders) against the per-call front-end (figrender.py):
the `fdc_call` records must be byte-identical.  Pure host code: runs without a GPU."""
import math

import numpy as np
import pytest

from figdraw_b200 import figrender, native_scene, scenes, scenes_synth as ss
from figdraw_b200.figbackend import TraceBackend
from figdraw_b200.fignodes import (BackdropBlurStyle, Fig, FigFlags, FigKind, FillGradientAxis, Glyph, ImageStyle, MsdfImageStyle,
                                   Renders, RenderShadow, RenderStroke, ShadowStyle, StrokeCap, StrokeJoin, TransformStyle, drawableArc, drawableBezier,
                                   drawableCircle, drawableEllipse, drawableLine, drawableRect, figCircle, figLine, fill, linear,
                                   rect, rgba)
from figdraw_b200.scenes_synth import Rng


def per_call_records(renders, w, h, images=(), ui_scale=1.0, pixel_scale=1.0, subpixel=False):
    figrender.setFigUiScale(ui_scale)
    try:
        tb = TraceBackend(pixelScale=pixel_scale)
        for key, img in images:
            tb.putImage(key, img)
        if subpixel:
            tb._subpixel = True
        figrender.renderFrame(tb, renders, (float(w), float(h)))
        return tb.trace().calls
    finally:
        figrender.setFigUiScale(1.0)


def native_records(renders, images=(), ui_scale=1.0, pixel_scale=1.0, subpixel=False):
    packed = native_scene.pack_renders(renders)
    return native_scene.flatten(packed, ui_scale=ui_scale, pixel_scale=pixel_scale, subpixel_enabled=subpixel,
                                image_keys=[k for k, _ in images])


def assert_same(a, b):
    assert len(a) == len(b), f"{len(a)} native records vs {len(b)} per-call records"
    if a.tobytes() == b.tobytes():
        return
    for i in range(len(a)):
        if a[i].tobytes() != b[i].tobytes():
            raise AssertionError(f"record {i} differs:\n native   {a[i]}\n per-call {b[i]}")


GOLDEN_BUILDERS = [scenes.rgb_boxes_sdf, scenes.linear_gradient, scenes.layers_clip, scenes.layers_rect_mask,
                   scenes.mixed_rect_mask_batch, scenes.line_rect, scenes.circle_rect]


@pytest.mark.parametrize("builder", GOLDEN_BUILDERS, ids=lambda b: b.__name__)
def test_golden_scenes_flatten_identically(builder):
    r = builder(800.0, 600.0)
    assert_same(native_records(r), per_call_records(r, 800, 600))


def test_image_scene_and_missing_images():
    r = scenes.image_scene(800.0, 600.0)
    img = [(scenes.IMG1_KEY, scenes.load_img1())]
    assert_same(native_records(r, img), per_call_records(r, 800, 600, img))
    assert_same(native_records(r), per_call_records(r, 800, 600))  # image not resident: drawImage still issued


def test_config_scenes_flatten_identically():
    r2 = ss.renderlist_100(1920.0, 1080.0)
    assert_same(native_records(r2), per_call_records(r2, 1920, 1080))
    imgs = ss.text_page_images()
    r3 = ss.text_page(1280.0, 720.0, n_glyphs=1500, msdf_glyphs=200)
    assert_same(native_records(r3, imgs), per_call_records(r3, 1280, 720, imgs))
    # glyphs whose bitmap is not resident are skipped (figrender.nim:470-476): half the keys only
    half = imgs[: len(imgs) // 2]
    assert_same(native_records(r3, half), per_call_records(r3, 1280, 720, half))
    for rect_mask in (False, True):
        r4 = ss.clip_mask_table(1920.0, 1080.0, rows=40, cols=6, rect_mask=rect_mask)
        assert_same(native_records(r4), per_call_records(r4, 1920, 1080))


def test_ui_scale_pixel_scale_and_subpixel_state():
    r = ss.renderlist_100(960.0, 540.0, copies=20)
    assert_same(native_records(r, ui_scale=1.5, pixel_scale=2.0), per_call_records(r, 960, 540, ui_scale=1.5, pixel_scale=2.0))
    imgs = ss.text_page_images()
    r3 = ss.text_page(640.0, 360.0, n_glyphs=300)
    assert_same(native_records(r3, imgs, subpixel=True), per_call_records(r3, 640, 360, imgs, subpixel=True))


def random_renders(seed: int) -> Renders:
    """Random trees over every node kind, flag and drawable op the front-end restates."""
    rng = Rng(seed)
    u = lambda lo=0.0, hi=1.0: float(rng.uniform(1, lo, hi)[0])
    ri = lambda n: int(rng.uniform(1, 0.0, float(n))[0]) % n

    def col():
        return rgba(ri(256), ri(256), ri(256), (0, 90, 155, 255)[ri(4)])

    def rfill():
        k = ri(4)
        if k == 0:
            return fill(col())
        if k == 1:
            return linear(col(), col(), axis=FillGradientAxis(ri(4)))
        return linear(col(), col(), col(), axis=FillGradientAxis(ri(4)), midPos=ri(256))

    def rnode(depth):
        kind = (FigKind.nkRectangle, FigKind.nkRectangle, FigKind.nkDrawable, FigKind.nkText, FigKind.nkImage, FigKind.nkMsdfImage,
                FigKind.nkMtsdfImage, FigKind.nkBackdropBlur, FigKind.nkTransform, FigKind.nkFrame)[ri(10)]
        n = Fig(kind=kind, screenBox=rect(u(-50, 600), u(-50, 400), u(-5, 300), u(-5, 200)), fill=rfill(),
                corners=tuple(u(0, 40) for _ in range(4)), cornerRadiiY=tuple(u(0, 40) for _ in range(4)))
        if ri(4) == 0:
            n.rotation = u(-180, 180)
        fl = 0
        for f, p in ((FigFlags.NfClipContent, 6), (FigFlags.NfRectMaskContent, 6), (FigFlags.NfEllipticalCorners, 4),
                     (FigFlags.NfInvertY, 5), (FigFlags.NfDisableRender, 12)):
            if ri(p) == 0:
                fl |= int(f)
        n.flags = FigFlags(fl)
        if kind == FigKind.nkRectangle:
            n.shadows = [RenderShadow(style=ShadowStyle(ri(3)), fill=rfill(), blur=u(-2, 20), spread=u(-2, 12), x=u(-8, 8), y=u(-8, 8))
                         for _ in range(ri(5))]
            n.stroke = RenderStroke(weight=u(-1, 6), fill=rfill())
        elif kind == FigKind.nkDrawable:
            n.drawStroke = RenderStroke(weight=u(-1, 8), fill=rfill(), cap=StrokeCap(ri(4)), join=StrokeJoin(ri(4)))
            n.drawAa = (0.0, 0.0, 1.2, 0.8, 2.0)[ri(5)]
            n.drawSteps = (0, 0, 3, 6)[ri(4)]
            for _ in range(ri(5)):
                k = ri(10)
                if k == 0:
                    n.drawOps.append(drawableLine((u(0, 100), u(0, 100)), (u(0, 100), u(0, 100))))
                elif k == 1:
                    n.drawOps.append(drawableCircle((u(0, 100), u(0, 100)), u(-2, 40)))
                elif k == 2:
                    n.drawOps.append(drawableRect(rect(u(0, 50), u(0, 50), u(0, 90), u(0, 90)), tuple(ri(20) for _ in range(4))))
                elif k == 3:
                    n.drawOps.append(drawableEllipse((u(0, 100), u(0, 100)), (u(-2, 40), u(-2, 30))))
                elif k == 4:
                    n.drawOps.append(drawableBezier((u(0, 100), u(0, 100)), (u(0, 100), u(0, 100)), (u(0, 100), u(0, 100))))
                elif k == 5:  # collinear controls: the line fallback of the Bezier path
                    a, d = (u(0, 50), u(0, 50)), (u(1, 20), u(1, 20))
                    n.drawOps.append(drawableBezier(a, (a[0] + d[0], a[1] + d[1]), (a[0] + 2 * d[0], a[1] + 2 * d[1])))
                elif k in (6, 7):  # higher-order (and 2-control) Beziers: adaptive or fixed spans, caps and joins
                    nc = (2, 4, 5, 7)[ri(4)]
                    sc = (1.0, 4.0, 12.0)[ri(3)]
                    n.drawOps.append(drawableBezier([(u(0, 60) * sc, u(-30, 60) * sc) for _ in range(nc)], steps=(0, 0, 2, 5)[ri(4)]))
                else:
                    n.drawOps.append(drawableArc((u(0, 100), u(0, 100)), u(-2, 120), u(-7, 7), (0.0, u(-7, 7), u(-2, 2))[ri(3)],
                                                 steps=(0, 0, 1, 4)[ri(4)]))
        elif kind == FigKind.nkText:
            n.glyphs = [Glyph(key=1000 + ri(40), pos=(u(0, 300), u(0, 100)), fill=rfill()) for _ in range(ri(12))]
            if ri(2):
                n.flags = FigFlags(int(n.flags) | int(FigFlags.NfSelectText))
            n.selectionRects = [rect(u(0, 200), u(0, 80), u(-1, 60), u(-2, 20)) for _ in range(ri(3))]
            n.decorations = [(rect(u(0, 200), u(0, 80), u(-5, 120), u(-1, 3)), rfill()) for _ in range(ri(3))]
        elif kind == FigKind.nkImage:
            n.image = ImageStyle(id=(0, 77, 78)[ri(3)], fill=rfill())
        elif kind in (FigKind.nkMsdfImage, FigKind.nkMtsdfImage):
            st = MsdfImageStyle(id=(0, 88)[ri(2)], fill=rfill(), pxRange=(0.0, 4.0, 6.0)[ri(3)], sdThreshold=(0.0, 0.5, 0.4, 1.5)[ri(4)],
                                strokeWeight=u(-1, 4))
            if kind == FigKind.nkMsdfImage:
                n.msdfImage = st
            else:
                n.mtsdfImage = st
        elif kind == FigKind.nkBackdropBlur:
            n.backdropBlur = BackdropBlurStyle(blur=u(-4, 30))
        elif kind == FigKind.nkTransform:
            m = np.eye(4, dtype=np.float32)
            a = u(-1, 1)
            m[0, 0], m[0, 1], m[1, 0], m[1, 1] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
            n.transform = TransformStyle(translation=(u(-20, 20), (0.0, u(-20, 20))[ri(2)]), matrix=m.reshape(16).tolist(),
                                         useMatrix=bool(ri(2)))
        return n

    r = Renders()
    for lvl in (3, -2, 0)[: 1 + ri(3)]:
        lst = r[lvl]
        for _ in range(1 + ri(4)):
            stack = [lst.addRoot(rnode(0))]
            for _ in range(ri(30)):
                parent = stack[ri(len(stack))]
                stack.append(lst.addChild(parent, rnode(1)))
        lst.addRoot(figLine(u(0, 300), u(0, 300), u(0, 300), u(0, 300), col(), u(0, 6)))
        lst.addRoot(figCircle(u(0, 300), u(0, 300), rfill(), u(0, 40)))
    return r


@pytest.mark.parametrize("seed", range(40))
def test_random_trees_flatten_identically(seed):
    r = random_renders(seed)
    keys = [(1000 + k, np.zeros((4, 4, 4), np.uint8)) for k in range(0, 40, 3)]
    ui = (1.0, 1.0, 2.0, 0.75)[seed % 4]
    sub = seed % 3 == 0  # subpixel positioning: glyph x snapped, the fraction goes to setTextSubpixelShift
    assert_same(native_records(r, keys, ui_scale=ui, subpixel=sub), per_call_records(r, 640, 480, keys, ui_scale=ui, subpixel=sub))


def test_unknown_drawable_kind_is_refused():
    r = Renders()
    n = Fig(kind=FigKind.nkDrawable, screenBox=rect(0, 0, 10, 10), fill=fill(rgba(0, 0, 0, 255)))
    n.drawStroke = RenderStroke(weight=2.0, fill=fill(rgba(0, 0, 0, 255)))
    n.drawOps.append(drawableLine((0, 0), (5, 5)))
    r.addRoot(0, n)
    packed = native_scene.pack_renders(r)
    packed.ops[0]["kind"] = 17
    with pytest.raises(ValueError):
        native_scene.flatten(packed)


def test_curve_drawables_flatten_identically():
    """Every curve case the reference's own tests pin (tests/ttransform.nim:269-525), natively and per call."""
    red = fill(rgba(255, 0, 0, 255))
    cases = [
        ([drawableBezier([(0, 0), (10, 20), (20, 0)], steps=4)], RenderStroke(weight=2.0, fill=red), 0),
        ([drawableLine((0, 0), (10, 0))], RenderStroke(weight=2.0, fill=red, cap=StrokeCap.scRound), 0),
        ([drawableLine((0, 0), (10, 0))], RenderStroke(weight=2.0, fill=red, cap=StrokeCap.scSquare), 0),
        ([drawableBezier([(0, 0), (10, 20), (20, -10), (30, 0)], steps=4)], RenderStroke(weight=2.0, fill=red), 0),
        ([drawableBezier([(0, 0), (40, 200), (80, -200), (120, 0)])], RenderStroke(weight=2.0, fill=red), 0),
        ([drawableArc((10, 10), 8.0, 0.0, 1.5707964, steps=4)], RenderStroke(weight=2.0, fill=red), 0),
        ([drawableArc((90, 90), 80.0, 0.0, 3.1415927)], RenderStroke(weight=2.0, fill=red), 0),
        ([drawableArc((10, 10), 8.0, 0.0, 1.5707964, steps=4)],
         RenderStroke(weight=2.0, fill=red, cap=StrokeCap.scButt, join=StrokeJoin.sjBevel), 0),
        ([drawableArc((10, 10), 8.0, 0.3, -2.5)], RenderStroke(weight=3.0, fill=red, cap=StrokeCap.scSquare, join=StrokeJoin.sjMiter), 0),
        ([drawableBezier([(0, 0), (10, 20), (20, 0)]), drawableArc((20, 10), 8.0, 0.0, 1.5707964, steps=2)],
         RenderStroke(weight=2.0, fill=red), 4),
        ([drawableBezier([(0, 0), (30, 5)])], RenderStroke(weight=2.0, fill=red, cap=StrokeCap.scSquare), 0),
    ]
    for ui in (1.0, 2.0):
        for ops_, stroke, node_steps in cases:
            n = Fig(kind=FigKind.nkDrawable, screenBox=rect(5.0, 7.0, 30.0, 20.0), drawSteps=node_steps)
            n.drawStroke = stroke
            n.drawOps = list(ops_)
            r = Renders()
            r.addRoot(0, n)
            assert_same(native_records(r, ui_scale=ui), per_call_records(r, 200, 200, ui_scale=ui))
