"""Scene builders for BASELINE.json's configs.

cfg1: the reference's six golden-test scenes, restated node for node from its tests.
cfg2..cfg5: synthetic scenes of the named shapes (see `scenes_synth.py`).

Each builder returns `Renders` exactly as the reference test's `makeRenderTree` does, so the same
front-end (figrender.py) turns it into backend calls.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Tuple

import numpy as np

from .figbackend import TraceBackend, Trace
from .fignodes import (Fig, FigFlags, FigKind, FillGradientAxis, ImageStyle, RenderList, RenderShadow, RenderStroke,
                       Renders, ShadowStyle, figCircle, figLine, linear, rect, rgba)
from .figrender import renderFrame, setFigUiScale

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
IMG1_KEY = 0x1D6A5E11  # stands for hash("img1.png").ImageId


def _single_layer(lst: RenderList) -> Renders:
    r = Renders()
    r.setLayer(0, lst)
    return r


def rgb_boxes_sdf(w: float, h: float) -> Renders:
    """tests/trender_rgb_boxes_sdf.nim:13-101."""
    lst = RenderList()
    root = lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(0, 0, w, h), fill=rgba(255, 255, 255, 255)))
    lst.addChild(root, Fig(kind=FigKind.nkRectangle, corners=(10, 20, 30, 40), screenBox=rect(60, 60, 220, 140),
                           fill=rgba(220, 40, 40, 255), stroke=RenderStroke(weight=5.0, fill=rgba(0, 0, 0, 255))))
    lst.addChild(root, Fig(
        kind=FigKind.nkRectangle, screenBox=rect(320, 120, 220, 140),
        fill=linear(rgba(24, 128, 72, 255), rgba(40, 180, 90, 255), rgba(54, 206, 170, 255),
                    axis=FillGradientAxis.fgaX, midPos=140),
        shadows=[RenderShadow(style=ShadowStyle.DropShadow, blur=10, spread=10, x=10, y=10, fill=rgba(0, 0, 0, 55))]))
    lst.addChild(root, Fig(
        kind=FigKind.nkRectangle, screenBox=rect(180, 300, 220, 140), fill=rgba(60, 90, 220, 255),
        shadows=[
            RenderShadow(style=ShadowStyle.InnerShadow, blur=12, spread=0, x=-6, y=-6,
                         fill=linear(rgba(25, 25, 25, 90), rgba(65, 65, 65, 175), axis=FillGradientAxis.fgaDiagTLBR)),
            RenderShadow(style=ShadowStyle.InnerShadow, blur=12, spread=0, x=6, y=6,
                         fill=linear(rgba(255, 255, 255, 255), rgba(205, 205, 205, 115),
                                     axis=FillGradientAxis.fgaDiagTLBR)),
        ]))
    return _single_layer(lst)


def linear_gradient(w: float, h: float) -> Renders:
    """tests/trender_linear_gradient.nim:13-96."""
    lst = RenderList()
    root = lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(0, 0, w, h), fill=rgba(255, 255, 255, 255)))
    lst.addChild(root, Fig(kind=FigKind.nkRectangle, screenBox=rect(80, 80, 440, 120), corners=(12, 12, 12, 12),
                           fill=linear(rgba(220, 40, 40, 255), rgba(40, 200, 90, 255), rgba(50, 90, 225, 255),
                                       axis=FillGradientAxis.fgaX, midPos=128)))
    lst.addChild(root, Fig(kind=FigKind.nkRectangle, screenBox=rect(80, 240, 220, 220), corners=(10, 10, 10, 10),
                           fill=linear(rgba(240, 210, 40, 255), rgba(110, 60, 210, 255), axis=FillGradientAxis.fgaY)))
    lst.addChild(root, Fig(kind=FigKind.nkRectangle, screenBox=rect(340, 250, 240, 180), fill=rgba(0, 0, 0, 0),
                           stroke=RenderStroke(weight=20, fill=linear(rgba(245, 70, 70, 255), rgba(70, 115, 245, 255),
                                                                      axis=FillGradientAxis.fgaX))))
    lst.addChild(root, Fig(
        kind=FigKind.nkRectangle, screenBox=rect(610, 300, 150, 200), fill=rgba(245, 245, 245, 255),
        shadows=[RenderShadow(style=ShadowStyle.DropShadow, blur=6, spread=14, x=0, y=0,
                              fill=linear(rgba(255, 70, 70, 170), rgba(70, 110, 255, 170),
                                          axis=FillGradientAxis.fgaX))]))
    return _single_layer(lst)


def layers_clip(w: float, h: float, rectMask: bool = False) -> Renders:
    """tests/trender_layers_clip.nim:76-173 (float32 arithmetic as in the Nim source)."""
    f = np.float32
    w, h = f(w), f(h)
    bg, container, button = rgba(255, 255, 255, 255), rgba(208, 208, 208, 255), rgba(43, 159, 234, 255)
    cW, cH, cY = w * f(0.30), w * f(0.40), h * f(0.10)
    cLX, cRX = w * f(0.03), w * f(0.50)
    bX, bW, bH = cW * f(0.10), cW * f(1.30), cH * f(0.20)
    bY1, bY2, bY3 = cH * f(0.15), cH * f(0.45), cH * f(0.75)

    def box(r, color, z, clip=False, rmask=False, corners=10):
        flags = FigFlags(0)
        if clip:
            flags |= FigFlags.NfClipContent
        if rmask:
            flags |= FigFlags.NfRectMaskContent
        return Fig(kind=FigKind.nkRectangle, zlevel=z, screenBox=r, fill=color, corners=(corners,) * 4, flags=flags)

    bgList = RenderList()
    bgList.addRoot(Fig(kind=FigKind.nkRectangle, zlevel=-20, screenBox=rect(0, 0, w, h), fill=bg))
    l0 = RenderList()
    left = l0.addRoot(box(rect(cLX, cY, cW, cH), container, 0))
    right = l0.addRoot(box(rect(cRX, cY, cW, cH), container, 0, clip=not rectMask, rmask=rectMask))
    l0.addChild(left, box(rect(cLX + bX, cY + bY2, bW, bH), button, 0))
    l0.addChild(right, box(rect(cRX + bX, cY + bY2, bW, bH), button, 0))
    low, top = RenderList(), RenderList()
    low.addRoot(box(rect(cLX + bX, cY + bY3, bW, bH), button, -5))
    top.addRoot(box(rect(cLX + bX, cY + bY1, bW, bH), button, 20))
    low.addRoot(box(rect(cRX + bX, cY + bY3, bW, bH), button, -5))
    top.addRoot(box(rect(cRX + bX, cY + bY1, bW, bH), button, 20))
    r = Renders()
    r.setLayer(-20, bgList)
    r.setLayer(0, l0)
    r.setLayer(-5, low)
    r.setLayer(20, top)
    r.sort()
    return r


def layers_rect_mask(w: float, h: float) -> Renders:
    return layers_clip(w, h, rectMask=True)


def mixed_rect_mask_batch(w: float, h: float) -> Renders:
    """tests/trender_layers_clip.nim:181-221."""
    lst = RenderList()

    def root(r, color, rmask=False):
        flags = FigFlags.NfRectMaskContent if rmask else FigFlags(0)
        return lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=r, fill=color, flags=flags))

    root(rect(0, 0, w, h), rgba(255, 255, 255, 255))
    root(rect(32, 48, 96, 80), rgba(230, 70, 52, 255))
    m = root(rect(180, 48, 80, 80), rgba(218, 218, 218, 255), rmask=True)
    lst.addChild(m, Fig(kind=FigKind.nkRectangle, screenBox=rect(150, 72, 150, 34), fill=rgba(56, 168, 88, 255)))
    root(rect(310, 48, 96, 80), rgba(54, 118, 230, 255))
    return _single_layer(lst)


def line_rect(w: float, h: float) -> Renders:
    """tests/trender_extras.nim:18-37."""
    lst = RenderList()
    root = lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(0, 0, w, h), fill=rgba(255, 255, 255, 255)))
    lst.addChild(root, figLine(90.0, 120.0, 710.0, 470.0, rgba(0, 0, 0, 255), 48.0))
    return _single_layer(lst)


def circle_rect(w: float, h: float) -> Renders:
    """tests/trender_extras.nim:39-57."""
    lst = RenderList()
    root = lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(0, 0, w, h), fill=rgba(255, 255, 255, 255)))
    lst.addChild(root, figCircle(400.0, 300.0, rgba(0, 0, 0, 255), 110.0))
    return _single_layer(lst)


def image_scene(w: float, h: float) -> Renders:
    """tests/trender_image.nim:13-39."""
    lst = RenderList()
    root = lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(0, 0, w, h), fill=rgba(160, 160, 160, 255)))
    lst.addChild(root, Fig(kind=FigKind.nkImage, screenBox=rect(60, 60, 160, 160),
                           image=ImageStyle(fill=rgba(255, 255, 255, 255), id=IMG1_KEY)))
    return _single_layer(lst)


def load_img1() -> np.ndarray:
    """data/img1.png as straight-alpha RGBA8 (what reaches glTexSubImage2D, textures.nim:88-104)."""
    from PIL import Image

    return np.asarray(Image.open(os.path.join(GOLDEN_DIR, "img1.png")).convert("RGBA"), dtype=np.uint8)


# name -> (builder, width, height, golden png or None)
GOLDEN_SCENES: Dict[str, Tuple[Callable[[float, float], Renders], int, int, str]] = {
    "rgb_boxes_sdf": (rgb_boxes_sdf, 800, 600, "render_rgb_boxes_sdf.png"),
    "linear_gradient": (linear_gradient, 800, 600, "render_linear_gradient.png"),
    "layers_clip": (layers_clip, 800, 375, "render_layers_clip.png"),
    "circle_rect": (circle_rect, 800, 600, "render_circle_rect.png"),
    "line_rect": (line_rect, 800, 600, "render_line_rect.png"),
    "image": (image_scene, 800, 600, "render_image.png"),
}


def trace_scene(builder: Callable[[float, float], Renders], width: int, height: int, atlasSize: int = 2048,
                images=None, uiScale: float = 1.0) -> Trace:
    """Build the scene, run the front-end over a TraceBackend, return the recorded frame.
    atlasSize 2048 is what the reference's render tests use (tests/opengl_test_utils.nim:33)."""
    setFigUiScale(uiScale)
    tb = TraceBackend(atlasSize=atlasSize)
    for key, img in (images or []):
        tb.putImage(key, img)
    renders = builder(float(width), float(height))
    renderFrame(tb, renders, (float(width), float(height)))
    return tb.trace()


def golden_trace(name: str) -> Trace:
    builder, w, h, _png = GOLDEN_SCENES[name]
    images = [(IMG1_KEY, load_img1())] if name == "image" else None
    return trace_scene(builder, w, h, images=images)


def load_golden(name: str) -> np.ndarray:
    from PIL import Image

    return np.asarray(Image.open(os.path.join(GOLDEN_DIR, GOLDEN_SCENES[name][3])).convert("RGBA"), dtype=np.uint8)
