"""Structural render checks of the reference restated ("assert coarse properties; no golden"):
tests/trender_image_msdf_invert.nim -- NfInvertY keeps mirrored nkImage / nkMsdfImage upright under a y-mirroring
nkTransform.  Runs on the oracle (CPU) and on the CUDA backend (-m gpu), which must also agree with each other."""
import numpy as np
import pytest

from figdraw_b200.fignodes import (Fig, FigFlags, FigKind, ImageStyle, MsdfImageStyle, RenderList, Renders, TransformStyle, fill,
                                   rect, rgba)
from figdraw_b200.scenes_synth import trace_renders
from oracle import oracle

W, H, S = 720, 520, 180.0
BITMAP_ID, MSDF_ID = 0x1A2B3C4D5E6F, 0x0F1E2D3C4B5A
RECTS = {"image_base": (40, 50), "image_noinvert": (260, 50), "image_invert": (480, 50),
         "msdf_base": (40, 270), "msdf_noinvert": (260, 270), "msdf_invert": (480, 270)}


def asymmetric_image():
    img = np.zeros((24, 24, 4), np.uint8)
    img[:8] = (0, 0, 0, 255)
    img[8:] = (255, 230, 0, 255)
    return img


def synthetic_msdf_field():
    img = np.zeros((24, 24, 4), np.uint8)
    img[:8] = (255, 255, 255, 255)
    img[8:] = (0, 0, 0, 255)
    return img


def invert_scene():
    lst = RenderList()
    lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(0, 0, W, H), fill=fill(rgba(255, 255, 255, 255))))
    white, black = fill(rgba(255, 255, 255, 255)), fill(rgba(0, 0, 0, 255))

    def image(pos, flags=0):
        return Fig(kind=FigKind.nkImage, screenBox=rect(pos[0], pos[1], S, S), flags=FigFlags(flags),
                   image=ImageStyle(id=BITMAP_ID, fill=white))

    def msdf(pos, flags=0):
        return Fig(kind=FigKind.nkMsdfImage, screenBox=rect(pos[0], pos[1], S, S), flags=FigFlags(flags),
                   msdfImage=MsdfImageStyle(id=MSDF_ID, fill=black, pxRange=4.0, sdThreshold=0.5))

    def mirrored(p):  # mirroredInputRect: the final rect seen through y -> h - y
        return (p[0], H - p[1] - S)

    lst.addRoot(image(RECTS["image_base"]))
    lst.addRoot(msdf(RECTS["msdf_base"]))
    m = np.diag([1.0, -1.0, 1.0, 1.0]).astype(np.float32)
    root = lst.addRoot(Fig(kind=FigKind.nkTransform,
                           transform=TransformStyle(translation=(0.0, float(H)), matrix=m.reshape(16).tolist(), useMatrix=True)))
    lst.addChild(root, image(mirrored(RECTS["image_noinvert"])))
    lst.addChild(root, image(mirrored(RECTS["image_invert"]), FigFlags.NfInvertY))
    lst.addChild(root, msdf(mirrored(RECTS["msdf_noinvert"])))
    lst.addChild(root, msdf(mirrored(RECTS["msdf_invert"]), FigFlags.NfInvertY))
    r = Renders()
    r.setLayer(0, lst)
    return r


def invert_trace():
    return trace_renders(invert_scene(), W, H, images=[(BITMAP_ID, asymmetric_image()), (MSDF_ID, synthetic_msdf_field())])


def row_profile(img, pos):
    x0, y0 = int(pos[0]), int(pos[1])
    block = img[y0:y0 + int(S), x0:x0 + int(S), :3].astype(np.int64)
    return (255 - block).sum(axis=(1, 2))


def check_invert_properties(img):
    p = {k: row_profile(img, v) for k, v in RECTS.items()}
    assert all(len(v) > 0 for v in p.values())
    assert p["image_base"].max() - p["image_base"].min() > 500 and p["msdf_base"].max() - p["msdf_base"].min() > 500
    direct = lambda a, b: int(np.abs(a - b).sum())
    flipped = lambda a, b: int(np.abs(a - b[::-1]).sum())
    for kind in ("image", "msdf"):
        base, noinv, inv = p[f"{kind}_base"], p[f"{kind}_noinvert"], p[f"{kind}_invert"]
        assert flipped(base, noinv) < direct(base, noinv), kind     # without the flag the mirror shows
        assert direct(base, inv) <= flipped(base, inv), kind        # NfInvertY keeps it upright


def test_invert_y_on_the_oracle():
    check_invert_properties(oracle.render_trace(invert_trace()))


@pytest.mark.gpu
def test_invert_y_on_the_cuda_backend():
    from figdraw_b200.cuda_context import render_trace

    tr = invert_trace()
    got = render_trace(tr)
    check_invert_properties(got)
    want = oracle.render_trace(tr)
    assert int(np.abs(got.astype(np.int16) - want.astype(np.int16)).max()) <= 2
