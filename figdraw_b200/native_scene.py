"""Marshals `Renders` (fignodes.py) into the POD arrays of the native front-end and drives it.

The reference keeps `Fig` as a <=256-byte variant record in `seq[Fig]` (fignodes.nim:44-97); the Nim shim would hand
`addr nodes[0]` over after copying the `seq` members (glyph arrangements, drawable ops) into side arrays.  This module
does the same from the Python mirror of the scene model, so that the tests can compare the native flattening
(`fdc_flatten_renders`, csrc/fdc_flatten.cu) record for record with the per-call front-end (figrender.py).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import abi
from .fignodes import DrawableKind, Fig, FigKind, Fill, FillKind, RenderList, Renders, RenderShadow, RenderStroke


def _put_fill(dst, fill: Fill) -> None:
    dst["kind"] = int(fill.kind)
    dst["axis"] = int(fill.axis)
    dst["mid_pos"] = int(fill.midPos) & 255
    if fill.kind == FillKind.flColor:
        dst["c"] = (fill.color & 0xFFFFFFFF, 0, 0)
    elif fill.kind == FillKind.flLinear2:
        dst["c"] = (fill.start & 0xFFFFFFFF, 0, fill.stop & 0xFFFFFFFF)
    else:
        dst["c"] = (fill.start & 0xFFFFFFFF, fill.mid & 0xFFFFFFFF, fill.stop & 0xFFFFFFFF)


def _put_stroke(dst, stroke: RenderStroke) -> None:
    dst["weight"] = stroke.weight
    _put_fill(dst["fill"], stroke.fill)
    dst["cap"] = int(stroke.cap)
    dst["join"] = int(stroke.join)


def _put_shadow(dst, sh: RenderShadow) -> None:
    dst["style"] = int(sh.style)
    _put_fill(dst["fill"], sh.fill)
    dst["blur"], dst["spread"], dst["x"], dst["y"] = sh.blur, sh.spread, sh.x, sh.y


@dataclass
class PackedScene:
    """POD arrays + the `fdc_scene` pointing into them (kept alive together)."""

    nodes: List[np.ndarray]
    roots: List[np.ndarray]
    glyphs: np.ndarray
    ops: np.ndarray
    lists: ctypes.Array
    points: Optional[np.ndarray] = None      # Bezier control points (x, y pairs) the ops index
    text_rects: Optional[np.ndarray] = None  # selection / decoration rects the text nodes index
    scene: Optional[abi.FdcScene] = None

    def __post_init__(self):
        if self.points is None:
            self.points = np.zeros(2, dtype=np.float32)
        if self.text_rects is None:
            self.text_rects = np.zeros(1, dtype=abi.TEXT_RECT_DTYPE)
        self.scene = abi.FdcScene(self.lists, len(self.nodes), self.glyphs.ctypes.data, self.text_rects.ctypes.data,
                                  self.ops.ctypes.data, self.points.ctypes.data)

    @property
    def n_nodes(self) -> int:
        return sum(len(n) for n in self.nodes)


def pack_renders(renders: Renders) -> PackedScene:
    glyphs: List[tuple] = []
    ops: List[tuple] = []
    node_arrays, root_arrays = [], []
    glyph_rows, op_rows, rect_rows = [], [], []
    for _lvl, lst in renders.pairs():
        arr = np.zeros(len(lst.nodes), dtype=abi.FIG_DTYPE)
        for i, n in enumerate(lst.nodes):
            r = arr[i]
            r["kind"] = int(n.kind)
            r["zlevel"] = int(n.zlevel)
            r["flags"] = int(n.flags)
            r["parent"] = int(n.parent)
            r["child_count"] = int(n.childCount)
            r["screen_box"] = n.screenBox.tuple()
            r["rotation"] = n.rotation
            _put_fill(r["fill"], n.fill)
            r["corners"] = tuple(n.corners)
            r["corner_radii_y"] = tuple(n.cornerRadiiY)
            pay = r["payload"]
            if n.kind == FigKind.nkRectangle:
                v = pay.view(abi.FIG_RECT_DTYPE)[0]
                if len(n.shadows) > 4:
                    raise ValueError("ShadowCount is 4 (figbasics.nim:12)")
                for k, sh in enumerate(n.shadows):
                    _put_shadow(v["shadows"][k], sh)
                _put_stroke(v["stroke"], n.stroke)
            elif n.kind == FigKind.nkText:
                v = pay[: abi.FIG_TEXT_DTYPE.itemsize].view(abi.FIG_TEXT_DTYPE)[0]
                v["first_glyph"], v["n_glyphs"] = len(glyph_rows), len(n.glyphs)
                glyph_rows.extend(n.glyphs)
                v["first_rect"], v["n_selection"], v["n_decoration"] = len(rect_rows), len(n.selectionRects), len(n.decorations)
                rect_rows.extend((r, None) for r in n.selectionRects)
                rect_rows.extend(n.decorations)
            elif n.kind == FigKind.nkDrawable:
                v = pay[: abi.FIG_DRAWABLE_DTYPE.itemsize].view(abi.FIG_DRAWABLE_DTYPE)[0]
                _put_stroke(v["stroke"], n.drawStroke)
                v["steps"], v["aa"] = int(n.drawSteps), n.drawAa
                v["first_op"], v["n_ops"] = len(op_rows), len(n.drawOps)
                op_rows.extend(n.drawOps)
            elif n.kind == FigKind.nkImage:
                v = pay[: abi.FIG_IMAGE_DTYPE.itemsize].view(abi.FIG_IMAGE_DTYPE)[0]
                v["id"] = int(n.image.id) & 0xFFFFFFFFFFFFFFFF
                _put_fill(v["fill"], n.image.fill)
            elif n.kind in (FigKind.nkMsdfImage, FigKind.nkMtsdfImage):
                st = n.msdfImage if n.kind == FigKind.nkMsdfImage else n.mtsdfImage
                v = pay[: abi.FIG_MSDF_DTYPE.itemsize].view(abi.FIG_MSDF_DTYPE)[0]
                v["id"] = int(st.id) & 0xFFFFFFFFFFFFFFFF
                _put_fill(v["fill"], st.fill)
                v["px_range"], v["sd_threshold"], v["stroke_weight"] = st.pxRange, st.sdThreshold, st.strokeWeight
            elif n.kind == FigKind.nkBackdropBlur:
                pay[:4].view("<f4")[0] = n.backdropBlur.blur
            elif n.kind == FigKind.nkTransform:
                v = pay[: abi.FIG_TRANSFORM_DTYPE.itemsize].view(abi.FIG_TRANSFORM_DTYPE)[0]
                v["translation"] = tuple(n.transform.translation)
                if n.transform.matrix is not None:
                    v["matrix"] = np.asarray(n.transform.matrix, dtype=np.float32).reshape(16)
                v["use_matrix"] = 1 if n.transform.useMatrix else 0
        node_arrays.append(arr)
        root_arrays.append(np.asarray(lst.rootIds, dtype=np.int32))
    garr = np.zeros(max(len(glyph_rows), 1), dtype=abi.GLYPH_DTYPE)
    for i, g in enumerate(glyph_rows):
        garr[i]["key"] = int(g.key) & 0xFFFFFFFFFFFFFFFF
        garr[i]["pos"] = tuple(g.pos)
        _put_fill(garr[i]["fill"], g.fill)
    oarr = np.zeros(max(len(op_rows), 1), dtype=abi.DRAW_OP_DTYPE)
    pts: List[float] = []
    for i, op in enumerate(op_rows):
        o = oarr[i]
        o["kind"] = int(op.kind)
        o["a"], o["b"], o["center"], o["radius"] = tuple(op.a), tuple(op.b), tuple(op.center), op.radius
        if op.box is not None:
            o["box"] = op.box.tuple()
        o["corners"] = tuple(op.corners)
        o["ellipse_radii"] = tuple(op.ellipseRadii)
        o["start_angle"], o["sweep_angle"] = op.startAngle, op.sweepAngle
        o["first_point"], o["n_points"] = len(pts) // 2, len(op.controls)
        o["steps"] = int(op.steps)
        for p in op.controls:
            pts.extend((float(p[0]), float(p[1])))
    parr = np.asarray(pts if pts else [0.0, 0.0], dtype=np.float32)
    rarr = np.zeros(max(len(rect_rows), 1), dtype=abi.TEXT_RECT_DTYPE)
    for i, (r, color) in enumerate(rect_rows):
        rarr[i]["rect"] = r.tuple()
        if color is not None:
            _put_fill(rarr[i]["fill"], color)
    lists = (abi.FdcRenderList * max(len(node_arrays), 1))()
    for i, (na, ra) in enumerate(zip(node_arrays, root_arrays)):
        lists[i].nodes = na.ctypes.data
        lists[i].n_nodes = len(na)
        lists[i].root_ids = ra.ctypes.data
        lists[i].n_roots = len(ra)
    return PackedScene(node_arrays, root_arrays, garr, oarr, lists, parr, rarr)


def flatten(scene: PackedScene, ui_scale: float = 1.0, pixel_scale: float = 1.0, aa_factor: float = 1.2,
            subpixel_enabled: bool = False, image_keys: Iterable[int] = ()) -> np.ndarray:
    """`fdc_flatten_renders`: the frame body as `fdc_call` records.  Pure host code, needs no GPU."""
    lib = abi.load_library()
    keys = np.asarray(sorted(int(k) & 0xFFFFFFFFFFFFFFFF for k in image_keys), dtype=np.uint64)
    env = abi.FdcFlattenEnv(ui_scale, pixel_scale, aa_factor, 1 if subpixel_enabled else 0,
                            keys.ctypes.data if len(keys) else None, len(keys))
    n = ctypes.c_size_t(0)
    cap = 64 + 8 * scene.n_nodes + 16 * len(scene.ops) + len(scene.text_rects)
    for _ in range(2):
        out = np.zeros(cap, dtype=abi.CALL_DTYPE)
        rc = lib.fdc_flatten_renders(ctypes.byref(scene.scene), ctypes.byref(env), out.ctypes.data, cap, ctypes.byref(n))
        if rc == 0:
            return out[: n.value].copy()
        if rc != 4:  # FDC_ERR_CAPACITY
            raise ValueError(f"fdc_flatten_renders failed (fdc_status {rc}): unsupported drawable or malformed list")
        cap = n.value
    raise RuntimeError("fdc_flatten_renders: capacity negotiation failed")


def render_frame(ctx, scene: PackedScene, frame_size: Sequence[float], ui_scale: float = 1.0, clear_main: bool = True,
                 clear_color=(1.0, 1.0, 1.0, 1.0)) -> None:
    """`fdc_render_frame` on a CudaContext: renderFrame with the DFS done natively."""
    rgba = (ctypes.c_float * 4)(*clear_color)
    ctx._ck(ctx._lib.fdc_render_frame(ctx._h, ctypes.byref(scene.scene), float(ui_scale), float(frame_size[0]), float(frame_size[1]), 1 if clear_main else 0, rgba))
