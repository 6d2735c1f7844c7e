"""Backend interface: host-side mirror of `BackendContext` (src/figdraw/figbackend.nim:185-705).

`BackendContext` is the reference's plugin seam: ~45 `method`s that raise "unavailable" by default.
Three implementations live in this repo:
  - `TraceBackend`  : records every call as a 128-byte `fdc_call` (include/figdraw_cuda.h) -- the exact
                      role of the reference's `RecordingBackend` (tests/ttransform.nim:7-122).
  - `CudaContext`   : figdraw_b200/cuda_context.py, forwards to libfigdraw_cuda.so (the product).
  - tests build small fakes on top of `BackendContext` like the reference's tests do.
Method names, argument meaning and error behaviour follow the reference (ValueError = "unavailable").
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .abi import CALL_DTYPE, FillKindAbi, Op, SdfMode
from .fignodes import Fill, FillGradientAxis, FillKind, StrokeCap, f32

DefaultSdfAaFactor = 1.2  # figbackend.nim:34


@dataclass(frozen=True)
class BackendFill:
    """figbackend.nim:96-107.  kind: 1 bfColor, 2 bfLinear2, 3 bfLinear3; 0 = explicit 4 vertex colours."""

    kind: int
    axis: int = 0
    c: Tuple[int, int, int, int] = (0, 0, 0, 0)
    midPos: float = 0.5


def toBackendFill(fill: Fill) -> BackendFill:
    """figbackend.nim:109-127."""
    if fill.kind == FillKind.flColor:
        return BackendFill(kind=FillKindAbi.COLOR, c=(fill.color, 0, 0, 0))
    if fill.kind == FillKind.flLinear2:
        return BackendFill(kind=FillKindAbi.LINEAR2, axis=int(fill.axis), c=(fill.start, fill.stop, 0, 0))
    mid = min(max(f32(fill.midPos) / f32(255.0), f32(0.01)), f32(0.99))
    return BackendFill(kind=FillKindAbi.LINEAR3, axis=int(fill.axis), c=(fill.start, fill.mid, fill.stop, 0),
                       midPos=float(mid))


def colors4(cs: Sequence[int]) -> BackendFill:
    return BackendFill(kind=FillKindAbi.COLORS4, c=tuple(int(x) for x in cs))


def solid(color: int) -> BackendFill:
    return BackendFill(kind=FillKindAbi.COLOR, c=(int(color), 0, 0, 0))


Radii = Tuple[Sequence[float], Sequence[float]]  # (x[TL,TR,BL,BR], y[TL,TR,BL,BR])  CornerRadii2D[float32]


def circularRadii(r: Sequence[float]) -> Radii:
    return (tuple(r), tuple(r))


ZeroRadii: Radii = ((0.0, 0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 0.0))


class BackendContext:
    """Method table of figbackend.nim:245-705; every method raises ValueError until overridden."""

    def _unavailable(self, name):
        raise ValueError(f"Backend {name} unavailable")

    # frame
    def beginFrame(self, frameSize, clearMain=False, clearMainColor=(1.0, 1.0, 1.0, 1.0)):
        self._unavailable("beginFrame")

    def endFrame(self):
        self._unavailable("endFrame")

    def readPixels(self, frame=(0, 0, 0, 0), readFront=False):
        self._unavailable("readPixels")

    def pixelScale(self) -> float:
        return 1.0

    # AA / text flags
    def sdfAaFactor(self) -> float:
        return DefaultSdfAaFactor

    def setSdfAaFactor(self, aaFactor: float):
        pass

    def textSubpixelPositioningEnabled(self) -> bool:
        return False

    def setTextSubpixelPositioningEnabled(self, enabled: bool):
        pass

    def setTextSubpixelShift(self, shift: float):
        pass

    # atlas
    def hasImage(self, key: int) -> bool:
        return False

    def putImage(self, key: int, image: np.ndarray):
        self._unavailable("putImage")

    def updateImage(self, key: int, image: np.ndarray):
        self._unavailable("updateImage")

    def removeImage(self, key: int):
        self._unavailable("removeImage")

    # draws
    def drawRoundedRectSdf(self, rect, fill: BackendFill, radii: Radii, mode=SdfMode.sdfModeClipAA, factor=4.0,
                           spread=0.0, shapeSize=(0.0, 0.0)):
        self._unavailable("drawRoundedRectSdf")

    def drawImage(self, key: int, pos, colors: Sequence[int], size=(0.0, 0.0), flipY=False):
        self._unavailable("drawImage")

    def drawMsdfImage(self, key, pos, color, size, pxRange, sdThreshold=0.5, strokeWeight=0.0, flipY=False):
        self._unavailable("drawMsdfImage")

    def drawMtsdfImage(self, key, pos, color, size, pxRange, sdThreshold=0.5, strokeWeight=0.0, flipY=False):
        self._unavailable("drawMtsdfImage")

    def drawQuadraticBezierSdf(self, rect, fill: BackendFill, p0, p1, p2, strokeWeight, cap):
        self._unavailable("drawQuadraticBezierSdf")

    def drawFilledQuad(self, verts, colors):
        self._unavailable("drawFilledQuad")

    def drawRect(self, rect, color: int):
        self._unavailable("drawRect")

    def drawBackdropBlur(self, rect, radii: Radii, blurRadius: float):
        self._unavailable("drawBackdropBlur")

    # masks
    def beginMask(self, clipRect, radii: Radii):
        self._unavailable("beginMask")

    def endMask(self):
        self._unavailable("endMask")

    def popMask(self):
        self._unavailable("popMask")

    def beginRectMask(self, maskRect, radii: Radii):
        # figbackend.nim:619-626: default falls back to the texture mask
        self.beginMask(maskRect, radii)
        self.endMask()

    def popRectMask(self):
        self.popMask()

    # transforms
    def translate(self, v):
        self._unavailable("translate")

    def rotate(self, angle: float):
        self._unavailable("rotate")

    def scale(self, s):
        self._unavailable("scale")

    def applyTransform(self, m):
        self._unavailable("applyTransform")

    def saveTransform(self):
        self._unavailable("saveTransform")

    def restoreTransform(self):
        self._unavailable("restoreTransform")


def _rect4(r) -> Tuple[float, float, float, float]:
    if hasattr(r, "tuple"):
        return r.tuple()
    return tuple(float(v) for v in r)


class Trace:
    """A recorded frame: the call array both the oracle and the CUDA backend consume, plus atlas uploads.

    `images[i] = (call_index, key, rgba uint8 [h, w, 4])`: putImage happened before call `call_index`.
    """

    def __init__(self, width: int, height: int, clear: Optional[Tuple[float, float, float, float]]):
        self.width, self.height = int(width), int(height)
        self.clear = clear
        self.calls = np.zeros(0, dtype=CALL_DTYPE)
        self.images: List[Tuple[int, int, np.ndarray]] = []
        self.atlas_size = 1024

    @property
    def n_draws(self) -> int:
        return int((self.calls["op"] >= 32).sum())


class TraceBackend(BackendContext):
    """Records backend calls into `fdc_call` records (cf. RecordingBackend, tests/ttransform.nim:7-122)."""

    def __init__(self, atlasSize: int = 1024, pixelScale: float = 1.0):
        self._cap = 256
        self._buf = np.zeros(self._cap, dtype=CALL_DTYPE)
        self._n = 0
        self._images: List[Tuple[int, int, np.ndarray]] = []
        self._keys: Dict[int, Tuple[int, int]] = {}
        self._aa = DefaultSdfAaFactor
        self._subpixel = False
        self._frame = None
        self._clear = None
        self._pixelScale = pixelScale
        self.atlasSize = atlasSize

    # -- recording helpers
    def _rec(self, op: int):
        if self._n == self._cap:
            self._cap *= 2
            nb = np.zeros(self._cap, dtype=CALL_DTYPE)
            nb[: self._n] = self._buf[: self._n]
            self._buf = nb
        r = self._buf[self._n]
        r["op"] = int(op)
        self._n += 1
        return r

    def extend(self, calls: np.ndarray):
        """Bulk append pre-built records (scene generators vectorise with numpy)."""
        n = len(calls)
        while self._n + n > self._cap:
            self._cap *= 2
        if self._cap != len(self._buf):
            nb = np.zeros(self._cap, dtype=CALL_DTYPE)
            nb[: self._n] = self._buf[: self._n]
            self._buf = nb
        self._buf[self._n : self._n + n] = calls
        self._n += n

    @staticmethod
    def _put_fill(r, fill: BackendFill):
        r["u"][1] = int(fill.kind)
        r["u"][2] = int(fill.axis)
        r["u"][3:7] = [int(c) & 0xFFFFFFFF for c in fill.c]
        r["f"][16] = fill.midPos

    @staticmethod
    def _put_rect_radii(r, rect, radii: Radii):
        r["f"][0:4] = _rect4(rect)
        r["f"][4:8] = radii[0]
        r["f"][8:12] = radii[1]

    def trace(self) -> Trace:
        if self._frame is None:
            raise ValueError("beginFrame was not called")
        t = Trace(self._frame[0], self._frame[1], self._clear)
        t.calls = self._buf[: self._n].copy()
        t.images = list(self._images)
        t.atlas_size = self.atlasSize
        return t

    # -- frame
    def beginFrame(self, frameSize, clearMain=False, clearMainColor=(1.0, 1.0, 1.0, 1.0)):
        self._frame = (int(frameSize[0]), int(frameSize[1]))
        self._clear = tuple(clearMainColor) if clearMain else None

    def endFrame(self):
        pass

    def pixelScale(self):
        return self._pixelScale

    # -- AA
    def sdfAaFactor(self):
        return self._aa

    def setSdfAaFactor(self, aaFactor):
        if self._aa == aaFactor:
            return
        self._aa = aaFactor
        self._rec(Op.SET_AA)["f"][0] = aaFactor

    def textSubpixelPositioningEnabled(self):
        return self._subpixel

    def setTextSubpixelPositioningEnabled(self, enabled):
        self._subpixel = bool(enabled)
        self._rec(Op.SET_SUBPIXEL)["u"][0] = 1 if enabled else 0

    def setTextSubpixelShift(self, shift):
        if not self._subpixel:
            return
        r = self._rec(Op.SET_SUBPIXEL)
        r["u"][0] = 1
        r["f"][0] = shift

    # -- atlas
    def hasImage(self, key):
        return key in self._keys

    def putImage(self, key, image):
        image = np.ascontiguousarray(image, dtype=np.uint8)
        assert image.ndim == 3 and image.shape[2] == 4
        self._keys[key] = (image.shape[1], image.shape[0])
        self._images.append((self._n, int(key), image))

    # -- draws
    def drawRoundedRectSdf(self, rect, fill, radii, mode=SdfMode.sdfModeClipAA, factor=4.0, spread=0.0,
                           shapeSize=(0.0, 0.0)):
        r = self._rec(Op.ROUNDED_RECT)
        self._put_rect_radii(r, rect, radii)
        r["f"][12] = factor
        r["f"][13] = spread
        r["f"][14:16] = shapeSize
        r["u"][0] = int(mode)
        self._put_fill(r, fill)

    def drawImage(self, key, pos, colors, size=(0.0, 0.0), flipY=False):
        r = self._rec(Op.IMAGE)
        r["u"][0] = key & 0xFFFFFFFF
        r["u"][1] = (key >> 32) & 0xFFFFFFFF
        r["u"][3:7] = [int(c) & 0xFFFFFFFF for c in colors]
        r["u"][7] = 1 if flipY else 0
        r["f"][0:2] = pos
        r["f"][2:4] = size

    def _msdf(self, mtsdf, key, pos, color, size, pxRange, sdThreshold, strokeWeight, flipY):
        r = self._rec(Op.MSDF)
        r["u"][0] = key & 0xFFFFFFFF
        r["u"][1] = (key >> 32) & 0xFFFFFFFF
        r["u"][2] = 1 if mtsdf else 0
        r["u"][3] = int(color) & 0xFFFFFFFF
        r["u"][7] = 1 if flipY else 0
        r["f"][0:2] = pos
        r["f"][2:4] = size
        r["f"][4] = pxRange
        r["f"][5] = sdThreshold
        r["f"][6] = strokeWeight

    def drawMsdfImage(self, key, pos, color, size, pxRange, sdThreshold=0.5, strokeWeight=0.0, flipY=False):
        self._msdf(False, key, pos, color, size, pxRange, sdThreshold, strokeWeight, flipY)

    def drawMtsdfImage(self, key, pos, color, size, pxRange, sdThreshold=0.5, strokeWeight=0.0, flipY=False):
        self._msdf(True, key, pos, color, size, pxRange, sdThreshold, strokeWeight, flipY)

    def drawQuadraticBezierSdf(self, rect, fill, p0, p1, p2, strokeWeight, cap):
        r = self._rec(Op.BEZIER)
        r["f"][0:4] = _rect4(rect)
        r["f"][4:6] = p0
        r["f"][6:8] = p1
        r["f"][8:10] = p2
        r["f"][10] = strokeWeight
        r["u"][0] = int(cap)
        self._put_fill(r, fill)

    def drawFilledQuad(self, verts, colors):
        r = self._rec(Op.FILLED_QUAD)
        r["f"][0:8] = [c for v in verts for c in v]
        r["u"][3:7] = [int(c) & 0xFFFFFFFF for c in colors]

    def drawRect(self, rect, color):
        r = self._rec(Op.RECT)
        r["f"][0:4] = _rect4(rect)
        r["u"][3] = int(color) & 0xFFFFFFFF

    def drawBackdropBlur(self, rect, radii, blurRadius):
        r = self._rec(Op.BACKDROP_BLUR)
        self._put_rect_radii(r, rect, radii)
        r["f"][12] = blurRadius

    # -- masks
    def beginMask(self, clipRect, radii):
        self._put_rect_radii(self._rec(Op.BEGIN_MASK), clipRect, radii)

    def endMask(self):
        self._rec(Op.END_MASK)

    def popMask(self):
        self._rec(Op.POP_MASK)

    def beginRectMask(self, maskRect, radii):
        self._put_rect_radii(self._rec(Op.BEGIN_RECT_MASK), maskRect, radii)

    def popRectMask(self):
        self._rec(Op.POP_RECT_MASK)

    # -- transforms
    def translate(self, v):
        self._rec(Op.TRANSLATE)["f"][0:2] = v

    def rotate(self, angle):
        self._rec(Op.ROTATE)["f"][0] = angle

    def scale(self, s):
        if np.isscalar(s):
            s = (s, s)
        self._rec(Op.SCALE)["f"][0:2] = s

    def applyTransform(self, m):
        self._rec(Op.APPLY_TRANSFORM)["f"][0:16] = np.asarray(m, dtype=np.float32).reshape(16)

    def saveTransform(self):
        self._rec(Op.SAVE_TRANSFORM)

    def restoreTransform(self):
        self._rec(Op.RESTORE_TRANSFORM)
