"""Scene model: the INPUT contract of the render path (unchanged API).

Host-side mirror of the reference's scene types so that scenes read like the reference's own tests:
  - `Fig`, `RenderList`, `Renders`, `ZLevel`, `FigIdx`      src/figdraw/fignodes.nim:44-92, :165-177, :393-421
  - `FigKind`, `FigFlags`, `RenderShadow`, `RenderStroke`    src/figdraw/figbasics.nim:12-113
  - `Fill`, `linear()`, `FillGradientAxis`                   src/figdraw/common/filltypes.nim:12-95
  - `figLine`, `figCircle`                                   src/figdraw/figextras.nim:3-49

Only data + the index bookkeeping the front-end walks (`addRoot`, `addChild`, `childIndex`).
Arithmetic that the reference does in float32 is done in float32 here (numpy scalars).
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32


# ----------------------------------------------------------------------------- colours / fills
def rgba(r: int, g: int, b: int, a: int = 255) -> int:
    """chroma `rgba()`: packed straight-alpha RGBA8, r | g<<8 | b<<16 | a<<24."""
    return (int(r) & 255) | ((int(g) & 255) << 8) | ((int(b) & 255) << 16) | ((int(a) & 255) << 24)


def rgba_tuple(c: int) -> Tuple[int, int, int, int]:
    return (c & 255, (c >> 8) & 255, (c >> 16) & 255, (c >> 24) & 255)


class FillGradientAxis(enum.IntEnum):
    fgaX = 0
    fgaY = 1
    fgaDiagTLBR = 2
    fgaDiagBLTR = 3


class FillKind(enum.IntEnum):
    flColor = 0
    flLinear2 = 1
    flLinear3 = 2


@dataclass(frozen=True)
class Fill:
    """filltypes.nim:34-42 (variant object flattened)."""

    kind: FillKind = FillKind.flColor
    color: int = 0  # flColor
    axis: FillGradientAxis = FillGradientAxis.fgaX
    start: int = 0
    mid: int = 0
    stop: int = 0
    midPos: int = 128  # uint8


def fill(color: int) -> Fill:
    return Fill(kind=FillKind.flColor, color=color)


def linear(start: int, *rest, axis: FillGradientAxis = FillGradientAxis.fgaX, midPos: int = 128) -> Fill:
    """`linear(start, stop, axis)` / `linear(start, mid, stop, axis, midPos)` (filltypes.nim:50-60)."""
    if len(rest) == 1:
        return Fill(kind=FillKind.flLinear2, axis=axis, start=start, stop=rest[0])
    if len(rest) == 2:
        return Fill(kind=FillKind.flLinear3, axis=axis, start=start, mid=rest[0], stop=rest[1], midPos=midPos)
    raise TypeError("linear(start, stop) or linear(start, mid, stop)")


def toFill(x) -> Fill:
    return x if isinstance(x, Fill) else fill(int(x))


# ----------------------------------------------------------------------------- enums
class FigKind(enum.IntEnum):
    nkFrame = 0
    nkText = 1
    nkRectangle = 2
    nkDrawable = 3
    nkScrollBar = 4
    nkImage = 5
    nkMsdfImage = 6
    nkMtsdfImage = 7
    nkBackdropBlur = 8
    nkTransform = 9


class FigFlags(enum.IntFlag):
    NfClipContent = 1
    NfDisableRender = 2
    NfRootWindow = 4
    NfInactive = 8
    NfSelectText = 16
    NfInvertY = 32
    NfRectMaskContent = 64
    NfEllipticalCorners = 128


class ShadowStyle(enum.IntEnum):
    NoShadow = 0
    DropShadow = 1
    InnerShadow = 2


class StrokeCap(enum.IntEnum):
    scAuto = 0
    scRound = 1
    scButt = 2
    scSquare = 3


class StrokeJoin(enum.IntEnum):
    sjAuto = 0
    sjRound = 1
    sjBevel = 2
    sjMiter = 3


class DrawableKind(enum.IntEnum):
    dkLine = 0
    dkCircle = 1
    dkRectangle = 2
    dkBezier = 3
    dkArc = 4
    dkEllipse = 5


ShadowCount = 4  # figbasics.nim:12


@dataclass
class Rect:
    x: float = 0.0
    y: float = 0.0
    w: float = 0.0
    h: float = 0.0

    def __post_init__(self):
        self.x, self.y, self.w, self.h = f32(self.x), f32(self.y), f32(self.w), f32(self.h)

    def scaled(self, s) -> "Rect":
        s = f32(s)
        return Rect(self.x * s, self.y * s, self.w * s, self.h * s)

    def tuple(self):
        return (float(self.x), float(self.y), float(self.w), float(self.h))


def rect(x, y, w, h) -> Rect:
    return Rect(x, y, w, h)


@dataclass
class RenderShadow:
    style: ShadowStyle = ShadowStyle.NoShadow
    fill: Fill = field(default_factory=Fill)
    blur: float = 0.0
    spread: float = 0.0
    x: float = 0.0
    y: float = 0.0

    def __post_init__(self):
        self.fill = toFill(self.fill)


@dataclass
class RenderStroke:
    weight: float = 0.0
    fill: Fill = field(default_factory=Fill)
    cap: StrokeCap = StrokeCap.scAuto
    join: StrokeJoin = StrokeJoin.sjAuto

    def __post_init__(self):
        self.fill = toFill(self.fill)


@dataclass
class ImageStyle:
    id: int = 0
    fill: Fill = field(default_factory=lambda: fill(rgba(255, 255, 255, 255)))

    def __post_init__(self):
        self.fill = toFill(self.fill)


@dataclass
class MsdfImageStyle:
    id: int = 0
    fill: Fill = field(default_factory=lambda: fill(rgba(255, 255, 255, 255)))
    pxRange: float = 0.0
    sdThreshold: float = 0.0
    strokeWeight: float = 0.0

    def __post_init__(self):
        self.fill = toFill(self.fill)


@dataclass
class BackdropBlurStyle:
    blur: float = 0.0


@dataclass
class TransformStyle:
    translation: Tuple[float, float] = (0.0, 0.0)
    matrix: Optional[Sequence[float]] = None  # 16 floats, vmath column-major
    useMatrix: bool = False


@dataclass
class DrawableOp:
    """fignodes.nim:21-42 (variant object flattened)."""

    kind: DrawableKind
    a: Tuple[float, float] = (0.0, 0.0)
    b: Tuple[float, float] = (0.0, 0.0)
    center: Tuple[float, float] = (0.0, 0.0)  # circle / ellipse centre, arcCenter
    radius: float = 0.0                        # circle radius, arcRadius
    box: Optional[Rect] = None
    corners: Sequence[int] = (0, 0, 0, 0)
    ellipseRadii: Tuple[float, float] = (0.0, 0.0)
    controls: Sequence[Tuple[float, float]] = ()
    steps: int = 0                             # dkBezier steps / dkArc arcSteps (uint16, 0 = adaptive)
    startAngle: float = 0.0
    sweepAngle: float = 0.0


def drawableLine(a, b) -> DrawableOp:
    return DrawableOp(kind=DrawableKind.dkLine, a=tuple(a), b=tuple(b))


def drawableCircle(center, radius) -> DrawableOp:
    return DrawableOp(kind=DrawableKind.dkCircle, center=tuple(center), radius=radius)


def drawableRect(box: Rect, corners=(0, 0, 0, 0)) -> DrawableOp:
    return DrawableOp(kind=DrawableKind.dkRectangle, box=box, corners=tuple(corners))


def drawableEllipse(center, radii) -> DrawableOp:
    return DrawableOp(kind=DrawableKind.dkEllipse, center=tuple(center), ellipseRadii=tuple(radii))


def drawableBezier(*args, steps: int = 0) -> DrawableOp:
    """`drawableBezier(controls, steps)` / `drawableBezier(p0, p1, p2, steps)` (fignodes.nim:258-270)."""
    controls = args[0] if len(args) == 1 else args
    return DrawableOp(kind=DrawableKind.dkBezier, controls=tuple(tuple(p) for p in controls), steps=int(steps))


def drawableArc(center, radius, startAngle, sweepAngle, steps: int = 0) -> DrawableOp:
    """fignodes.nim:278-307."""
    return DrawableOp(kind=DrawableKind.dkArc, center=tuple(center), radius=radius, startAngle=startAngle,
                      sweepAngle=sweepAngle, steps=int(steps))


@dataclass
class Glyph:
    """The per-glyph data `renderText` consumes (figrender.nim:456-493): position is already
    `glyphLocalPos(pos, descent) + imageOffset.scaled()`; `key` is the atlas key of its bitmap."""

    key: int
    pos: Tuple[float, float]
    fill: Fill = field(default_factory=lambda: fill(rgba(0, 0, 0, 255)))

    def __post_init__(self):
        self.fill = toFill(self.fill)


@dataclass
class Fig:
    """fignodes.nim:53-92.  Corner arrays are (TopLeft, TopRight, BottomLeft, BottomRight)."""

    kind: FigKind = FigKind.nkFrame
    zlevel: int = 0
    parent: int = -1
    flags: FigFlags = FigFlags(0)
    childCount: int = 0
    screenBox: Rect = field(default_factory=Rect)
    rotation: float = 0.0
    fill: Fill = field(default_factory=Fill)
    corners: Sequence[int] = (0, 0, 0, 0)
    cornerRadiiY: Sequence[int] = (0, 0, 0, 0)
    # nkRectangle
    shadows: Sequence[RenderShadow] = ()
    stroke: RenderStroke = field(default_factory=RenderStroke)
    # nkText: what the (upstream) text layout produced -- glyphs, selection rects for NfSelectText (selectionRectsFor),
    # underline / strikethrough rects with their span colour (renderTextDecorations, figrender.nim:370-415)
    glyphs: Sequence[Glyph] = ()
    selectionRects: Sequence[Rect] = ()
    decorations: Sequence[Tuple[Rect, Fill]] = ()
    # nkDrawable
    drawStroke: RenderStroke = field(default_factory=RenderStroke)
    drawSteps: int = 0
    drawAa: float = 0.0
    drawOps: List[DrawableOp] = field(default_factory=list)
    # nkImage / nkMsdfImage / nkMtsdfImage
    image: ImageStyle = field(default_factory=ImageStyle)
    msdfImage: MsdfImageStyle = field(default_factory=MsdfImageStyle)
    mtsdfImage: MsdfImageStyle = field(default_factory=MsdfImageStyle)
    # nkBackdropBlur / nkTransform
    backdropBlur: BackdropBlurStyle = field(default_factory=BackdropBlurStyle)
    transform: TransformStyle = field(default_factory=TransformStyle)

    def __post_init__(self):
        self.fill = toFill(self.fill)
        self.flags = FigFlags(int(self.flags))


FigIdxMax = 32767  # FigIdx = int16 (fignodes.nim:51)


@dataclass
class RenderList:
    """Flat pre-order node array + root indices (fignodes.nim:44-46)."""

    nodes: List[Fig] = field(default_factory=list)
    rootIds: List[int] = field(default_factory=list)

    def addRoot(self, root: Fig) -> int:
        idx = len(self.nodes)
        if idx > FigIdxMax:
            raise OverflowError("RenderList exceeds FigIdx int16 capacity")
        root.parent = -1
        self.nodes.append(root)
        self.rootIds.append(idx)
        return idx

    def addChild(self, parentIdx: int, child: Fig) -> int:
        if not (0 <= parentIdx < len(self.nodes)):
            raise IndexError("bad parent index")
        idx = len(self.nodes)
        if idx > FigIdxMax:
            raise OverflowError("RenderList exceeds FigIdx int16 capacity")
        if self.nodes[parentIdx].childCount >= 32767:
            raise ValueError("RenderList parent childCount overflow")
        self.nodes[parentIdx].childCount += 1
        child.parent = parentIdx
        self.nodes.append(child)
        return idx

    def childIndex(self, current: int) -> Iterator[int]:
        """fignodes.nim:165-177."""
        cnt_wanted = self.nodes[current].childCount
        idx, cnt = current + 1, 0
        while cnt < cnt_wanted:
            if idx >= len(self.nodes):
                break
            if self.nodes[idx].parent == current:
                cnt += 1
                yield idx
            idx += 1

    def __len__(self):
        return len(self.nodes)


class Renders:
    """`Renders = ref object layers: OrderedTable[ZLevel, RenderList]` (fignodes.nim:48-49).
    Iteration order is insertion order -- the front-end does NOT sort (figrender.nim:1951); callers do."""

    def __init__(self):
        self.layers: Dict[int, RenderList] = {}

    def __getitem__(self, lvl: int) -> RenderList:
        if not (-128 <= lvl <= 127):
            raise OverflowError("ZLevel is int8")
        return self.layers.setdefault(lvl, RenderList())

    def setLayer(self, lvl: int, lst: RenderList) -> None:
        self.layers[lvl] = lst

    def sort(self) -> None:
        self.layers = dict(sorted(self.layers.items(), key=lambda kv: kv[0]))

    def pairs(self):
        return self.layers.items()

    def addRoot(self, lvl: int, root: Fig) -> int:
        root.zlevel = lvl
        return self[lvl].addRoot(root)

    def addChild(self, lvl: int, parentIdx: int, child: Fig) -> int:
        child.zlevel = lvl
        return self[lvl].addChild(parentIdx, child)


def newRenders() -> Renders:
    return Renders()


# ----------------------------------------------------------------------------- figextras.nim
def figLine(x1, y1, x2, y2, fillv, weight, zlevel: int = 0) -> Fig:
    """figextras.nim:3-30."""
    ax, ay, bx, by = f32(x1), f32(y1), f32(x2), f32(y2)
    dx, dy = bx - ax, by - ay
    hw = max(f32(0), f32(weight)) / f32(2)
    bounds = Rect(min(ax, bx) - hw, min(ay, by) - hw, abs(dx) + hw * f32(2), abs(dy) + hw * f32(2))
    node = Fig(kind=FigKind.nkDrawable, zlevel=zlevel, screenBox=bounds, fill=toFill(fillv))
    node.drawStroke = RenderStroke(weight=weight, fill=toFill(fillv))
    node.drawOps.append(drawableLine((ax - bounds.x, ay - bounds.y), (bx - bounds.x, by - bounds.y)))
    return node


def figCircle(x, y, fillv, radius, zlevel: int = 0) -> Fig:
    """figextras.nim:32-49."""
    r = max(f32(0), f32(radius))
    d = r * f32(2)
    node = Fig(kind=FigKind.nkDrawable, zlevel=zlevel, fill=toFill(fillv))
    node.screenBox = Rect(f32(x) - r, f32(y) - r, d, d)
    node.drawOps.append(drawableCircle((r, r), r))
    return node
