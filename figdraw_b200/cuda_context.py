"""`CudaContext`: the B200 backend behind figdraw's `BackendContext` interface.

Python twin of the Nim shim in bindings/nim/cuda_context.nim: every `BackendContext` method
(src/figdraw/figbackend.nim:245-705, as overridden by `OpenGlContext` in src/figdraw/opengl/glcontext.nim)
forwards to the C-ABI function of the same meaning in libfigdraw_cuda.so.  No arithmetic happens here.

There is no CPU path: constructing a context without the built extension or without a B200 raises.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np

from . import abi
from .abi import FdcFill, FdcFrameStats, SdfMode, Status
from .figbackend import BackendContext, BackendFill, Radii, Trace, _rect4


class FigDrawError(RuntimeError):
    """common/shared.nim:19 `FigDrawError`; carries the fdc_status code."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[fdc_status {code}] {message}")
        self.code = code


def _f4(v) -> ctypes.Array:
    return (ctypes.c_float * 4)(*[float(x) for x in v])


def _f2(v) -> ctypes.Array:
    return (ctypes.c_float * 2)(*[float(x) for x in v])


def _u4(v) -> ctypes.Array:
    return (ctypes.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in v])


def _fill(fill: BackendFill) -> FdcFill:
    f = FdcFill()
    f.kind = int(fill.kind)
    f.axis = int(fill.axis)
    for i in range(4):
        f.c[i] = int(fill.c[i]) & 0xFFFFFFFF
    f.mid_pos = float(fill.midPos)
    return f


class CudaContext(BackendContext):
    def __init__(self, atlasSize: int = 1024, pixelScale: float = 1.0, device: int = 0, rank: int = 0,
                 nRanks: int = 1, pixelate: bool = False):
        self._lib = abi.load_library()
        self._h = ctypes.c_void_p()
        rc = self._lib.fdc_create(ctypes.byref(self._h), int(device), int(atlasSize), float(pixelScale), int(rank),
                                  int(nRanks))
        if rc != 0:
            msg = self._lib.fdc_last_error(None)
            self._h = None
            raise FigDrawError(rc, msg.decode() if msg else "fdc_create failed")
        self.missing_images = 0
        self._frame: Optional[Tuple[int, int]] = None
        if pixelate:  # newContext(pixelate = true), glcontext.nim:255-282
            self._ck(self._lib.fdc_set_pixelate(self._h, 1))

    # -- plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._lib.fdc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            msg = self._lib.fdc_last_error(self._h)
            raise FigDrawError(rc, msg.decode() if msg else "")

    # -- frame
    def beginFrame(self, frameSize, clearMain=False, clearMainColor=(1.0, 1.0, 1.0, 1.0)):
        w, h = int(frameSize[0]), int(frameSize[1])
        self._frame = (w, h)
        self._ck(self._lib.fdc_begin_frame(self._h, w, h, 1 if clearMain else 0, _f4(clearMainColor)))

    def endFrame(self):
        self._ck(self._lib.fdc_end_frame(self._h))

    def replayFrame(self):
        self._ck(self._lib.fdc_replay_frame(self._h))

    def setReplayGraph(self, enabled: bool):
        """replayFrame as one CUDA-graph launch (default) or launch by launch (which also times the phases)."""
        self._ck(self._lib.fdc_set_replay_graph(self._h, 1 if enabled else 0))

    def sync(self):
        self._ck(self._lib.fdc_sync(self._h))

    def syncStatus(self) -> int:
        """fdc_sync without raising on FDC_ERR_RETRY (6): under a tile-band partition the host all-reduces this status
        and calls retryFrame() on every rank when any rank reported 6 (bands.resolve_across_ranks does that)."""
        rc = int(self._lib.fdc_sync(self._h))
        if rc not in (0, int(Status.ERR_RETRY)):
            self._ck(rc)
        return rc

    def retryFrame(self):
        self._ck(self._lib.fdc_retry_frame(self._h))

    def abortFrame(self):
        self._ck(self._lib.fdc_abort_frame(self._h))

    def debugLimitLists(self, coarseEntries: int = 0, tileEntries: int = 0):
        self._ck(self._lib.fdc_debug_limit_lists(self._h, int(coarseEntries), int(tileEntries)))

    def readPixels(self, frame=(0, 0, 0, 0), readFront=False, out: Optional[np.ndarray] = None) -> np.ndarray:
        x, y, w, h = (int(v) for v in frame)
        if w <= 0 or h <= 0:
            x, y = 0, 0
            w, h = self._frame
        if out is None:
            out = np.empty((h, w, 4), dtype=np.uint8)
        self._ck(self._lib.fdc_read_pixels(self._h, x, y, w, h, out.ctypes.data))
        return out

    def readPixelsAsync(self, out: np.ndarray, frame=(0, 0, 0, 0)) -> None:
        """Queue the read-back behind the frame and return; `out` (pinned host memory) is valid after `sync()`."""
        x, y, w, h = (int(v) for v in frame)
        self._ck(self._lib.fdc_read_pixels_async(self._h, x, y, w, h, out.ctypes.data))

    def pixelScale(self) -> float:
        return float(self._lib.fdc_pixel_scale(self._h))

    # -- AA / text
    def sdfAaFactor(self) -> float:
        return float(self._lib.fdc_sdf_aa_factor(self._h))

    def setSdfAaFactor(self, aaFactor: float):
        self._ck(self._lib.fdc_set_sdf_aa_factor(self._h, float(aaFactor)))

    def textSubpixelPositioningEnabled(self) -> bool:
        return bool(getattr(self, "_subpixel", False))

    def setTextSubpixelPositioningEnabled(self, enabled: bool):
        self._subpixel = bool(enabled)
        self._ck(self._lib.fdc_set_text_subpixel_positioning_enabled(self._h, 1 if enabled else 0))

    def setTextSubpixelShift(self, shift: float):
        self._ck(self._lib.fdc_set_text_subpixel_shift(self._h, float(shift)))

    # -- atlas
    def atlasSize(self) -> int:
        return int(self._lib.fdc_atlas_size(self._h))

    def atlasPackedArea(self) -> int:
        return int(self._lib.fdc_atlas_packed_area(self._h))

    def hasImage(self, key: int) -> bool:
        return bool(self._lib.fdc_has_image(self._h, ctypes.c_uint64(key & (2**64 - 1))))

    def putImage(self, key: int, image: np.ndarray):
        """Returns (normalised atlas rect, atlas_rebuilt)."""
        image = np.ascontiguousarray(image, dtype=np.uint8)
        h, w = image.shape[:2]
        rect = (ctypes.c_float * 4)()
        rebuilt = ctypes.c_int(0)
        self._ck(self._lib.fdc_put_image(self._h, ctypes.c_uint64(key & (2**64 - 1)), w, h, image.ctypes.data, rect,
                                         ctypes.byref(rebuilt)))
        return tuple(rect), bool(rebuilt.value)

    def updateImage(self, key: int, image: np.ndarray):
        image = np.ascontiguousarray(image, dtype=np.uint8)
        h, w = image.shape[:2]
        self._ck(self._lib.fdc_update_image(self._h, ctypes.c_uint64(key & (2**64 - 1)), w, h, image.ctypes.data))

    def imageRect(self, key: int):
        rect = (ctypes.c_float * 4)()
        self._ck(self._lib.fdc_get_image_rect(self._h, ctypes.c_uint64(key & (2**64 - 1)), rect))
        return tuple(rect)

    def removeImage(self, key: int):
        self._ck(self._lib.fdc_remove_image(self._h, ctypes.c_uint64(key & (2**64 - 1))))

    # -- atlas residency without the reference's Nim tables (figbackend.nim:355-468)
    def markEntry(self, key: int, kind: int, idA: int = 0, idB: int = 0):
        """kind: 1 image (idA = ImageId), 2 glyph (idA = FontId, idB = TypefaceId), 3 generated."""
        self._ck(self._lib.fdc_mark_entry(self._h, ctypes.c_uint64(key & (2**64 - 1)), int(kind), ctypes.c_uint64(idA), ctypes.c_uint64(idB)))

    def clearFontGlyphs(self, fontId: int) -> int:
        return int(self._lib.fdc_clear_font_glyphs(self._h, ctypes.c_uint64(fontId)))

    def clearTypefaceGlyphs(self, typefaceId: int) -> int:
        return int(self._lib.fdc_clear_typeface_glyphs(self._h, ctypes.c_uint64(typefaceId)))

    def retainOwner(self, what: int, id_: int, token: int):
        self._ck(self._lib.fdc_retain_owner(self._h, int(what), ctypes.c_uint64(id_), ctypes.c_uint64(token)))

    def releaseOwner(self, what: int, id_: int, token: int) -> bool:
        last = ctypes.c_int(0)
        self._ck(self._lib.fdc_release_owner(self._h, int(what), ctypes.c_uint64(id_), ctypes.c_uint64(token), ctypes.byref(last)))
        return bool(last.value)

    def atlasUsage(self):
        u = abi.FdcAtlasUsage()
        self._ck(self._lib.fdc_get_atlas_usage(self._h, ctypes.byref(u)))
        return u

    def setAtlasReplay(self, enabled: bool):
        self._ck(self._lib.fdc_set_atlas_replay(self._h, 1 if enabled else 0))

    def rasterizeGlyphs(self, jobs: np.ndarray, segs: np.ndarray, lcdFilter: bool = False) -> bool:
        """Glyph bitmaps from outlines, rasterised on the GPU straight into the atlas (abi.GLYPH_JOB_DTYPE /
        abi.OUTLINE_SEG_DTYPE arrays).  Returns whether the atlas was rebuilt."""
        jobs = np.ascontiguousarray(jobs, dtype=abi.GLYPH_JOB_DTYPE)
        segs = np.ascontiguousarray(segs, dtype=abi.OUTLINE_SEG_DTYPE)
        rebuilt = ctypes.c_int(0)
        self._ck(self._lib.fdc_rasterize_glyphs(self._h, jobs.ctypes.data, len(jobs), segs.ctypes.data, len(segs),
                                                1 if lcdFilter else 0, ctypes.byref(rebuilt)))
        return bool(rebuilt.value)

    def resetImageAtlas(self, minimumSize: int):
        self._ck(self._lib.fdc_reset_image_atlas(self._h, int(minimumSize)))

    # -- draws
    def drawRoundedRectSdf(self, rect, fill, radii, mode=SdfMode.sdfModeClipAA, factor=4.0, spread=0.0,
                           shapeSize=(0.0, 0.0)):
        f = _fill(fill)
        self._ck(self._lib.fdc_draw_rounded_rect_sdf(self._h, _f4(_rect4(rect)), ctypes.byref(f), _f4(radii[0]),
                                                     _f4(radii[1]), int(mode), float(factor), float(spread),
                                                     _f2(shapeSize)))

    def drawImage(self, key, pos, colors, size=(0.0, 0.0), flipY=False):
        rc = self._lib.fdc_draw_image(self._h, ctypes.c_uint64(key & (2**64 - 1)), _f2(pos), _u4(colors), _f2(size),
                                      1 if flipY else 0)
        if rc == Status.ERR_MISSING_IMAGE:  # glcontext.nim:1305-1310: warn and skip
            self.missing_images += 1
            return
        self._ck(rc)

    def _msdf(self, mtsdf, key, pos, color, size, pxRange, sdThreshold, strokeWeight, flipY):
        rc = self._lib.fdc_draw_msdf_image(self._h, ctypes.c_uint64(key & (2**64 - 1)), _f2(pos),
                                           int(color) & 0xFFFFFFFF, _f2(size), float(pxRange), float(sdThreshold),
                                           float(strokeWeight), 1 if flipY else 0, 1 if mtsdf else 0)
        if rc == Status.ERR_MISSING_IMAGE:
            self.missing_images += 1
            return
        self._ck(rc)

    def drawMsdfImage(self, key, pos, color, size, pxRange, sdThreshold=0.5, strokeWeight=0.0, flipY=False):
        self._msdf(False, key, pos, color, size, pxRange, sdThreshold, strokeWeight, flipY)

    def drawMtsdfImage(self, key, pos, color, size, pxRange, sdThreshold=0.5, strokeWeight=0.0, flipY=False):
        self._msdf(True, key, pos, color, size, pxRange, sdThreshold, strokeWeight, flipY)

    def drawQuadraticBezierSdf(self, rect, fill, p0, p1, p2, strokeWeight, cap):
        f = _fill(fill)
        self._ck(self._lib.fdc_draw_quadratic_bezier_sdf(self._h, _f4(_rect4(rect)), ctypes.byref(f), _f2(p0), _f2(p1),
                                                         _f2(p2), float(strokeWeight), int(cap)))

    def drawFilledQuad(self, verts, colors):
        v = (ctypes.c_float * 8)(*[float(c) for p in verts for c in p])
        self._ck(self._lib.fdc_draw_filled_quad(self._h, v, _u4(colors)))

    def drawRect(self, rect, color):
        self._ck(self._lib.fdc_draw_rect(self._h, _f4(_rect4(rect)), int(color) & 0xFFFFFFFF))

    def drawBackdropBlur(self, rect, radii, blurRadius):
        self._ck(self._lib.fdc_draw_backdrop_blur(self._h, _f4(_rect4(rect)), _f4(radii[0]), _f4(radii[1]),
                                                  float(blurRadius)))

    # -- masks
    def beginMask(self, clipRect, radii):
        self._ck(self._lib.fdc_begin_mask(self._h, _f4(_rect4(clipRect)), _f4(radii[0]), _f4(radii[1])))

    def endMask(self):
        self._ck(self._lib.fdc_end_mask(self._h))

    def popMask(self):
        self._ck(self._lib.fdc_pop_mask(self._h))

    def beginRectMask(self, maskRect, radii):
        self._ck(self._lib.fdc_begin_rect_mask(self._h, _f4(_rect4(maskRect)), _f4(radii[0]), _f4(radii[1])))

    def popRectMask(self):
        self._ck(self._lib.fdc_pop_rect_mask(self._h))

    # -- transforms
    def translate(self, v):
        self._ck(self._lib.fdc_translate(self._h, float(v[0]), float(v[1])))

    def rotate(self, angle):
        self._ck(self._lib.fdc_rotate(self._h, float(angle)))

    def scale(self, s):
        if np.isscalar(s):
            s = (s, s)
        self._ck(self._lib.fdc_scale(self._h, float(s[0]), float(s[1])))

    def applyTransform(self, m):
        arr = (ctypes.c_float * 16)(*[float(x) for x in np.asarray(m, dtype=np.float32).reshape(16)])
        self._ck(self._lib.fdc_apply_transform(self._h, arr))

    def saveTransform(self):
        self._ck(self._lib.fdc_save_transform(self._h))

    def restoreTransform(self):
        self._ck(self._lib.fdc_restore_transform(self._h))

    def transformMirrorsY(self) -> bool:
        return bool(self._lib.fdc_transform_mirrors_y(self._h))

    def getTransform(self) -> np.ndarray:
        out = (ctypes.c_float * 16)()
        self._ck(self._lib.fdc_get_transform(self._h, out))
        return np.array(out, dtype=np.float32)

    # -- display list + introspection
    def submitCalls(self, calls: np.ndarray):
        calls = np.ascontiguousarray(calls)
        assert calls.dtype.itemsize == 128
        self._ck(self._lib.fdc_submit_calls(self._h, calls.ctypes.data, len(calls)))

    def submitDraws(self, draws: np.ndarray):
        """`draws` holds only draw records (op >= 32): one run, the host never reads them."""
        assert draws.dtype.itemsize == 128 and draws.flags["C_CONTIGUOUS"]
        self._ck(self._lib.fdc_submit_draws(self._h, draws.ctypes.data, len(draws)))

    def submitPrepared(self, prepared):
        """Replay a `prepare_calls()` result: state ops record by record, runs of draws through fdc_submit_draws,
        runs of compact rounded-rect records through fdc_submit_rects64."""
        calls, runs = prepared
        for run in runs:
            kind, a, b = run[0], run[1], run[2]
            if kind == "rects64":
                r = run[3]
                self._ck(self._lib.fdc_submit_rects64(self._h, r.ctypes.data, len(r)))
            elif kind in ("draws", True):
                self._ck(self._lib.fdc_submit_draws(self._h, calls[a:b].ctypes.data, b - a))
            else:
                self._ck(self._lib.fdc_submit_calls(self._h, calls[a:b].ctypes.data, b - a))

    def submitRects64(self, rects: np.ndarray):
        rects = np.ascontiguousarray(rects, dtype=abi.RECT64_DTYPE)
        self._ck(self._lib.fdc_submit_rects64(self._h, rects.ctypes.data, len(rects)))

    def renderFrameNative(self, scene, frameSize, uiScale: float = 1.0, clearMain: bool = True,
                          clearColor=(1.0, 1.0, 1.0, 1.0)):
        """renderFrame (figrender.nim:1960-2002) with the scene DFS run natively: `scene` is a
        native_scene.PackedScene (POD `fdc_fig` arrays); one FFI call per frame instead of one per backend call."""
        from . import native_scene

        self._frame = (int(float(frameSize[0]) * uiScale), int(float(frameSize[1]) * uiScale))
        native_scene.render_frame(self, scene, frameSize, uiScale, clearMain, clearColor)

    def frameStats(self) -> FdcFrameStats:
        st = FdcFrameStats()
        self._ck(self._lib.fdc_get_frame_stats(self._h, ctypes.byref(st)))
        return st

    def debugBins(self, segment: int = 0):
        """(tile_offsets[tiles+1], entries) -- entries are backend-call ordinals in paint order."""
        n_off, n_ent = ctypes.c_size_t(0), ctypes.c_size_t(0)
        self._ck(self._lib.fdc_debug_bins(self._h, segment, None, 0, None, 0, ctypes.byref(n_off), ctypes.byref(n_ent)))
        off = np.zeros(n_off.value, dtype=np.uint32)
        ent = np.zeros(max(n_ent.value, 1), dtype=np.uint32)
        self._ck(self._lib.fdc_debug_bins(self._h, segment, off.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                          off.size, ent.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ent.size,
                                          ctypes.byref(n_off), ctypes.byref(n_ent)))
        return off, ent[: n_ent.value]

    def shadeStats(self):
        out = (ctypes.c_uint64 * 8)()
        self._ck(self._lib.fdc_debug_shade_stats(self._h, out))
        return {"visits": out[0], "visits_full": out[1], "visits_general": out[2], "list_steps": out[3], "occl_steps": out[4]}

    def bandRows(self) -> Tuple[int, int]:
        y0, y1 = ctypes.c_int(0), ctypes.c_int(0)
        self._ck(self._lib.fdc_band_rows(self._h, ctypes.byref(y0), ctypes.byref(y1)))
        return y0.value, y1.value

    def tileRowCosts(self) -> np.ndarray:
        """Tile entries per 16-px tile row of the last frame (this rank's rows; 0 elsewhere): fdc_get_tile_row_costs."""
        n = ctypes.c_int(0)
        self._ck(self._lib.fdc_get_tile_row_costs(self._h, None, 0, ctypes.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.uint32)
        self._ck(self._lib.fdc_get_tile_row_costs(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), int(out.size), ctypes.byref(n)))
        return out[: n.value]

    def setBandTileRows(self, bounds) -> None:
        """Bands chosen by the host: n_ranks + 1 tile-row boundaries (same on every rank); None / empty = equal bands."""
        b = [int(v) for v in (bounds if bounds is not None else [])]
        arr = (ctypes.c_int * max(len(b), 1))(*b)
        self._ck(self._lib.fdc_set_band_tile_rows(self._h, arr, len(b)))

    def bindFramebuffer(self, device_ptr: Optional[int]):
        self._ck(self._lib.fdc_bind_framebuffer(self._h, ctypes.c_void_p(device_ptr or 0)))

    def framebufferPtr(self) -> int:
        return int(self._lib.fdc_framebuffer_ptr(self._h) or 0)

    def stream(self) -> int:
        return int(self._lib.fdc_stream(self._h) or 0)

    def reserveFramebuffer(self, width: int, rows: int):
        self._ck(self._lib.fdc_reserve_framebuffer(self._h, int(width), int(rows)))

    def framebufferIpcHandle(self) -> bytes:
        buf = (ctypes.c_uint8 * 64)()
        self._ck(self._lib.fdc_framebuffer_ipc_handle(self._h, buf))
        return bytes(buf)

    def openPeerFramebuffer(self, handle: bytes) -> int:
        buf = (ctypes.c_uint8 * 64)(*handle)
        out = ctypes.c_void_p()
        self._ck(self._lib.fdc_open_peer_framebuffer(self._h, buf, ctypes.byref(out)))
        return int(out.value)

    def setPeerGather(self, mode: str = "stores", subBands: int = 4):
        """"stores": the shade kernel writes every pixel to every peer; "copy": copy engines ship finished slices."""
        self._ck(self._lib.fdc_set_peer_gather(self._h, {"stores": 0, "copy": 1}[mode], int(subBands)))

    def bindSharedFramebuffer(self, localPtr: int, nbytes: int, peerPtrs: Sequence[int], multicastPtr: int, width: int,
                              rows: int):
        """A framebuffer every rank can reach (torch symmetric memory): peers' mappings + the NVSwitch multicast mapping.
        The band all-gather is then fused into the shade kernel's copy-out and every frame ends with a flag barrier."""
        arr = (ctypes.c_void_p * len(peerPtrs))(*[ctypes.c_void_p(p) for p in peerPtrs])
        self._ck(self._lib.fdc_bind_shared_framebuffer(self._h, ctypes.c_void_p(localPtr), int(nbytes), arr, len(peerPtrs),
                                                       ctypes.c_void_p(multicastPtr or 0), int(width), int(rows)))

    def exportFramebuffer(self, width: int, rows: int) -> Tuple[int, int]:
        """The framebuffer as an exportable allocation: (POSIX file descriptor, bytes).  A presenter imports the descriptor
        (Vulkan / GL external memory, or CUDA) and reads the rows after sync(); no read-back over PCIe."""
        fd, nbytes = ctypes.c_int(-1), ctypes.c_size_t(0)
        self._ck(self._lib.fdc_export_framebuffer(self._h, int(width), int(rows), ctypes.byref(fd), ctypes.byref(nbytes)))
        return fd.value, nbytes.value

    def setFrameBarrier(self, enabled: bool):
        self._ck(self._lib.fdc_set_frame_barrier(self._h, 1 if enabled else 0))

    def setPeerFramebuffers(self, ptrs: Sequence[int]):
        arr = (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(p) for p in ptrs])
        self._ck(self._lib.fdc_set_peer_framebuffers(self._h, arr, len(ptrs)))


def pack_rects64(calls: np.ndarray):
    """Vectorised fdc_pack_rect64: (mask of representable records, their fdc_rect64 form).  A record is representable
    when it is a rounded rect with radii_x == radii_y, a solid / 2-stop / 3-stop fill and a midPos on the uint8 grid;
    the result expands back to the identical 128 bytes (tests/test_rect64.py checks that through the C helper)."""
    calls = np.ascontiguousarray(calls)
    u, f = calls["u"], calls["f"]
    fb = f.view(np.uint32)
    kind, axis, mode = u[:, 1], u[:, 2], u[:, 0]
    ok = (calls["op"] == int(abi.Op.ROUNDED_RECT)) & (kind >= 1) & (kind <= 3) & (axis <= 3) & (mode <= 255)
    ok &= (fb[:, 4:8] == fb[:, 8:12]).all(axis=1) & (u[:, 6:9] == 0).all(axis=1) & (fb[:, 17:22] == 0).all(axis=1)
    lin3 = kind == 3
    mid = f[:, 16]
    m = np.zeros(len(calls), dtype=np.uint32)
    found = ~lin3 & (fb[:, 16] == np.float32(0.5).view(np.uint32))
    with np.errstate(invalid="ignore"):
        m0 = np.clip(np.rint(np.nan_to_num(mid.astype(np.float64)) * 255.0), 0, 255).astype(np.int64)
    for dm in (0, -1, 1):
        cand = np.clip(m0 + dm, 0, 255)
        back = np.clip(cand.astype(np.float32) / np.float32(255.0), np.float32(0.01), np.float32(0.99))
        hit = lin3 & ~found & (back.view(np.uint32) == fb[:, 16])
        m[hit] = cand[hit].astype(np.uint32)
        found |= hit
    ok &= found
    r = np.zeros(int(ok.sum()), dtype=abi.RECT64_DTYPE)
    sel = calls[ok]
    r["rect"], r["radii"] = sel["f"][:, 0:4], sel["f"][:, 4:8]
    r["factor"], r["spread"], r["shape_size"] = sel["f"][:, 12], sel["f"][:, 13], sel["f"][:, 14:16]
    r["packed"] = sel["u"][:, 0] | (sel["u"][:, 1] << 8) | (sel["u"][:, 2] << 10) | (m[ok] << 16)
    r["c"] = sel["u"][:, 3:6]
    return ok, r


def prepare_calls(calls: np.ndarray, compact: bool = False, min_compact_run: int = 256):
    """Split a call array once into maximal runs of draw records / other records (what a host that emits the calls
    knows anyway), so replaying it needs no per-record inspection.  `compact`: runs of representable rounded rects are
    converted to 64-byte fdc_rect64 records (half the bytes over PCIe); runs: (kind, a, b[, rects64])."""
    calls = np.ascontiguousarray(calls)
    is_draw = calls["op"] >= abi.FIRST_DRAW_OP
    cls = is_draw.astype(np.int8)
    rects = None
    if compact:
        ok, rects = pack_rects64(calls)
        cls = cls + ok.astype(np.int8)  # 0 state, 1 draw, 2 compact draw
        # short compact runs are not worth a separate submission: demote them to plain draws
        edges = np.flatnonzero(np.diff(cls)) + 1
        bounds = [0, *edges.tolist(), len(calls)]
        for a, b in zip(bounds[:-1], bounds[1:]):
            if cls[a] == 2 and b - a < min_compact_run:
                cls[a:b] = 1
    edges = np.flatnonzero(np.diff(cls)) + 1
    bounds = [0, *edges.tolist(), len(calls)]
    runs = []
    pos = np.cumsum(ok) - ok if compact else None  # index of each record in `rects`
    for a, b in zip(bounds[:-1], bounds[1:]):
        if b <= a:
            continue
        if cls[a] == 2:
            runs.append(("rects64", a, b, np.ascontiguousarray(rects[pos[a]:pos[a] + (b - a)])))
        else:
            runs.append(("draws" if cls[a] == 1 else "state", a, b))
    return calls, runs


def prepared_upload_bytes(prepared) -> int:
    """Bytes that cross host->device when a prepared frame is submitted."""
    _calls, runs = prepared
    return sum((64 if r[0] == "rects64" else 128) * (r[2] - r[1]) for r in runs)


def submit_trace(trace: Trace, ctx: CudaContext) -> None:
    """putImage + beginFrame + fdc_submit_calls + endFrame; the frame is left in flight on the context's stream."""
    for _idx, key, img in trace.images:
        ctx.putImage(key, img)
    ctx.beginFrame((trace.width, trace.height), clearMain=trace.clear is not None,
                   clearMainColor=trace.clear or (1.0, 1.0, 1.0, 1.0))
    ctx.submitCalls(trace.calls)
    ctx.endFrame()


def render_trace(trace: Trace, ctx: Optional[CudaContext] = None, device: int = 0) -> np.ndarray:
    """Render a recorded frame through the C ABI: putImage for its images, one fdc_submit_calls, readPixels."""
    own = ctx is None
    ctx = ctx or CudaContext(atlasSize=trace.atlas_size, device=device)
    try:
        for _idx, key, img in trace.images:
            ctx.putImage(key, img)
        ctx.beginFrame((trace.width, trace.height), clearMain=trace.clear is not None,
                       clearMainColor=trace.clear or (1.0, 1.0, 1.0, 1.0))
        ctx.submitCalls(trace.calls)
        ctx.endFrame()
        return ctx.readPixels()
    finally:
        if own:
            ctx.close()
