"""ctypes binding of the CPU oracle (oracle/figdraw_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, `__graft_entry__.smoke()`, and bench.py's `cpu_baseline` / `--impl reference`
legs.  Nothing under figdraw_b200/ imports this module; the product has no CPU path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfigdraw_oracle.so")
_LIB: Optional[ctypes.CDLL] = None

N_MODES = 24


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "figdraw_oracle.c"), os.path.join(_HERE, "glyph_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(p) for p in srcs):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _SO


def _lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        build()
        lib = ctypes.CDLL(_SO)
        c = ctypes
        lib.orc_create.restype = c.c_void_p
        lib.orc_create.argtypes = [c.c_int]
        lib.orc_destroy.argtypes = [c.c_void_p]
        lib.orc_set_pixelate.argtypes = [c.c_void_p, c.c_int]
        lib.orc_atlas_size.argtypes = [c.c_void_p]
        lib.orc_rebuilds.argtypes = [c.c_void_p]
        lib.orc_get_image_rect.argtypes = [c.c_void_p, c.c_uint64, c.POINTER(c.c_float)]
        lib.orc_put_image.argtypes = [c.c_void_p, c.c_uint64, c.c_int, c.c_int, c.c_void_p, c.POINTER(c.c_float)]
        lib.orc_render.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.POINTER(c.c_float), c.c_void_p,
                                   c.c_int64, c.c_void_p, c.c_void_p, c.c_int]
        lib.orc_render_rows.argtypes = lib.orc_render.argtypes + [c.c_int, c.c_int]
        lib.orc_max_threads.restype = c.c_int
        lib.orc_count_fragments.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_int64, c.c_void_p, c.c_int]
        lib.orc_rasterize_glyph.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p]
        lib.orc_collect_quads.restype = c.c_int64
        lib.orc_collect_quads.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_int64,
                                          c.c_void_p, c.c_int64]
        _LIB = lib
    return _LIB


def max_threads() -> int:
    """Host threads the oracle may use.  Launchers such as torchrun export OMP_NUM_THREADS=1; the oracle passes an
    explicit num_threads() clause, so the core count is what matters."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


class Oracle:
    """One GL-context-equivalent: an atlas plus the frame interpreter."""

    def __init__(self, atlas_size: int = 1024, pixelate: bool = False):
        self._h = _lib().orc_create(int(atlas_size))
        if pixelate:
            _lib().orc_set_pixelate(self._h, 1)

    def close(self):
        if self._h:
            _lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def atlas_size(self) -> int:
        return int(_lib().orc_atlas_size(self._h))

    def put_image(self, key: int, rgba: np.ndarray) -> Tuple[Tuple[float, float, float, float], bool]:
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w = rgba.shape[:2]
        out = (ctypes.c_float * 4)()
        rebuilt = _lib().orc_put_image(self._h, ctypes.c_uint64(key & (2**64 - 1)), w, h, rgba.ctypes.data, out)
        return tuple(out), bool(rebuilt)

    def get_image_rect(self, key: int):
        out = (ctypes.c_float * 4)()
        ok = _lib().orc_get_image_rect(self._h, ctypes.c_uint64(key & (2**64 - 1)), out)
        return tuple(out) if ok else None

    def render(self, width: int, height: int, calls: np.ndarray, clear=(1.0, 1.0, 1.0, 1.0),
               fb: Optional[np.ndarray] = None, n_threads: int = 0, want_counts: bool = False,
               rows: Optional[Tuple[int, int]] = None):
        """Returns RGBA8 [H, W, 4] (top-left origin) and optionally per-SdfMode fragment counts.
        `rows=(r0, r1)` renders only that row range (bounded CPU-timing sample)."""
        calls = np.ascontiguousarray(calls)
        assert calls.dtype.itemsize == 128
        if fb is None:
            fb = np.zeros((height, width, 4), dtype=np.uint8)
        else:
            fb = np.ascontiguousarray(fb, dtype=np.uint8).copy()
            assert fb.shape == (height, width, 4)
        c4 = (ctypes.c_float * 4)(*(clear if clear is not None else (0, 0, 0, 0)))
        counts = np.zeros(2 * N_MODES, dtype=np.int64)  # [mode] solid, [N_MODES + mode] gradient fills
        nt = n_threads if n_threads > 0 else max_threads()
        r0, r1 = rows if rows is not None else (0, height)
        rc = _lib().orc_render_rows(self._h, width, height, 1 if clear is not None else 0, c4, calls.ctypes.data,
                                    len(calls), fb.ctypes.data, counts.ctypes.data, nt, int(r0), int(r1))
        if rc != 0:
            raise RuntimeError(f"oracle: render failed with status {rc}")
        return (fb, counts) if want_counts else fb


def count_fragments(trace, n_threads: int = 0, oracle: Optional["Oracle"] = None) -> np.ndarray:
    """Exact fragments per SdfMode ([mode] solid fills, [N_MODES+mode] gradient fills) without shading."""
    o = oracle or Oracle(trace.atlas_size)
    if oracle is None:
        for _idx, key, img in trace.images:
            o.put_image(key, img)
    calls = np.ascontiguousarray(trace.calls)
    counts = np.zeros(2 * N_MODES, dtype=np.int64)
    nt = n_threads if n_threads > 0 else max_threads()
    rc = _lib().orc_count_fragments(o._h, trace.width, trace.height, calls.ctypes.data, len(calls), counts.ctypes.data, nt)
    if rc != 0:
        raise RuntimeError(f"oracle: count failed with status {rc}")
    return counts


def reference_bins(trace, tile_w: int = 16, tile_h: int = 16, band: Optional[Tuple[int, int]] = None,
                   oracle: Optional["Oracle"] = None):
    """Reference bin lists per segment: list of (tile_offsets[tiles+1], entries) with entries = backend-call
    ordinals in paint order -- the same form `fdc_debug_bins` returns."""
    o = oracle or Oracle(trace.atlas_size)
    if oracle is None:
        for _idx, key, img in trace.images:
            o.put_image(key, img)
    W, H = trace.width, trace.height
    y0, y1 = band if band is not None else (0, H)
    calls = np.ascontiguousarray(trace.calls)
    n = _lib().orc_collect_quads(o._h, W, H, y0, y1, calls.ctypes.data, len(calls), None, 0)
    rec = np.zeros((max(n, 1), 8), dtype=np.int32)
    _lib().orc_collect_quads(o._h, W, H, y0, y1, calls.ctypes.data, len(calls), rec.ctypes.data, n)
    rec = rec[:n]
    tx_n, ty_n = (W + tile_w - 1) // tile_w, (H + tile_h - 1) // tile_h
    out = []
    n_seg = max(int(rec[:, 0].max()), 0) + 1 if n else 1  # segment -1: placeholder records (dropped first mask quads)
    for s in range(n_seg):
        r = rec[rec[:, 0] == s]
        tx0, ty0 = r[:, 2] // tile_w, r[:, 3] // tile_h
        tx1, ty1 = (r[:, 4] - 1) // tile_w, (r[:, 5] - 1) // tile_h
        nx, ny = tx1 - tx0 + 1, ty1 - ty0 + 1
        cnt = (nx * ny).astype(np.int64)
        total = int(cnt.sum())
        prim = np.repeat(np.arange(len(r)), cnt)
        start = np.repeat(np.cumsum(cnt) - cnt, cnt)
        local = np.arange(total) - start
        nxr = np.repeat(nx, cnt)
        tiles = (np.repeat(ty0, cnt) + local // nxr) * tx_n + np.repeat(tx0, cnt) + local % nxr
        order = np.argsort(tiles, kind="stable")  # stable: emission order inside each tile
        counts = np.bincount(tiles, minlength=tx_n * ty_n)
        offsets = np.zeros(tx_n * ty_n + 1, dtype=np.uint32)
        offsets[1:] = np.cumsum(counts)
        out.append((offsets, r[prim[order], 1].astype(np.uint32)))
    return out


def render_trace(trace, n_threads: int = 0, want_counts: bool = False, oracle: Optional[Oracle] = None,
                 rows: Optional[Tuple[int, int]] = None, pixelate: bool = False):
    """Render a figdraw_b200.figbackend.Trace: uploads its images (in order), then replays its calls."""
    o = oracle or Oracle(trace.atlas_size, pixelate=pixelate)
    if oracle is None:
        for _idx, key, img in trace.images:
            o.put_image(key, img)
    return o.render(trace.width, trace.height, trace.calls, clear=trace.clear, n_threads=n_threads,
                    want_counts=want_counts, rows=rows)


def rasterize_glyph(segs: np.ndarray, width: int, height: int, lcd_filter: bool = False) -> np.ndarray:
    """CPU oracle of the glyph coverage rasteriser (oracle/glyph_oracle.c): `segs` is an array of 32-byte outline
    segments (x0, y0, x1, y1, cx, cy, kind, pad); returns height x width x 4 straight-alpha RGBA8."""
    segs = np.ascontiguousarray(segs)
    assert segs.dtype.itemsize == 32
    out = np.zeros((height, width, 4), dtype=np.uint8)
    rc = _lib().orc_rasterize_glyph(segs.ctypes.data, len(segs), int(width), int(height), 1 if lcd_filter else 0, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle: glyph rasterisation failed with status {rc}")
    return out
