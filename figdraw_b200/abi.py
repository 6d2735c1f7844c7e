"""ctypes view of include/figdraw_cuda.h: enums, the 128-byte `fdc_call` record and the library loader.

The product path FAILS LOUDLY when the CUDA extension is missing: `load_library()` raises, there is
no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes
import enum
import os
from typing import Optional

import numpy as np

ABI_VERSION = 1


class Status(enum.IntEnum):
    OK = 0
    ERR_INVALID = 1
    ERR_CUDA = 2
    ERR_STATE = 3
    ERR_CAPACITY = 4
    ERR_MISSING_IMAGE = 5
    ERR_RETRY = 6


class SdfMode(enum.IntEnum):
    """figbackend.nim:36-52."""

    sdfModeAtlas = 0
    sdfModeClipAA = 3
    sdfModeDropShadow = 7
    sdfModeDropShadowAA = 8
    sdfModeInsetShadow = 9
    sdfModeInsetShadowAnnular = 10
    sdfModeAnnular = 11
    sdfModeAnnularAA = 12
    sdfModeMsdf = 13
    sdfModeMtsdf = 14
    sdfModeMsdfAnnular = 15
    sdfModeMtsdfAnnular = 16
    sdfModeBackdropBlur = 17
    sdfModeBezierStrokeAA = 18
    sdfModeBezierStrokeButtAA = 19
    sdfModeBezierStrokeSquareAA = 20


class FillKindAbi(enum.IntEnum):
    COLORS4 = 0
    COLOR = 1
    LINEAR2 = 2
    LINEAR3 = 3


class Op(enum.IntEnum):
    NOP = 0
    SAVE_TRANSFORM = 1
    RESTORE_TRANSFORM = 2
    TRANSLATE = 3
    ROTATE = 4
    SCALE = 5
    APPLY_TRANSFORM = 6
    SET_AA = 7
    BEGIN_MASK = 8
    END_MASK = 9
    POP_MASK = 10
    BEGIN_RECT_MASK = 11
    POP_RECT_MASK = 12
    BACKDROP_BLUR = 13
    SET_SUBPIXEL = 14
    ROUNDED_RECT = 32
    IMAGE = 33
    MSDF = 34
    BEZIER = 35
    FILLED_QUAD = 36
    RECT = 37


FIRST_DRAW_OP = 32

CALL_DTYPE = np.dtype([("op", "<u4"), ("u", "<u4", (9,)), ("f", "<f4", (22,))])
assert CALL_DTYPE.itemsize == 128
# compact rounded-rect record (fdc_rect64)
RECT64_DTYPE = np.dtype([("rect", "<f4", (4,)), ("radii", "<f4", (4,)), ("factor", "<f4"), ("spread", "<f4"),
                         ("shape_size", "<f4", (2,)), ("packed", "<u4"), ("c", "<u4", (3,))])
assert RECT64_DTYPE.itemsize == 64


class FdcFill(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_uint32),
        ("axis", ctypes.c_uint32),
        ("c", ctypes.c_uint32 * 4),
        ("mid_pos", ctypes.c_float),
    ]


# ---- scene PODs of the native front-end (include/figdraw_cuda.h, "Scene flattening in native code")
NODE_FILL_DTYPE = np.dtype([("kind", "u1"), ("axis", "u1"), ("mid_pos", "u1"), ("_pad", "u1"), ("c", "<u4", (3,))])
NODE_SHADOW_DTYPE = np.dtype([("style", "<u4"), ("fill", NODE_FILL_DTYPE), ("blur", "<f4"), ("spread", "<f4"), ("x", "<f4"),
                              ("y", "<f4")])
NODE_STROKE_DTYPE = np.dtype([("weight", "<f4"), ("fill", NODE_FILL_DTYPE), ("cap", "u1"), ("join", "u1"), ("_pad", "u1", (2,))])
FIG_PAYLOAD_BYTES = 168
FIG_DTYPE = np.dtype({
    "names": ["kind", "zlevel", "flags", "parent", "child_count", "screen_box", "rotation", "fill", "corners",
              "corner_radii_y", "payload"],
    "formats": ["u1", "i1", "<u2", "<i4", "<i4", ("<f4", (4,)), "<f4", NODE_FILL_DTYPE, ("<f4", (4,)), ("<f4", (4,)),
                ("u1", (FIG_PAYLOAD_BYTES,))],
    "offsets": [0, 1, 2, 4, 8, 12, 28, 32, 48, 64, 80],
    "itemsize": 248,
})
# kind-specific views of `payload` (the C union)
FIG_RECT_DTYPE = np.dtype([("shadows", NODE_SHADOW_DTYPE, (4,)), ("stroke", NODE_STROKE_DTYPE)])
FIG_TEXT_DTYPE = np.dtype([("first_glyph", "<u4"), ("n_glyphs", "<u4"), ("first_rect", "<u4"), ("n_selection", "<u4"),
                           ("n_decoration", "<u4")])
TEXT_RECT_DTYPE = np.dtype([("rect", "<f4", (4,)), ("fill", NODE_FILL_DTYPE)])
FIG_DRAWABLE_DTYPE = np.dtype([("stroke", NODE_STROKE_DTYPE), ("steps", "<i4"), ("aa", "<f4"), ("first_op", "<u4"),
                               ("n_ops", "<u4")])
FIG_IMAGE_DTYPE = np.dtype([("id", "<u8"), ("fill", NODE_FILL_DTYPE)])
FIG_MSDF_DTYPE = np.dtype([("id", "<u8"), ("fill", NODE_FILL_DTYPE), ("px_range", "<f4"), ("sd_threshold", "<f4"),
                           ("stroke_weight", "<f4")])
FIG_BACKDROP_DTYPE = np.dtype([("blur", "<f4")])
FIG_TRANSFORM_DTYPE = np.dtype([("translation", "<f4", (2,)), ("matrix", "<f4", (16,)), ("use_matrix", "<u4")])
GLYPH_DTYPE = np.dtype([("key", "<u8"), ("pos", "<f4", (2,)), ("fill", NODE_FILL_DTYPE)])
DRAW_OP_DTYPE = np.dtype([("kind", "<u4"), ("a", "<f4", (2,)), ("b", "<f4", (2,)), ("center", "<f4", (2,)), ("radius", "<f4"),
                          ("box", "<f4", (4,)), ("corners", "<f4", (4,)), ("ellipse_radii", "<f4", (2,)),
                          ("start_angle", "<f4"), ("sweep_angle", "<f4"), ("first_point", "<u4"), ("n_points", "<u4"),
                          ("steps", "<u4")])
assert NODE_FILL_DTYPE.itemsize == 16 and NODE_SHADOW_DTYPE.itemsize == 36 and NODE_STROKE_DTYPE.itemsize == 24
assert TEXT_RECT_DTYPE.itemsize == 32
assert FIG_RECT_DTYPE.itemsize == FIG_PAYLOAD_BYTES and GLYPH_DTYPE.itemsize == 32 and DRAW_OP_DTYPE.itemsize == 92


class FdcRenderList(ctypes.Structure):
    _fields_ = [("nodes", ctypes.c_void_p), ("n_nodes", ctypes.c_uint32), ("root_ids", ctypes.c_void_p),
                ("n_roots", ctypes.c_uint32)]


class FdcScene(ctypes.Structure):
    _fields_ = [("lists", ctypes.POINTER(FdcRenderList)), ("n_lists", ctypes.c_uint32), ("glyphs", ctypes.c_void_p),
                ("text_rects", ctypes.c_void_p), ("ops", ctypes.c_void_p), ("points", ctypes.c_void_p)]


class FdcFlattenEnv(ctypes.Structure):
    _fields_ = [("ui_scale", ctypes.c_float), ("pixel_scale", ctypes.c_float), ("aa_factor", ctypes.c_float),
                ("subpixel_enabled", ctypes.c_uint32), ("image_keys", ctypes.c_void_p), ("n_image_keys", ctypes.c_size_t)]


class FdcFrameStats(ctypes.Structure):
    _fields_ = [
        ("n_prims", ctypes.c_uint32),
        ("n_segments", ctypes.c_uint32),
        ("tiles_x", ctypes.c_uint32),
        ("tiles_y", ctypes.c_uint32),
        ("tile_w", ctypes.c_uint32),
        ("tile_h", ctypes.c_uint32),
        ("n_tile_entries", ctypes.c_uint64),
        ("n_launches", ctypes.c_uint32),
        ("gpu_ms", ctypes.c_float),
        ("shade_ms", ctypes.c_float),
        ("bin_ms", ctypes.c_float),
        ("blur_ms", ctypes.c_float),
    ]


OUTLINE_SEG_DTYPE = np.dtype([("x0", "<f4"), ("y0", "<f4"), ("x1", "<f4"), ("y1", "<f4"), ("cx", "<f4"), ("cy", "<f4"),
                              ("kind", "<u4"), ("_pad", "<u4")])
GLYPH_JOB_DTYPE = np.dtype([("key", "<u8"), ("first_seg", "<u4"), ("n_segs", "<u4"), ("width", "<i4"), ("height", "<i4")])
assert OUTLINE_SEG_DTYPE.itemsize == 32 and GLYPH_JOB_DTYPE.itemsize == 24


class FdcAtlasUsage(ctypes.Structure):
    """AtlasUsage, figbackend.nim:76-89."""

    _fields_ = [("atlas_size", ctypes.c_int32), ("entry_count", ctypes.c_int32), ("image_count", ctypes.c_int32),
                ("glyph_count", ctypes.c_int32), ("generated_count", ctypes.c_int32), ("unknown_count", ctypes.c_int32),
                ("atlas_area", ctypes.c_int64), ("used_area", ctypes.c_int64), ("packed_area", ctypes.c_int64),
                ("generation", ctypes.c_uint64), ("rebuild_count", ctypes.c_uint64)]


# Every symbol include/figdraw_cuda.h declares (checked by tests/test_abi.py against the header text).
EXPORTS = [
    "fdc_create", "fdc_destroy", "fdc_last_error", "fdc_abi_version",
    "fdc_begin_frame", "fdc_end_frame", "fdc_read_pixels", "fdc_read_pixels_async", "fdc_sync", "fdc_replay_frame",
    "fdc_retry_frame", "fdc_abort_frame", "fdc_debug_limit_lists", "fdc_set_replay_graph",
    "fdc_translate", "fdc_rotate", "fdc_scale", "fdc_apply_transform", "fdc_save_transform",
    "fdc_restore_transform", "fdc_transform_mirrors_y", "fdc_get_transform",
    "fdc_sdf_aa_factor", "fdc_set_sdf_aa_factor", "fdc_set_text_subpixel_positioning_enabled",
    "fdc_set_text_subpixel_shift", "fdc_pixel_scale", "fdc_set_pixelate",
    "fdc_draw_rounded_rect_sdf", "fdc_draw_image", "fdc_draw_msdf_image", "fdc_draw_quadratic_bezier_sdf",
    "fdc_draw_filled_quad", "fdc_draw_rect", "fdc_draw_backdrop_blur",
    "fdc_begin_mask", "fdc_end_mask", "fdc_pop_mask", "fdc_begin_rect_mask", "fdc_pop_rect_mask",
    "fdc_submit_calls", "fdc_submit_draws", "fdc_pack_rect64", "fdc_expand_rect64", "fdc_submit_rects64",
    "fdc_put_image", "fdc_update_image", "fdc_has_image", "fdc_get_image_rect", "fdc_remove_image",
    "fdc_reset_image_atlas", "fdc_atlas_size", "fdc_atlas_packed_area",
    "fdc_mark_entry", "fdc_clear_font_glyphs", "fdc_clear_typeface_glyphs", "fdc_retain_owner", "fdc_release_owner",
    "fdc_get_atlas_usage", "fdc_set_atlas_replay", "fdc_rasterize_glyphs",
    "fdc_bind_framebuffer", "fdc_framebuffer_ptr", "fdc_band_rows", "fdc_get_tile_row_costs", "fdc_set_band_tile_rows", "fdc_stream", "fdc_set_peer_framebuffers", "fdc_bind_shared_framebuffer", "fdc_set_frame_barrier", "fdc_export_framebuffer", "fdc_set_peer_gather", "fdc_reserve_framebuffer", "fdc_framebuffer_ipc_handle", "fdc_open_peer_framebuffer",
    "fdc_get_frame_stats", "fdc_debug_bins", "fdc_debug_shade_stats",
    "fdc_flatten_renders", "fdc_render_frame",
]

_LIB: Optional[ctypes.CDLL] = None


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libfigdraw_cuda.so")


def load_library() -> ctypes.CDLL:
    """Load libfigdraw_cuda.so (built in-tree by `__graft_entry__.build()`); raises if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"figdraw_b200: CUDA extension {path} is not built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc). There is no CPU fallback."
        )
    lib = ctypes.CDLL(path)
    c = ctypes
    P = c.c_void_p
    fp = c.POINTER(c.c_float)
    u32p = c.POINTER(c.c_uint32)

    def sig(name, res, *args):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("fdc_create", c.c_int, c.POINTER(P), c.c_int, c.c_int, c.c_float, c.c_int, c.c_int)
    sig("fdc_destroy", None, P)
    sig("fdc_last_error", c.c_char_p, P)
    sig("fdc_abi_version", c.c_int)
    sig("fdc_begin_frame", c.c_int, P, c.c_int, c.c_int, c.c_int, fp)
    sig("fdc_end_frame", c.c_int, P)
    sig("fdc_read_pixels", c.c_int, P, c.c_int, c.c_int, c.c_int, c.c_int, P)
    sig("fdc_sync", c.c_int, P)
    sig("fdc_replay_frame", c.c_int, P)
    sig("fdc_retry_frame", c.c_int, P)
    sig("fdc_set_replay_graph", c.c_int, P, c.c_int)
    sig("fdc_abort_frame", c.c_int, P)
    sig("fdc_debug_limit_lists", c.c_int, P, c.c_uint32, c.c_uint32)
    sig("fdc_translate", c.c_int, P, c.c_float, c.c_float)
    sig("fdc_rotate", c.c_int, P, c.c_float)
    sig("fdc_scale", c.c_int, P, c.c_float, c.c_float)
    sig("fdc_apply_transform", c.c_int, P, fp)
    sig("fdc_save_transform", c.c_int, P)
    sig("fdc_restore_transform", c.c_int, P)
    sig("fdc_transform_mirrors_y", c.c_int, P)
    sig("fdc_get_transform", c.c_int, P, fp)
    sig("fdc_sdf_aa_factor", c.c_float, P)
    sig("fdc_set_sdf_aa_factor", c.c_int, P, c.c_float)
    sig("fdc_set_text_subpixel_positioning_enabled", c.c_int, P, c.c_int)
    sig("fdc_set_text_subpixel_shift", c.c_int, P, c.c_float)
    sig("fdc_pixel_scale", c.c_float, P)
    sig("fdc_set_pixelate", c.c_int, P, c.c_int)
    sig("fdc_draw_rounded_rect_sdf", c.c_int, P, fp, c.POINTER(FdcFill), fp, fp, c.c_int, c.c_float, c.c_float, fp)
    sig("fdc_draw_image", c.c_int, P, c.c_uint64, fp, u32p, fp, c.c_int)
    sig("fdc_draw_msdf_image", c.c_int, P, c.c_uint64, fp, c.c_uint32, fp, c.c_float, c.c_float, c.c_float,
        c.c_int, c.c_int)
    sig("fdc_draw_quadratic_bezier_sdf", c.c_int, P, fp, c.POINTER(FdcFill), fp, fp, fp, c.c_float, c.c_int)
    sig("fdc_draw_filled_quad", c.c_int, P, fp, u32p)
    sig("fdc_draw_rect", c.c_int, P, fp, c.c_uint32)
    sig("fdc_draw_backdrop_blur", c.c_int, P, fp, fp, fp, c.c_float)
    sig("fdc_begin_mask", c.c_int, P, fp, fp, fp)
    sig("fdc_end_mask", c.c_int, P)
    sig("fdc_pop_mask", c.c_int, P)
    sig("fdc_begin_rect_mask", c.c_int, P, fp, fp, fp)
    sig("fdc_pop_rect_mask", c.c_int, P)
    sig("fdc_submit_calls", c.c_int, P, P, c.c_size_t)
    sig("fdc_submit_draws", c.c_int, P, P, c.c_size_t)
    sig("fdc_put_image", c.c_int, P, c.c_uint64, c.c_int, c.c_int, P, fp, c.POINTER(c.c_int))
    sig("fdc_update_image", c.c_int, P, c.c_uint64, c.c_int, c.c_int, P)
    sig("fdc_has_image", c.c_int, P, c.c_uint64)
    sig("fdc_get_image_rect", c.c_int, P, c.c_uint64, fp)
    sig("fdc_remove_image", c.c_int, P, c.c_uint64)
    sig("fdc_reset_image_atlas", c.c_int, P, c.c_int)
    sig("fdc_atlas_size", c.c_int, P)
    sig("fdc_atlas_packed_area", c.c_int, P)
    sig("fdc_mark_entry", c.c_int, P, c.c_uint64, c.c_int, c.c_uint64, c.c_uint64)
    sig("fdc_clear_font_glyphs", c.c_int, P, c.c_uint64)
    sig("fdc_clear_typeface_glyphs", c.c_int, P, c.c_uint64)
    sig("fdc_retain_owner", c.c_int, P, c.c_int, c.c_uint64, c.c_uint64)
    sig("fdc_release_owner", c.c_int, P, c.c_int, c.c_uint64, c.c_uint64, c.POINTER(c.c_int))
    sig("fdc_get_atlas_usage", c.c_int, P, c.POINTER(FdcAtlasUsage))
    sig("fdc_set_atlas_replay", c.c_int, P, c.c_int)
    sig("fdc_rasterize_glyphs", c.c_int, P, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_int, c.POINTER(c.c_int))
    sig("fdc_bind_framebuffer", c.c_int, P, P)
    sig("fdc_framebuffer_ptr", P, P)
    sig("fdc_band_rows", c.c_int, P, c.POINTER(c.c_int), c.POINTER(c.c_int))
    sig("fdc_get_tile_row_costs", c.c_int, P, c.POINTER(c.c_uint32), c.c_int, c.POINTER(c.c_int))
    sig("fdc_set_band_tile_rows", c.c_int, P, c.POINTER(c.c_int), c.c_int)
    sig("fdc_stream", P, P)
    sig("fdc_set_peer_framebuffers", c.c_int, P, c.POINTER(P), c.c_int)
    sig("fdc_reserve_framebuffer", c.c_int, P, c.c_int, c.c_int)
    sig("fdc_framebuffer_ipc_handle", c.c_int, P, c.POINTER(c.c_uint8))
    sig("fdc_open_peer_framebuffer", c.c_int, P, c.POINTER(c.c_uint8), c.POINTER(P))
    sig("fdc_get_frame_stats", c.c_int, P, c.POINTER(FdcFrameStats))
    sig("fdc_debug_shade_stats", c.c_int, P, c.POINTER(c.c_uint64))
    sig("fdc_pack_rect64", c.c_int, c.c_void_p, c.c_void_p)
    sig("fdc_expand_rect64", None, c.c_void_p, c.c_void_p)
    sig("fdc_submit_rects64", c.c_int, P, c.c_void_p, c.c_size_t)
    sig("fdc_read_pixels_async", c.c_int, P, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p)
    sig("fdc_set_peer_gather", c.c_int, P, c.c_int, c.c_int)
    sig("fdc_set_frame_barrier", c.c_int, P, c.c_int)
    sig("fdc_export_framebuffer", c.c_int, P, c.c_int, c.c_int, c.POINTER(c.c_int), c.POINTER(c.c_size_t))
    sig("fdc_bind_shared_framebuffer", c.c_int, P, P, c.c_size_t, c.POINTER(P), c.c_int, P, c.c_int, c.c_int)
    sig("fdc_flatten_renders", c.c_int, c.POINTER(FdcScene), c.POINTER(FdcFlattenEnv), c.c_void_p, c.c_size_t,
        c.POINTER(c.c_size_t))
    sig("fdc_render_frame", c.c_int, P, c.POINTER(FdcScene), c.c_float, c.c_float, c.c_float, c.c_int, c.POINTER(c.c_float))
    sig("fdc_debug_bins", c.c_int, P, c.c_int, u32p, c.c_size_t, u32p, c.c_size_t, c.POINTER(c.c_size_t),
        c.POINTER(c.c_size_t))
    if lib.fdc_abi_version() != ABI_VERSION:
        raise RuntimeError("figdraw_b200: libfigdraw_cuda.so ABI version mismatch; rebuild")
    _LIB = lib
    return lib
