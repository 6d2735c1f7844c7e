/*
 * figdraw_cuda.h -- C ABI of the B200 (sm_100a) render backend for figdraw.
 *
 * This is the drop-in boundary for ONE path of elcritch/figdraw: the backend half of the
 * `figdraw/figrender` frame entry, i.e. what `src/figdraw/opengl/glcontext.nim` plus
 * `src/figdraw/opengl/glsl/{atlas,atlas_rect_mask,mask,blur}.frag` do today.  Each entry point
 * replaces one `method` of `BackendContext` (`src/figdraw/figbackend.nim:185-705`), as overridden by
 * `OpenGlContext`; the reference file:line each one replaces is cited beside it.  A Nim
 * `CudaContext = ref object of BackendContext` forwards every method to the function of the same
 * meaning through `{.importc.}` (see INTEGRATION.md and bindings/nim/cuda_context.nim).
 *
 * Conventions
 *  - Plain C: pointers, sizes, POD structs.  No torch / C++ types cross this boundary.
 *  - Handle based, no globals: several contexts may coexist (multi-window tests in the reference).
 *  - All calls on one context come from ONE thread (the reference's render thread,
 *    `figrender.nim:1758` `forbids: [AppMainThreadEff]`).
 *  - Every function returning `int` returns FDC_OK (0) or an fdc_status error code;
 *    `fdc_last_error(ctx)` gives the message.  The Nim shim turns non-zero into `FigDrawError`.
 *  - There is NO CPU fallback: without a usable CUDA device `fdc_create` fails with FDC_ERR_CUDA.
 *  - Colours are straight-alpha RGBA8 packed little-endian in a uint32_t: r | g<<8 | b<<16 | a<<24
 *    (the byte order of chroma's `ColorRGBA`).
 *  - Corner arrays are in `DirectionCorners` order: TopLeft, TopRight, BottomLeft, BottomRight
 *    (`figbasics.nim:24-28`).  Vertex colour arrays are BL, BR, TR, TL (`figbackend.nim:162`).
 *  - Matrices are vmath `Mat4`: 16 floats, column-major, m[col*4+row].
 */
#ifndef FIGDRAW_CUDA_H
#define FIGDRAW_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDC_ABI_VERSION 1

typedef struct fdc_ctx fdc_ctx; /* opaque; one per BackendContext */

typedef enum fdc_status {
  FDC_OK = 0,
  FDC_ERR_INVALID = 1,     /* bad argument */
  FDC_ERR_CUDA = 2,        /* CUDA runtime/driver failure, or no device */
  FDC_ERR_STATE = 3,       /* call order violated (asserts in glcontext.nim:1888-1889, :1985-1986) */
  FDC_ERR_CAPACITY = 4,    /* a documented fixed limit was exceeded (mask depth > 15, atlas > 16384) */
  FDC_ERR_MISSING_IMAGE = 5, /* draw of an image key that is not in the atlas (warn + skip in GL) */
  FDC_ERR_RETRY = 6         /* tile-band partition only: this rank's bin lists overflowed and were regrown; the frame
                               must be re-run with fdc_retry_frame on EVERY rank (all-reduce the status) */
} fdc_status;

/* `SdfMode` -- values identical to figbackend.nim:36-52. */
typedef enum fdc_sdf_mode {
  FDC_SDF_ATLAS = 0,
  FDC_SDF_CLIP_AA = 3,
  FDC_SDF_DROP_SHADOW = 7,
  FDC_SDF_DROP_SHADOW_AA = 8,
  FDC_SDF_INSET_SHADOW = 9,
  FDC_SDF_INSET_SHADOW_ANNULAR = 10,
  FDC_SDF_ANNULAR = 11,
  FDC_SDF_ANNULAR_AA = 12,
  FDC_SDF_MSDF = 13,
  FDC_SDF_MTSDF = 14,
  FDC_SDF_MSDF_ANNULAR = 15,
  FDC_SDF_MTSDF_ANNULAR = 16,
  FDC_SDF_BACKDROP_BLUR = 17,
  FDC_SDF_BEZIER_STROKE_AA = 18,
  FDC_SDF_BEZIER_STROKE_BUTT_AA = 19,
  FDC_SDF_BEZIER_STROKE_SQUARE_AA = 20
} fdc_sdf_mode;

/* `FillGradientAxis` (common/filltypes.nim:12-16) and `BackendFillKind` (figbackend.nim:91-94). */
typedef enum fdc_axis { FDC_AXIS_X = 0, FDC_AXIS_Y = 1, FDC_AXIS_DIAG_TLBR = 2, FDC_AXIS_DIAG_BLTR = 3 } fdc_axis;
typedef enum fdc_fill_kind {
  FDC_FILL_COLORS4 = 0, /* explicit vertex colours BL,BR,TR,TL (the `colors: array[4,ColorRGBA]` overload) */
  FDC_FILL_COLOR = 1,   /* bfColor   : c[0] */
  FDC_FILL_LINEAR2 = 2, /* bfLinear2 : c[0]=start c[1]=stop */
  FDC_FILL_LINEAR3 = 3  /* bfLinear3 : c[0]=start c[1]=mid c[2]=stop, mid_pos in [0.01,0.99] */
} fdc_fill_kind;

/* `StrokeCap` (figbasics.nim:62-66). */
typedef enum fdc_cap { FDC_CAP_AUTO = 0, FDC_CAP_ROUND = 1, FDC_CAP_BUTT = 2, FDC_CAP_SQUARE = 3 } fdc_cap;

/* `BackendFill` (figbackend.nim:96-107) flattened. */
typedef struct fdc_fill {
  uint32_t kind;  /* fdc_fill_kind */
  uint32_t axis;  /* fdc_axis */
  uint32_t c[4];  /* packed RGBA8, meaning by kind */
  float mid_pos;  /* lin3MidPos */
} fdc_fill;

/* ---------------------------------------------------------------------------------------------
 * Display-list record.  One 128-byte record per backend call; `fdc_submit_calls` replays an array of
 * them exactly as if the corresponding fdc_* functions had been called one by one.  The same array is
 * what the test oracle consumes, so both sides see identical input bytes.
 * ------------------------------------------------------------------------------------------- */
typedef enum fdc_op {
  FDC_OP_NOP = 0,
  FDC_OP_SAVE_TRANSFORM = 1,
  FDC_OP_RESTORE_TRANSFORM = 2,
  FDC_OP_TRANSLATE = 3,        /* f[0..1] */
  FDC_OP_ROTATE = 4,           /* f[0] radians */
  FDC_OP_SCALE = 5,            /* f[0..1] */
  FDC_OP_APPLY_TRANSFORM = 6,  /* f[0..15] Mat4 */
  FDC_OP_SET_AA = 7,           /* f[0] */
  FDC_OP_BEGIN_MASK = 8,       /* f[0..3] rect, f[4..7] radii.x, f[8..11] radii.y */
  FDC_OP_END_MASK = 9,
  FDC_OP_POP_MASK = 10,
  FDC_OP_BEGIN_RECT_MASK = 11, /* as BEGIN_MASK */
  FDC_OP_POP_RECT_MASK = 12,
  FDC_OP_BACKDROP_BLUR = 13,   /* f[0..3] rect, f[4..11] radii, f[12] blurRadius */
  FDC_OP_SET_SUBPIXEL = 14,    /* u[0] positioning enabled, f[0] shift */
  /* draws */
  FDC_OP_ROUNDED_RECT = 32,    /* f[0..3] rect, f[4..7] radii.x, f[8..11] radii.y, f[12] factor, f[13] spread,
                                  f[14..15] shapeSize, f[16] mid_pos; u[0] SdfMode, u[1] fill kind, u[2] axis,
                                  u[3..6] fill.c[0..3] */
  FDC_OP_IMAGE = 33,           /* u[0..1] key lo/hi, u[3..6] colours BL,BR,TR,TL, u[7] flipY; f[0..1] pos, f[2..3] size */
  FDC_OP_MSDF = 34,            /* u[0..1] key, u[2] 1=MTSDF, u[3] colour, u[7] flipY; f[0..1] pos, f[2..3] size,
                                  f[4] pxRange, f[5] sdThreshold, f[6] strokeWeight */
  FDC_OP_BEZIER = 35,          /* f[0..3] rect, f[4..5] p0, f[6..7] p1, f[8..9] p2, f[10] strokeWeight, f[16] mid_pos;
                                  u[0] StrokeCap, u[1] fill kind, u[2] axis, u[3..6] fill.c */
  FDC_OP_FILLED_QUAD = 36,     /* f[0..7] four vertices, u[3..6] colours */
  FDC_OP_RECT = 37             /* f[0..3] rect, u[3] colour (legacy drawRect) */
} fdc_op;

typedef struct fdc_call {
  uint32_t op;   /* fdc_op */
  uint32_t u[9];
  float f[22];
} fdc_call; /* sizeof == 128 */

/* ---------------------------------------------------------------------------------------------
 * Context lifetime.  Replaces `newContext` (glcontext.nim:255-535).
 *   device     : CUDA ordinal.
 *   atlas_size : initial atlas edge in texels (GL default 1024), atlas margin is 4 (glcontext.nim:257).
 *   pixel_scale: `pixelScale` (glcontext.nim:260); returned by fdc_pixel_scale.
 *   rank/n_ranks: tile-band partition for multi-GPU.  Rank r shades tile rows
 *               [r*ceil(rows/n), min(rows,(r+1)*ceil(rows/n))) of every frame and leaves other
 *               rows untouched; n_ranks = 1 renders the full frame.
 * ------------------------------------------------------------------------------------------- */
int fdc_create(fdc_ctx** out, int device, int atlas_size, float pixel_scale, int rank, int n_ranks);
void fdc_destroy(fdc_ctx* ctx);
const char* fdc_last_error(fdc_ctx* ctx); /* ctx may be NULL: message of the last failed fdc_create */
int fdc_abi_version(void);

/* --- frame: beginFrame glcontext.nim:2080-2092 (+beginFrameProj :1951-1980), endFrame :1982-1989 --- */
/* clear_main = 0 keeps the previous frame's pixels, as GL keeps the back buffer. */
int fdc_begin_frame(fdc_ctx* ctx, int width, int height, int clear_main, const float clear_rgba[4]);
/* Launches the frame's kernels on the context stream; asynchronous. */
int fdc_end_frame(fdc_ctx* ctx);
/* readPixels glcontext.nim:2094-2135: RGBA8, top-left origin, tightly packed; w<=0||h<=0 reads the whole frame.
 * Synchronises the context stream. */
int fdc_read_pixels(fdc_ctx* ctx, int x, int y, int w, int h, uint8_t* out_rgba);
/* Asynchronous variant for a pipelined presenter: queues the device-to-host copy behind the frame on fdc_stream and
 * returns; `out_rgba` (pinned host memory) holds the pixels after the next fdc_sync.  One request per frame. */
int fdc_read_pixels_async(fdc_ctx* ctx, int x, int y, int w, int h, uint8_t* out_rgba);
/* Blocks until all submitted frames are complete. */
int fdc_sync(fdc_ctx* ctx);
/* Re-runs the last completed frame on the data already resident in device memory (no host->device copy).  The first
 * replay captures the frame's launches into a CUDA graph, later ones are a single cudaGraphLaunch;
 * fdc_set_replay_graph(ctx, 0) re-issues the launches one by one instead, which also times the phases for
 * fdc_get_frame_stats (bin_ms / shade_ms / blur_ms are not available from inside a graph, gpu_ms is). */
int fdc_replay_frame(fdc_ctx* ctx);
int fdc_set_replay_graph(fdc_ctx* ctx, int enabled);
/* Tile-band partitions: re-run the last frame after any rank reported FDC_ERR_RETRY (call on every rank, then gather
 * again).  Restores pixels a first attempt already blended (frames without clear_main) before re-running. */
int fdc_retry_frame(fdc_ctx* ctx);
/* Drop a frame that was begun but cannot be ended because a call in between failed (mask nesting beyond the limit,
 * restoreTransform on an empty stack, an unknown record ...).  Resets the recording and the mask / transform stacks;
 * fdc_render_frame does this itself on its error paths.  (GL has no equivalent: its asserts abort the program.) */
int fdc_abort_frame(fdc_ctx* ctx);

/* --- transforms: glcontext.nim:1991-2017 --- */
int fdc_translate(fdc_ctx* ctx, float x, float y);
int fdc_rotate(fdc_ctx* ctx, float angle);
int fdc_scale(fdc_ctx* ctx, float sx, float sy); /* scale(float32) is sx == sy */
int fdc_apply_transform(fdc_ctx* ctx, const float mat4[16]);
int fdc_save_transform(fdc_ctx* ctx);
int fdc_restore_transform(fdc_ctx* ctx);
int fdc_transform_mirrors_y(fdc_ctx* ctx); /* returns 0/1; glcontext.nim:2019-2024 */
int fdc_get_transform(fdc_ctx* ctx, float out_mat4[16]);

/* --- AA factor and text flags: glcontext.nim:1157-1167, :2052-2077 --- */
float fdc_sdf_aa_factor(fdc_ctx* ctx);
int fdc_set_sdf_aa_factor(fdc_ctx* ctx, float aa);
int fdc_set_text_subpixel_positioning_enabled(fdc_ctx* ctx, int enabled);
int fdc_set_text_subpixel_shift(fdc_ctx* ctx, float shift);
float fdc_pixel_scale(fdc_ctx* ctx);
/* `pixelate` argument of newContext (glcontext.nim:255-282): GL_NEAREST magnification of the atlas (:165-168). */
int fdc_set_pixelate(fdc_ctx* ctx, int enabled);

/* --- draws --- */
/* drawRoundedRectSdf, all three overloads (glcontext.nim:1420-1617): `fill->kind` selects which. */
int fdc_draw_rounded_rect_sdf(fdc_ctx* ctx, const float rect[4], const fdc_fill* fill, const float radii_x[4],
                              const float radii_y[4], int mode, float factor, float spread, const float shape_size[2]);
/* drawImage(path, pos, colors, size, flipY) glcontext.nim:1350-1367; size<=0 draws 1:1 texels. */
int fdc_draw_image(fdc_ctx* ctx, uint64_t key, const float pos[2], const uint32_t colors[4], const float size[2],
                   int flip_y);
/* drawMsdfImage / drawMtsdfImage glcontext.nim:1097-1155. */
int fdc_draw_msdf_image(fdc_ctx* ctx, uint64_t key, const float pos[2], uint32_t color, const float size[2],
                        float px_range, float sd_threshold, float stroke_weight, int flip_y, int is_mtsdf);
/* drawQuadraticBezierSdf glcontext.nim:1619-1741. */
int fdc_draw_quadratic_bezier_sdf(fdc_ctx* ctx, const float rect[4], const fdc_fill* fill, const float p0[2],
                                  const float p1[2], const float p2[2], float stroke_weight, int cap);
/* drawFilledQuad glcontext.nim:963-982. */
int fdc_draw_filled_quad(fdc_ctx* ctx, const float verts[8], const uint32_t colors[4]);
/* drawRect glcontext.nim:1402-1418 (legacy). */
int fdc_draw_rect(fdc_ctx* ctx, const float rect[4], uint32_t color);
/* drawBackdropBlur glcontext.nim:1788-1841 (+ runBackdropSeparableBlur :1743-1786). */
int fdc_draw_backdrop_blur(fdc_ctx* ctx, const float rect[4], const float radii_x[4], const float radii_y[4],
                           float blur_radius);

/* --- masks: glcontext.nim:1886-1949 --- */
int fdc_begin_mask(fdc_ctx* ctx, const float rect[4], const float radii_x[4], const float radii_y[4]);
int fdc_end_mask(fdc_ctx* ctx);
int fdc_pop_mask(fdc_ctx* ctx);
int fdc_begin_rect_mask(fdc_ctx* ctx, const float rect[4], const float radii_x[4], const float radii_y[4]);
int fdc_pop_rect_mask(fdc_ctx* ctx);

/* --- display list --- */
/* Replays `n` records in order.  Equivalent to the individual calls.  Runs of >= 2048 consecutive draw records
 * are transferred to the device in one asynchronous copy straight from `calls` (no host staging pass): if that
 * memory is page-locked it must stay unmodified until the next fdc_sync / fdc_read_pixels / fdc_begin_frame. */
int fdc_submit_calls(fdc_ctx* ctx, const fdc_call* calls, size_t n);
/* As fdc_submit_calls for an array the caller KNOWS holds only draw records (op >= FDC_OP_ROUNDED_RECT) issued under
 * the current state: the host does not even read them -- one run, one asynchronous copy; records with any other op
 * are dropped on the device.  Same lifetime rule for page-locked memory. */
int fdc_submit_draws(fdc_ctx* ctx, const fdc_call* draws, size_t n);

/* Compact form of the most common draw -- drawRoundedRectSdf with circular corner radii and a solid / 2-stop / 3-stop
 * fill -- at half the size of fdc_call: what crosses PCIe per frame is mostly these.  A record expands to exactly the
 * fdc_call it was packed from (fdc_expand_rect64 is what the setup kernel does on the device). */
typedef struct fdc_rect64 {
  float rect[4];
  float radii[4];        /* TL, TR, BL, BR; used for both axes */
  float factor, spread;
  float shape_size[2];
  uint32_t packed;       /* bits 0-7 fdc_sdf_mode, 8-9 fdc_fill_kind (1..3), 10-11 fdc_axis, 16-23 lin3 midPos as the uint8 of
                          * the scene's Fill (mid_pos = clamp(u8 / 255, 0.01, 0.99), figbackend.nim:109-127) */
  uint32_t c[3];         /* as fdc_fill.c[0..2] */
} fdc_rect64;            /* 64 bytes */
/* Host helpers.  fdc_pack_rect64 returns 1 and fills *out when `in` is representable (rounded rect, radii_x == radii_y,
 * no 4-colour fill, midPos on the uint8 grid), else 0.  fdc_expand_rect64 is the inverse. */
int fdc_pack_rect64(const fdc_call* in, fdc_rect64* out);
void fdc_expand_rect64(const fdc_rect64* in, fdc_call* out);
/* As fdc_submit_draws for a run of compact records (one run, one asynchronous copy of 64 bytes per draw). */
int fdc_submit_rects64(fdc_ctx* ctx, const fdc_rect64* rects, size_t n);

/* --- atlas: glcontext.nim:536-641, textures.nim:88-119 --- */
/* putImage: packs (skyline, margin 4), uploads straight-alpha RGBA8 texels + a 2x2-box mip chain.
 * Re-putting an existing key allocates a new slot like GL does.  `out_rect` receives the
 * normalised atlas rect (x,y,w,h)/atlasSize that `entries[key]` holds in the reference.
 * If the atlas is full it doubles (`grow`, glcontext.nim:536-539), drops every entry and returns
 * FDC_OK with *out_rebuilt = 1: the caller must then `noteAtlasRebuilt()` (replay images). */
int fdc_put_image(fdc_ctx* ctx, uint64_t key, int w, int h, const uint8_t* rgba, float out_rect[4], int* out_rebuilt);
/* updateImage glcontext.nim:591-605: same size, in place. */
int fdc_update_image(fdc_ctx* ctx, uint64_t key, int w, int h, const uint8_t* rgba);
int fdc_has_image(fdc_ctx* ctx, uint64_t key);
int fdc_get_image_rect(fdc_ctx* ctx, uint64_t key, float out_rect[4]);
int fdc_remove_image(fdc_ctx* ctx, uint64_t key);
int fdc_reset_image_atlas(fdc_ctx* ctx, int minimum_size); /* resetImageAtlas glcontext.nim:634-641 */
int fdc_atlas_size(fdc_ctx* ctx);
int fdc_atlas_packed_area(fdc_ctx* ctx);
/* Atlas residency for hosts that do not keep the reference's Nim tables.  A Nim CudaContext keeps
 * using the base-class procs over its own tables (figbackend.nim:355-468); these are their native equivalents:
 *   fdc_mark_entry            markImageEntry / markGlyphEntry / markGeneratedEntry :359-398 (id_a = ImageId or FontId,
 *                             id_b = TypefaceId)
 *   fdc_clear_font_glyphs, fdc_clear_typeface_glyphs   :416-432 (return the number of entries removed)
 *   fdc_retain_owner, fdc_release_owner                :434-468; releasing the last owner evicts what it kept alive
 *   fdc_get_atlas_usage       atlasUsage :303-333
 *   fdc_set_atlas_replay(1)   when the atlas has to double, re-pack the live entries and carry their texels over on the
 *                             device instead of dropping everything for the host to replay (noteAtlasRebuilt :202-207);
 *                             removed entries are not carried over, which is what reclaims their space. */
typedef enum fdc_entry_kind { FDC_ENTRY_UNKNOWN = 0, FDC_ENTRY_IMAGE = 1, FDC_ENTRY_GLYPH = 2, FDC_ENTRY_GENERATED = 3 } fdc_entry_kind;
typedef enum fdc_owner_kind { FDC_OWNER_IMAGE = 0, FDC_OWNER_FONT = 1 } fdc_owner_kind;
typedef struct fdc_atlas_usage { /* AtlasUsage, figbackend.nim:76-89 */
  int32_t atlas_size, entry_count, image_count, glyph_count, generated_count, unknown_count;
  int64_t atlas_area, used_area, packed_area;
  uint64_t generation, rebuild_count;
} fdc_atlas_usage;
int fdc_mark_entry(fdc_ctx* ctx, uint64_t key, int kind, uint64_t id_a, uint64_t id_b);
int fdc_clear_font_glyphs(fdc_ctx* ctx, uint64_t font_id);
int fdc_clear_typeface_glyphs(fdc_ctx* ctx, uint64_t typeface_id);
int fdc_retain_owner(fdc_ctx* ctx, int what, uint64_t id, uint64_t token);
int fdc_release_owner(fdc_ctx* ctx, int what, uint64_t id, uint64_t token, int* out_last);
int fdc_get_atlas_usage(fdc_ctx* ctx, fdc_atlas_usage* out);
int fdc_set_atlas_replay(fdc_ctx* ctx, int enabled);

/* --- glyph coverage rasterisation on the GPU.  Replaces, for hosts that hand over OUTLINES, the
 * per-glyph CPU rasterisation + putImage of the reference (common/textrasters/pixie_raster.nim:45-95 renderPixieGlyph:
 * typeset one rune, image.fillText, optional applyLcdFilter :12-43, loadGlyphImage -> putImage).  Each job is one glyph
 * bitmap: `width` x `height` texels, its outline as `n_segs` segments starting at `first_seg` in the shared segment
 * array -- lines and quadratic Beziers in the bitmap's pixel space (x right, y down), closed contours, holes wound
 * opposite to the outer contour (TrueType order).  The library packs a slot for `key` (as fdc_put_image does), writes
 * area-coverage texels (255,255,255,alpha) straight into the atlas and builds the mip chain; `lcd_filter` applies the
 * reference's 5-tap filter.  PARITY UNPINNED against pixie's own anti-aliasing (pixie is not vendored); the oracle is
 * the same signed-area accumulation written sequentially (oracle/glyph_oracle.c). */
typedef struct fdc_outline_seg {
  float x0, y0, x1, y1; /* end points */
  float cx, cy;         /* control point (kind 1) */
  uint32_t kind;        /* 0 line, 1 quadratic Bezier */
  uint32_t _pad;
} fdc_outline_seg;      /* 32 bytes */
typedef struct fdc_glyph_job {
  uint64_t key;         /* atlas key, e.g. hash((2344, fontId, glyphId, lcd, variant)), common/fontglyphs.nim:54-59 */
  uint32_t first_seg, n_segs;
  int32_t width, height;
} fdc_glyph_job;        /* 24 bytes */
int fdc_rasterize_glyphs(fdc_ctx* ctx, const fdc_glyph_job* jobs, size_t n_jobs, const fdc_outline_seg* segs, size_t n_segs,
                         int lcd_filter, int* out_rebuilt);

/* --- multi-GPU / zero-copy plumbing (new; no reference equivalent) --- */
/* Render into caller-owned device memory (W*H*4 bytes, row pitch W*4) instead of the context's own
 * framebuffer; NULL restores the internal one.  Lets a host framework all-gather bands in place. */
int fdc_bind_framebuffer(fdc_ctx* ctx, void* device_rgba8);
void* fdc_framebuffer_ptr(fdc_ctx* ctx);
/* Rows [*y0,*y1) this rank owns for the current frame size. */
int fdc_band_rows(fdc_ctx* ctx, int* y0, int* y1);
/* Band balancing: tile entries per 16-px tile row of the last completed frame (this rank's rows; 0 elsewhere -- sum over
 * the ranks for the whole frame), and bands chosen by the host instead of equal ones: n_ranks + 1 tile-row boundaries,
 * bounds[0] = 0, non-decreasing, bounds[n_ranks] = ceil(H / 16); the same on every rank; n_bounds = 0 restores equal
 * bands.  Takes effect for replays of the recorded frame and for later frames of that height (a frame of another height
 * is partitioned into equal bands until boundaries for it are set).  Needs a framebuffer the
 * ranks share or reach (fdc_bind_shared_framebuffer / fdc_set_peer_framebuffers); an all-gather of equal slices does
 * not apply to unequal bands. */
int fdc_get_tile_row_costs(fdc_ctx* ctx, uint32_t* out, int cap, int* n_rows);
int fdc_set_band_tile_rows(fdc_ctx* ctx, const int* bounds, int n_bounds);
/* The CUDA stream (cudaStream_t) frames are launched on. */
void* fdc_stream(fdc_ctx* ctx);
/* Peer framebuffers: when set (n_ranks entries, own entry may be NULL), the shade kernel stores every
 * finished pixel of its band to all peers directly over NVLink, fusing the band all-gather into the kernel.
 * After every rank's frame has completed (any cross-rank barrier on fdc_stream) each framebuffer holds the
 * whole frame. */
int fdc_set_peer_framebuffers(fdc_ctx* ctx, void* const* device_ptrs, int n);
/* --- present without a read-back (replaces readPixels glcontext.nim:2094-2135 for a presenter on the
 * same machine).  The framebuffer becomes an exportable allocation of at least width*rows*4 bytes (CUDA virtual memory
 * management) and *out_fd receives its POSIX file descriptor: a Vulkan (VK_KHR_external_memory_fd, OPAQUE_FD), OpenGL
 * (EXT_memory_object_fd) or CUDA (cuMemImportFromShareableHandle) consumer imports it once and reads RGBA8 rows of
 * pitch width*4 after fdc_sync -- no pixel crosses PCIe.  The context keeps owning the descriptor. */
int fdc_export_framebuffer(fdc_ctx* ctx, int width, int rows, int* out_fd, size_t* out_bytes);
/* Multi-GPU with a framebuffer every rank can reach (the host allocates it: CUDA VMM / torch symmetric memory): this
 * rank's mapping (`bytes` >= width*rows*4 rounded up to 256, + 4096 bytes of cross-rank flags), the n_ranks peers'
 * mappings of THEIR copies (own entry ignored) and, behind an NVSwitch, the multicast mapping of all copies (or NULL).
 * With a multicast mapping the shade kernel's copy-out writes every finished 16-byte chunk ONCE with multimem.st and
 * the switch delivers it to all ranks -- the band all-gather is fused into the kernel; without one it stores the chunk
 * into each peer.  Every frame then ends with a cross-rank flag barrier on fdc_stream: when the stream has drained,
 * every rank's copy holds the whole frame.  Backdrop blur halos are read from the peers' copies. */
int fdc_bind_shared_framebuffer(fdc_ctx* ctx, void* local_ptr, size_t bytes, void* const* peer_ptrs, int n,
                                void* multicast_ptr, int width, int rows);
int fdc_set_frame_barrier(fdc_ctx* ctx, int enabled); /* the end-of-frame flag barrier of a shared framebuffer (default 1) */
/* How the band reaches the peers registered above.  FDC_GATHER_STORES (default): the shade kernel stores every finished
 * pixel to every peer (fused, SM-driven).  FDC_GATHER_COPY: the last segment is shaded in `sub_bands` slices of tile
 * rows and each finished slice is copied to every peer by the copy engines (cudaMemcpyAsync over NVLink) while the
 * next slice is being shaded; fdc_stream waits for the copies, so the same cross-rank barrier completes the frame. */
typedef enum fdc_gather_mode { FDC_GATHER_STORES = 0, FDC_GATHER_COPY = 1 } fdc_gather_mode;
int fdc_set_peer_gather(fdc_ctx* ctx, int mode, int sub_bands);
/* CUDA IPC plumbing for the above between processes (one process per GPU): make sure the context owns a
 * framebuffer of `rows` x width x 4 bytes (rows >= height; a dedicated cudaMalloc), export its 64-byte
 * cudaIpcMemHandle_t, and map a peer's handle into this process. */
int fdc_reserve_framebuffer(fdc_ctx* ctx, int width, int rows);
int fdc_framebuffer_ipc_handle(fdc_ctx* ctx, uint8_t out_handle[64]);
int fdc_open_peer_framebuffer(fdc_ctx* ctx, const uint8_t handle[64], void** out_device_ptr);

/* --- introspection for parity tests and benchmarks --- */
typedef struct fdc_frame_stats {
  uint32_t n_prims;        /* primitives shaded (after host/device early-outs) */
  uint32_t n_segments;     /* 1 + number of backdrop blurs */
  uint32_t tiles_x, tiles_y, tile_w, tile_h;
  uint64_t n_tile_entries; /* total (tile, primitive) pairs in the bin lists */
  uint32_t n_launches;     /* kernels launched for the last frame */
  float gpu_ms;            /* device time of the last frame (CUDA events on the context stream) */
  float shade_ms;          /* device time of the shade kernel(s) of the last frame */
  float bin_ms;            /* setup + binning kernels */
  float blur_ms;
} fdc_frame_stats;
int fdc_get_frame_stats(fdc_ctx* ctx, fdc_frame_stats* out);
/* Bin lists of segment `segment` of the last frame: tile_offsets has tiles_x*tiles_y+1 entries, entry ids
 * are indices into that segment's primitive array in emission order.  Pass NULL to query sizes. */
int fdc_debug_bins(fdc_ctx* ctx, int segment, uint32_t* tile_offsets, size_t offsets_cap, uint32_t* entries,
                   size_t entries_cap, size_t* n_offsets, size_t* n_entries);

/* Test hook: pretend the coarse / tile bin lists hold at most this many entries (0 = real capacity) until the next
 * regrow, to exercise the overflow -> regrow -> re-run path deterministically. */
int fdc_debug_limit_lists(fdc_ctx* ctx, uint32_t coarse_entries, uint32_t tile_entries);

/* Replays the last frame with counters enabled: out[0] primitive visits by shading warps, [1] of which with full
 * coverage, [2] of which on the general path, [3] 32-entry list steps walked, [4] occlusion-scan steps. */
int fdc_debug_shade_stats(fdc_ctx* ctx, uint64_t out[8]);

/* ---------------------------------------------------------------------------------------------
 * Scene flattening in native code.  Replaces the Nim front-end DFS for hosts that hand
 * over the scene itself instead of making ~45 backend calls per node:
 *   renderFrame figrender.nim:1960-2002, renderRoot :1946-1958, render :1756-1839 (stage order),
 *   renderDropShadows :654-689, renderInnerShadows :716-744, renderRoundedShapeScaledCorners :806-873,
 *   renderText :417-497 (glyph loop), renderImage/renderMsdfImage/renderMtsdfImage/renderBackdropBlur :1673-1754,
 *   renderDrawable :1653-1667 (line :946-995, circle :1111-1126, rect :1128-1132, ellipse :1613-1630, Beziers of any
 *   order :1327-1486 incl. adaptive quadratic spans, arcs :1535-1611, endpoint caps :1010-1039, bevel/miter/round
 *   joins :1059-1109), gradientColors :623-647, toBackendFill figbackend.nim:109-127.
 * The node records are PODs mirroring `Fig` (fignodes.nim:53-92); `seq` members (glyphs, drawable ops, Bezier control
 * points) live in side arrays the nodes index.  Layers are rendered in array order (the reference does not sort, figrender.nim:1951).
 */
typedef struct fdc_node_fill { /* Fill, common/filltypes.nim:34-42 */
  uint8_t kind;                /* 0 flColor, 1 flLinear2, 2 flLinear3 */
  uint8_t axis;                /* fdc_axis */
  uint8_t mid_pos;             /* uint8, flLinear3 */
  uint8_t _pad;
  uint32_t c[3];               /* flColor: c[0]; flLinear2: start c[0], stop c[2]; flLinear3: start, mid, stop */
} fdc_node_fill;               /* 16 bytes */

typedef struct fdc_node_shadow { /* RenderShadow, figbasics.nim:78-90 */
  uint32_t style;              /* 0 none, 1 DropShadow, 2 InnerShadow */
  fdc_node_fill fill;
  float blur, spread, x, y;
} fdc_node_shadow;             /* 36 bytes */

typedef struct fdc_node_stroke { /* RenderStroke, figbasics.nim:92-100 */
  float weight;
  fdc_node_fill fill;
  uint8_t cap;                 /* 0 auto, 1 round, 2 butt, 3 square */
  uint8_t join;
  uint8_t _pad[2];
} fdc_node_stroke;             /* 24 bytes */

typedef enum fdc_fig_kind { /* FigKind, figbasics.nim:14-26 */
  FDC_NK_FRAME = 0, FDC_NK_TEXT = 1, FDC_NK_RECTANGLE = 2, FDC_NK_DRAWABLE = 3, FDC_NK_SCROLLBAR = 4, FDC_NK_IMAGE = 5,
  FDC_NK_MSDF_IMAGE = 6, FDC_NK_MTSDF_IMAGE = 7, FDC_NK_BACKDROP_BLUR = 8, FDC_NK_TRANSFORM = 9
} fdc_fig_kind;

typedef enum fdc_fig_flags { /* FigFlags, figbasics.nim:28-38 */
  FDC_NF_CLIP_CONTENT = 1, FDC_NF_DISABLE_RENDER = 2, FDC_NF_ROOT_WINDOW = 4, FDC_NF_INACTIVE = 8, FDC_NF_SELECT_TEXT = 16,
  FDC_NF_INVERT_Y = 32, FDC_NF_RECT_MASK_CONTENT = 64, FDC_NF_ELLIPTICAL_CORNERS = 128
} fdc_fig_flags;

typedef struct fdc_fig {
  uint8_t kind;                /* fdc_fig_kind */
  int8_t zlevel;
  uint16_t flags;              /* fdc_fig_flags */
  int32_t parent;              /* index in the same list, -1 for roots */
  int32_t child_count;
  float screen_box[4];         /* x, y, w, h (before uiScale) */
  float rotation;              /* degrees */
  fdc_node_fill fill;
  float corners[4];            /* TL, TR, BL, BR */
  float corner_radii_y[4];     /* used with FDC_NF_ELLIPTICAL_CORNERS */
  union {                      /* kind-specific payload (the reference's variant object) */
    struct { fdc_node_shadow shadows[4]; fdc_node_stroke stroke; } rect;                       /* nkRectangle */
    struct { uint32_t first_glyph, n_glyphs, first_rect, n_selection, n_decoration; } text;     /* nkText -> fdc_glyph[], fdc_text_rect[] */
    struct { fdc_node_stroke stroke; int32_t steps; float aa; uint32_t first_op, n_ops; } drawable; /* -> fdc_draw_op[] */
    struct { uint64_t id; fdc_node_fill fill; } image;                                         /* nkImage */
    struct { uint64_t id; fdc_node_fill fill; float px_range, sd_threshold, stroke_weight; } msdf; /* nkMsdfImage / nkMtsdfImage */
    struct { float blur; } backdrop;                                                           /* nkBackdropBlur */
    struct { float translation[2]; float matrix[16]; uint32_t use_matrix; } transform;         /* nkTransform */
  } u;
} fdc_fig;                     /* 248 bytes */

typedef struct fdc_glyph {     /* what renderText consumes per glyph (figrender.nim:456-493) */
  uint64_t key;                /* atlas key of the glyph bitmap */
  float pos[2];                /* glyphLocalPos + imageOffset, already scaled (snapped to whole pixels by the flattener when
                                * subpixel positioning is on, the fraction going to setTextSubpixelShift) */
  fdc_node_fill fill;
} fdc_glyph;                   /* 32 bytes */

typedef struct fdc_text_rect { /* what the text layout hands renderText besides glyphs (figrender.nim:355-415, :434-453):
                                * first the node's selection rects (drawn with the node fill when NfSelectText is set), then
                                * its underline / strikethrough rects with their span colour; local text-box coordinates */
  float rect[4];
  fdc_node_fill fill;          /* decorations only */
} fdc_text_rect;               /* 32 bytes */

typedef struct fdc_draw_op {   /* DrawableOp, fignodes.nim:21-42 */
  uint32_t kind;               /* 0 line, 1 circle, 2 rectangle, 3 bezier, 4 arc, 5 ellipse */
  float a[2], b[2];            /* line */
  float center[2];             /* circle, ellipse, arc */
  float radius;                /* circle, arc */
  float box[4];                /* rectangle */
  float corners[4];            /* rectangle */
  float ellipse_radii[2];
  float start_angle, sweep_angle; /* arc, radians */
  uint32_t first_point, n_points; /* bezier: control points in the `points` side array (x, y pairs), any count >= 2 */
  uint32_t steps;              /* bezier `steps` / arc `arcSteps`; 0 = adaptive (or the node's drawSteps) */
} fdc_draw_op;                 /* 92 bytes */

typedef struct fdc_render_list { /* RenderList, fignodes.nim:44-46 */
  const fdc_fig* nodes;
  uint32_t n_nodes;
  const int32_t* root_ids;
  uint32_t n_roots;
} fdc_render_list;

typedef struct fdc_flatten_env {
  float ui_scale;              /* figUiScale(), common/shared.nim:67-71 */
  float pixel_scale;           /* BackendContext.pixelScale */
  float aa_factor;             /* sdfAaFactor at frame start */
  uint32_t subpixel_enabled;   /* textSubpixelPositioningEnabled at frame start */
  const uint64_t* image_keys;  /* sorted ascending: keys for which hasImage() is true */
  size_t n_image_keys;
} fdc_flatten_env;

typedef struct fdc_scene {     /* Renders (fignodes.nim:48-49) + the side arrays the node records index */
  const fdc_render_list* lists;  /* layers in table order */
  uint32_t n_lists;
  const fdc_glyph* glyphs;
  const fdc_text_rect* text_rects;
  const fdc_draw_op* ops;
  const float* points;
} fdc_scene;

/* Pure host function (no context, no GPU): the body of renderFrame between beginFrame and endFrame as `fdc_call`
 * records -- saveTransform, scale(pixelScale), every layer's roots in order, restoreTransform.  Writes at most `cap`
 * records; *n_out is the number needed (FDC_ERR_CAPACITY when cap was too small). */
int fdc_flatten_renders(const fdc_scene* scene, const fdc_flatten_env* env, fdc_call* out, size_t cap, size_t* n_out);

/* renderFrame on a context: beginFrame(frame size * uiScale), the flattened scene, endFrame. */
int fdc_render_frame(fdc_ctx* ctx, const fdc_scene* scene, float ui_scale, float frame_w, float frame_h, int clear_main,
                     const float clear_rgba[4]);

#ifdef __cplusplus
}
#endif
#endif /* FIGDRAW_CUDA_H */
