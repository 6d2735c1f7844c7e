## cuda_context.nim -- the B200 backend as a sibling of `OpenGlContext`, `VulkanContext` and `MetalContext`.
##
## Drop this file at `src/figdraw/cuda/cuda_context.nim` in a figdraw checkout, build with `-d:figdraw.cuda=on`
## (INTEGRATION.md shows the three-line switch in `commons.nim` / `figrender.nim`) and link `libfigdraw_cuda.so`.
##
## It is deliberately thin: every `BackendContext` method (src/figdraw/figbackend.nim:245-705) forwards to the
## C-ABI function of the same meaning in include/figdraw_cuda.h.  The transform stack, `ceil`, radii packing, mode
## encoding and gradient vertex colours that `glcontext.nim` does on the host all happen behind the ABI (on the
## device, in prim_setup_kernel), so the shim keeps no rendering state of its own except the `entries` table that
## the front-end reads directly (figrender.nim:59-60, :477).
##
## NOT COMPILED in the build container (no Nim toolchain there); kept reviewable against the header.

import std/[hashes, tables]
import pkg/chroma
import pkg/pixie
import pkg/vmath
import pkg/bumpy

import ../commons
import ../figbackend as figbackend
import ../common/filltypes
import ../common/shared
import ../fignodes

{.passL: "-lfigdraw_cuda".}

type
  FdcCtx = distinct pointer

  FdcFill {.bycopy.} = object
    kind, axis: uint32
    c: array[4, uint32]
    midPos: float32

  CudaContext* = ref object of figbackend.BackendContext
    h: FdcCtx
    entries*: Table[Hash, Rect] ## image key -> normalised atlas rect, as in OpenGlContext
    atlasEntryMeta: Table[Hash, figbackend.AtlasEntryMeta]
    frameSize: Vec2

const hdr = "figdraw_cuda.h"

proc fdc_create(outCtx: ptr FdcCtx, device, atlasSize: cint, pixelScale: cfloat, rank, nRanks: cint): cint {.importc, header: hdr.}
proc fdc_destroy(ctx: FdcCtx) {.importc, header: hdr.}
proc fdc_last_error(ctx: FdcCtx): cstring {.importc, header: hdr.}
proc fdc_begin_frame(ctx: FdcCtx, w, h, clearMain: cint, clearRgba: ptr cfloat): cint {.importc, header: hdr.}
proc fdc_end_frame(ctx: FdcCtx): cint {.importc, header: hdr.}
proc fdc_read_pixels(ctx: FdcCtx, x, y, w, h: cint, outRgba: ptr uint8): cint {.importc, header: hdr.}
proc fdc_translate(ctx: FdcCtx, x, y: cfloat): cint {.importc, header: hdr.}
proc fdc_rotate(ctx: FdcCtx, angle: cfloat): cint {.importc, header: hdr.}
proc fdc_scale(ctx: FdcCtx, sx, sy: cfloat): cint {.importc, header: hdr.}
proc fdc_apply_transform(ctx: FdcCtx, m: ptr cfloat): cint {.importc, header: hdr.}
proc fdc_save_transform(ctx: FdcCtx): cint {.importc, header: hdr.}
proc fdc_restore_transform(ctx: FdcCtx): cint {.importc, header: hdr.}
proc fdc_transform_mirrors_y(ctx: FdcCtx): cint {.importc, header: hdr.}
proc fdc_sdf_aa_factor(ctx: FdcCtx): cfloat {.importc, header: hdr.}
proc fdc_set_sdf_aa_factor(ctx: FdcCtx, aa: cfloat): cint {.importc, header: hdr.}
proc fdc_set_text_subpixel_positioning_enabled(ctx: FdcCtx, enabled: cint): cint {.importc, header: hdr.}
proc fdc_set_text_subpixel_shift(ctx: FdcCtx, shift: cfloat): cint {.importc, header: hdr.}
proc fdc_pixel_scale(ctx: FdcCtx): cfloat {.importc, header: hdr.}
proc fdc_draw_rounded_rect_sdf(ctx: FdcCtx, rect: ptr cfloat, fill: ptr FdcFill, rx, ry: ptr cfloat, mode: cint,
                               factor, spread: cfloat, shapeSize: ptr cfloat): cint {.importc, header: hdr.}
proc fdc_draw_image(ctx: FdcCtx, key: uint64, pos: ptr cfloat, colors: ptr uint32, size: ptr cfloat,
                    flipY: cint): cint {.importc, header: hdr.}
proc fdc_draw_msdf_image(ctx: FdcCtx, key: uint64, pos: ptr cfloat, color: uint32, size: ptr cfloat,
                         pxRange, sdThreshold, strokeWeight: cfloat, flipY, isMtsdf: cint): cint {.importc, header: hdr.}
proc fdc_draw_quadratic_bezier_sdf(ctx: FdcCtx, rect: ptr cfloat, fill: ptr FdcFill, p0, p1, p2: ptr cfloat,
                                   strokeWeight: cfloat, cap: cint): cint {.importc, header: hdr.}
proc fdc_draw_filled_quad(ctx: FdcCtx, verts: ptr cfloat, colors: ptr uint32): cint {.importc, header: hdr.}
proc fdc_draw_rect(ctx: FdcCtx, rect: ptr cfloat, color: uint32): cint {.importc, header: hdr.}
proc fdc_draw_backdrop_blur(ctx: FdcCtx, rect, rx, ry: ptr cfloat, blurRadius: cfloat): cint {.importc, header: hdr.}
proc fdc_begin_mask(ctx: FdcCtx, rect, rx, ry: ptr cfloat): cint {.importc, header: hdr.}
proc fdc_end_mask(ctx: FdcCtx): cint {.importc, header: hdr.}
proc fdc_pop_mask(ctx: FdcCtx): cint {.importc, header: hdr.}
proc fdc_begin_rect_mask(ctx: FdcCtx, rect, rx, ry: ptr cfloat): cint {.importc, header: hdr.}
proc fdc_pop_rect_mask(ctx: FdcCtx): cint {.importc, header: hdr.}
proc fdc_put_image(ctx: FdcCtx, key: uint64, w, h: cint, rgba: ptr uint8, outRect: ptr cfloat,
                   outRebuilt: ptr cint): cint {.importc, header: hdr.}
proc fdc_update_image(ctx: FdcCtx, key: uint64, w, h: cint, rgba: ptr uint8): cint {.importc, header: hdr.}
proc fdc_remove_image(ctx: FdcCtx, key: uint64): cint {.importc, header: hdr.}
proc fdc_reset_image_atlas(ctx: FdcCtx, minimumSize: cint): cint {.importc, header: hdr.}
proc fdc_atlas_size(ctx: FdcCtx): cint {.importc, header: hdr.}
proc fdc_atlas_packed_area(ctx: FdcCtx): cint {.importc, header: hdr.}
# round 2: error recovery, replay, pixelate, atlas residency, glyph rasterisation, present, shared framebuffer
proc fdc_sync(ctx: FdcCtx): cint {.importc, header: hdr.}
proc fdc_abort_frame(ctx: FdcCtx): cint {.importc, header: hdr.}
proc fdc_retry_frame(ctx: FdcCtx): cint {.importc, header: hdr.}
proc fdc_replay_frame(ctx: FdcCtx): cint {.importc, header: hdr.}
proc fdc_set_replay_graph(ctx: FdcCtx, enabled: cint): cint {.importc, header: hdr.}
proc fdc_set_pixelate(ctx: FdcCtx, enabled: cint): cint {.importc, header: hdr.}
proc fdc_mark_entry(ctx: FdcCtx, key: uint64, kind: cint, idA, idB: uint64): cint {.importc, header: hdr.}
proc fdc_clear_font_glyphs(ctx: FdcCtx, fontId: uint64): cint {.importc, header: hdr.}
proc fdc_clear_typeface_glyphs(ctx: FdcCtx, typefaceId: uint64): cint {.importc, header: hdr.}
proc fdc_retain_owner(ctx: FdcCtx, what: cint, id, token: uint64): cint {.importc, header: hdr.}
proc fdc_release_owner(ctx: FdcCtx, what: cint, id, token: uint64, outLast: ptr cint): cint {.importc, header: hdr.}
proc fdc_set_atlas_replay(ctx: FdcCtx, enabled: cint): cint {.importc, header: hdr.}
proc fdc_export_framebuffer(ctx: FdcCtx, width, rows: cint, outFd: ptr cint, outBytes: ptr csize_t): cint {.importc, header: hdr.}
proc fdc_get_tile_row_costs(ctx: FdcCtx, outCosts: ptr uint32, cap: cint, nRows: ptr cint): cint {.importc, header: hdr.}
proc fdc_set_band_tile_rows(ctx: FdcCtx, bounds: ptr cint, nBounds: cint): cint {.importc, header: hdr.}
proc fdc_bind_shared_framebuffer(ctx: FdcCtx, local: pointer, bytes: csize_t, peers: ptr pointer, n: cint,
                                 multicast: pointer, width, rows: cint): cint {.importc, header: hdr.}

type
  FdcOutlineSeg {.bycopy.} = object ## 32 bytes: a line (kind 0) or quadratic Bezier (kind 1) in bitmap pixel space
    x0, y0, cx, cy, x1, y1: cfloat
    kind, pad: uint32
  FdcGlyphJob {.bycopy.} = object   ## 24 bytes
    key: uint64
    firstSeg, nSegs: uint32
    width, height: int32

proc fdc_rasterize_glyphs(ctx: FdcCtx, jobs: ptr FdcGlyphJob, nJobs: csize_t, segs: ptr FdcOutlineSeg, nSegs: csize_t,
                          lcdFilter: cint, outRebuilt: ptr cint): cint {.importc, header: hdr.}

template ck(ctx: CudaContext, call: untyped) =
  ## Non-zero fdc_status -> FigDrawError (common/shared.nim:19), like GL errors surface as exceptions.
  let rc = call
  if rc != 0 and rc != 5: # 5 = FDC_ERR_MISSING_IMAGE: GL only warns (glcontext.nim:1305-1310)
    let msg = $fdc_last_error(ctx.h)
    discard fdc_abort_frame(ctx.h) # drop a half-recorded frame so the next beginFrame is legal (no-op outside a frame)
    raise newException(FigDrawError, "cuda backend: " & msg)

func pack(c: ColorRGBA): uint32 =
  c.r.uint32 or (c.g.uint32 shl 8) or (c.b.uint32 shl 16) or (c.a.uint32 shl 24)

func toFdc(fill: figbackend.BackendFill): FdcFill =
  case fill.kind
  of figbackend.bfColor:
    FdcFill(kind: 1, c: [fill.color.pack, 0, 0, 0], midPos: 0.5)
  of figbackend.bfLinear2:
    FdcFill(kind: 2, axis: fill.lin2Axis.uint32, c: [fill.lin2Start.pack, fill.lin2Stop.pack, 0, 0], midPos: 0.5)
  of figbackend.bfLinear3:
    FdcFill(kind: 3, axis: fill.lin3Axis.uint32,
            c: [fill.lin3Start.pack, fill.lin3Mid.pack, fill.lin3Stop.pack, 0], midPos: fill.lin3MidPos)

template rectArr(r: Rect): array[4, cfloat] = [r.x.cfloat, r.y.cfloat, r.w.cfloat, r.h.cfloat]
template radX(r: CornerRadii2D[float32]): array[4, cfloat] =
  [r.x[dcTopLeft].cfloat, r.x[dcTopRight].cfloat, r.x[dcBottomLeft].cfloat, r.x[dcBottomRight].cfloat]
template radY(r: CornerRadii2D[float32]): array[4, cfloat] =
  [r.y[dcTopLeft].cfloat, r.y[dcTopRight].cfloat, r.y[dcBottomLeft].cfloat, r.y[dcBottomRight].cfloat]

proc newContext*(atlasSize = 1024, pixelScale = 1.0, pixelate = false, device = 0, rank = 0, nRanks = 1): CudaContext =
  ## Mirrors `newContext` of glcontext.nim:255 (`pixelate` = GL_NEAREST magnification of atlas texels).
  ## Raises when there is no B200: there is no CPU fallback.
  result = CudaContext()
  if fdc_create(result.h.addr, device.cint, atlasSize.cint, pixelScale.cfloat, rank.cint, nRanks.cint) != 0:
    raise newException(FigDrawError, "cuda backend: " & $fdc_last_error(FdcCtx(nil)))
  if pixelate: discard fdc_set_pixelate(result.h, 1)
  result.entries = initTable[Hash, Rect]()
  result.atlasEntryMeta = initTable[Hash, figbackend.AtlasEntryMeta]()
  result.ensureImageMessageSubscription()
  result.noteAtlasCreated()

method kind*(ctx: CudaContext): figbackend.RendererBackendKind = figbackend.RendererBackendKind.rbOpenGL # or a new rbCuda
method entriesPtr*(ctx: CudaContext): ptr Table[Hash, Rect] = ctx.entries.addr
method atlasEntryMetaPtr*(ctx: CudaContext): var Table[Hash, figbackend.AtlasEntryMeta] = ctx.atlasEntryMeta
method atlasSize*(ctx: CudaContext): int = fdc_atlas_size(ctx.h).int
method atlasPackedArea*(ctx: CudaContext): int = fdc_atlas_packed_area(ctx.h).int
method pixelScale*(ctx: CudaContext): float32 = fdc_pixel_scale(ctx.h)
method hasImage*(ctx: CudaContext, key: Hash): bool = key in ctx.entries

method putImage*(ctx: CudaContext, path: Hash, image: Image) =
  ## glcontext.nim:581-586.  pixie stores premultiplied RGBX; the ABI takes straight alpha like glTexSubImage2D gets
  ## after the `ColorRGBX -> ColorRGBA` conversion of textures.nim:90-92.
  var data = newSeq[ColorRGBA](image.width * image.height)
  for i in 0 ..< data.len: data[i] = image.data[i].rgba()
  var r: array[4, cfloat]
  var rebuilt: cint
  ctx.ck fdc_put_image(ctx.h, cast[uint64](path), image.width.cint, image.height.cint,
                       cast[ptr uint8](data[0].addr), r[0].addr, rebuilt.addr)
  if rebuilt != 0:
    ctx.entries.clear()
    ctx.atlasEntryMeta.clear()
  ctx.entries[path] = rect(r[0], r[1], r[2], r[3])
  ctx.markGeneratedEntry(path)
  if rebuilt != 0: ctx.noteAtlasRebuilt() # replays every live image, figbackend.nim:202-207

method addImage*(ctx: CudaContext, key: Hash, image: Image) = ctx.putImage(key, image)

method updateImage*(ctx: CudaContext, path: Hash, image: Image) =
  var data = newSeq[ColorRGBA](image.width * image.height)
  for i in 0 ..< data.len: data[i] = image.data[i].rgba()
  ctx.ck fdc_update_image(ctx.h, cast[uint64](path), image.width.cint, image.height.cint, cast[ptr uint8](data[0].addr))

method putImage*(ctx: CudaContext, imgObj: ImgObj) =
  case imgObj.kind
  of PixieImg: ctx.putImage(imgObj.id.Hash, imgObj.pimg)
  of FlippyImg: ctx.putImage(imgObj.id.Hash, imgObj.flippy.mipmaps[0]) # the mip chain is rebuilt on the device
  ctx.markImageEntry(imgObj.id)

method resetImageAtlas*(ctx: CudaContext, minimumSize: int) =
  ctx.ck fdc_reset_image_atlas(ctx.h, minimumSize.cint)
  ctx.entries.clear()
  ctx.atlasEntryMeta.clear()
  ctx.noteAtlasRebuilt()

method clearImageAtlas*(ctx: CudaContext) = ctx.resetImageAtlas(0)

method beginFrame*(ctx: CudaContext, frameSize: Vec2, clearMain = false, clearMainColor: Color = whiteColor) =
  var c = [clearMainColor.r.cfloat, clearMainColor.g.cfloat, clearMainColor.b.cfloat, clearMainColor.a.cfloat]
  ctx.frameSize = frameSize
  ctx.ck fdc_begin_frame(ctx.h, frameSize.x.cint, frameSize.y.cint, clearMain.cint, c[0].addr)

method endFrame*(ctx: CudaContext) = ctx.ck fdc_end_frame(ctx.h)

method readPixels*(ctx: CudaContext, frame: Rect, readFront: bool): Image =
  var (x, y, w, h) = (frame.x.int, frame.y.int, frame.w.int, frame.h.int)
  if w <= 0 or h <= 0: (x, y, w, h) = (0, 0, ctx.frameSize.x.int, ctx.frameSize.y.int)
  if w <= 0 or h <= 0: return newImage(0, 0)
  result = newImage(w, h)
  var data = newSeq[ColorRGBA](w * h)
  ctx.ck fdc_read_pixels(ctx.h, x.cint, y.cint, w.cint, h.cint, cast[ptr uint8](data[0].addr))
  for i in 0 ..< data.len: result.data[i] = data[i].rgbx() # already top-left origin: no flipVertical needed

method translate*(ctx: CudaContext, v: Vec2) = ctx.ck fdc_translate(ctx.h, v.x, v.y)
method rotate*(ctx: CudaContext, angle: float32) = ctx.ck fdc_rotate(ctx.h, angle)
method scale*(ctx: CudaContext, s: float32) = ctx.ck fdc_scale(ctx.h, s, s)
method scale*(ctx: CudaContext, s: Vec2) = ctx.ck fdc_scale(ctx.h, s.x, s.y)
method applyTransform*(ctx: CudaContext, m: Mat4) =
  var a: array[16, cfloat]
  for i in 0 ..< 4:
    for j in 0 ..< 4: a[i * 4 + j] = m[i, j] # vmath: m[col, row]
  ctx.ck fdc_apply_transform(ctx.h, a[0].addr)
method saveTransform*(ctx: CudaContext) = ctx.ck fdc_save_transform(ctx.h)
method restoreTransform*(ctx: CudaContext) = ctx.ck fdc_restore_transform(ctx.h)
method transformMirrorsY*(ctx: CudaContext): bool = fdc_transform_mirrors_y(ctx.h) != 0

method sdfAaFactor*(ctx: CudaContext): float32 = fdc_sdf_aa_factor(ctx.h)
method setSdfAaFactor*(ctx: CudaContext, aaFactor: float32) = ctx.ck fdc_set_sdf_aa_factor(ctx.h, aaFactor)
method setTextSubpixelPositioningEnabled*(ctx: CudaContext, enabled: bool) =
  ctx.ck fdc_set_text_subpixel_positioning_enabled(ctx.h, enabled.cint)
method setTextSubpixelShift*(ctx: CudaContext, shift: float32) = ctx.ck fdc_set_text_subpixel_shift(ctx.h, shift)

proc drawRR(ctx: CudaContext, rect: Rect, fill: FdcFill, radii: CornerRadii2D[float32], mode: figbackend.SdfMode,
            factor, spread: float32, shapeSize: Vec2) =
  var (r, f, rx, ry) = (rectArr(rect), fill, radX(radii), radY(radii))
  var ss = [shapeSize.x.cfloat, shapeSize.y.cfloat]
  ctx.ck fdc_draw_rounded_rect_sdf(ctx.h, r[0].addr, f.addr, rx[0].addr, ry[0].addr, mode.cint, factor, spread, ss[0].addr)

method drawRoundedRectSdf*(ctx: CudaContext, rect: Rect, colors: array[4, ColorRGBA], radii: CornerRadii2D[float32],
                           mode: figbackend.SdfMode = sdfModeClipAA, factor: float32 = 4.0, spread: float32 = 0.0,
                           shapeSize: Vec2 = vec2(0, 0)) =
  ctx.drawRR(rect, FdcFill(kind: 0, c: [colors[0].pack, colors[1].pack, colors[2].pack, colors[3].pack], midPos: 0.5),
             radii, mode, factor, spread, shapeSize)

method drawRoundedRectSdf*(ctx: CudaContext, rect: Rect, fill: figbackend.BackendFill, radii: CornerRadii2D[float32],
                           mode: figbackend.SdfMode = sdfModeClipAA, factor: float32 = 4.0, spread: float32 = 0.0,
                           shapeSize: Vec2 = vec2(0, 0)) =
  ctx.drawRR(rect, fill.toFdc, radii, mode, factor, spread, shapeSize)

method drawRoundedRectSdf*(ctx: CudaContext, rect: Rect, color: Color, radii: CornerRadii2D[float32],
                           mode: figbackend.SdfMode = sdfModeClipAA, factor: float32 = 4.0, spread: float32 = 0.0,
                           shapeSize: Vec2 = vec2(0, 0)) =
  ctx.drawRR(rect, FdcFill(kind: 1, c: [color.rgba().pack, 0, 0, 0], midPos: 0.5), radii, mode, factor, spread, shapeSize)

method drawImage*(ctx: CudaContext, imageId: Hash, pos: Vec2, colors: array[4, ColorRGBA], size: Vec2, flipY: bool) =
  var (p, s) = ([pos.x.cfloat, pos.y.cfloat], [size.x.cfloat, size.y.cfloat])
  var c = [colors[0].pack, colors[1].pack, colors[2].pack, colors[3].pack]
  ctx.ck fdc_draw_image(ctx.h, cast[uint64](imageId), p[0].addr, c[0].addr, s[0].addr, flipY.cint)

proc drawSdfImage(ctx: CudaContext, imageId: Hash, pos: Vec2, color: Color, size: Vec2, pxRange, sdThreshold,
                  strokeWeight: float32, flipY, mtsdf: bool) =
  var (p, s) = ([pos.x.cfloat, pos.y.cfloat], [size.x.cfloat, size.y.cfloat])
  ctx.ck fdc_draw_msdf_image(ctx.h, cast[uint64](imageId), p[0].addr, color.rgba().pack, s[0].addr, pxRange, sdThreshold,
                             strokeWeight, flipY.cint, mtsdf.cint)

method drawMsdfImage*(ctx: CudaContext, imageId: Hash, pos: Vec2 = vec2(0, 0), color = color(1, 1, 1, 1), size: Vec2,
                      pxRange: float32, sdThreshold: float32 = 0.5, strokeWeight: float32 = 0.0, flipY = false) =
  ctx.drawSdfImage(imageId, pos, color, size, pxRange, sdThreshold, strokeWeight, flipY, false)

method drawMtsdfImage*(ctx: CudaContext, imageId: Hash, pos: Vec2 = vec2(0, 0), color = color(1, 1, 1, 1), size: Vec2,
                       pxRange: float32, sdThreshold: float32 = 0.5, strokeWeight: float32 = 0.0, flipY = false) =
  ctx.drawSdfImage(imageId, pos, color, size, pxRange, sdThreshold, strokeWeight, flipY, true)

method drawQuadraticBezierSdf*(ctx: CudaContext, rect: Rect, fill: figbackend.BackendFill, p0, p1, p2: Vec2,
                               strokeWeight: float32, cap: StrokeCap) =
  var (r, f) = (rectArr(rect), fill.toFdc)
  var (a, b, c) = ([p0.x.cfloat, p0.y.cfloat], [p1.x.cfloat, p1.y.cfloat], [p2.x.cfloat, p2.y.cfloat])
  ctx.ck fdc_draw_quadratic_bezier_sdf(ctx.h, r[0].addr, f.addr, a[0].addr, b[0].addr, c[0].addr, strokeWeight, cap.cint)

method drawFilledQuad*(ctx: CudaContext, verts: array[4, Vec2], colors: array[4, ColorRGBA]) =
  var v: array[8, cfloat]
  for i in 0 ..< 4: (v[2 * i], v[2 * i + 1]) = (verts[i].x.cfloat, verts[i].y.cfloat)
  var c = [colors[0].pack, colors[1].pack, colors[2].pack, colors[3].pack]
  ctx.ck fdc_draw_filled_quad(ctx.h, v[0].addr, c[0].addr)

method drawRect*(ctx: CudaContext, rect: Rect, color: Color) =
  var r = rectArr(rect)
  ctx.ck fdc_draw_rect(ctx.h, r[0].addr, color.rgba().pack)

method drawBackdropBlur*(ctx: CudaContext, rect: Rect, radii: CornerRadii2D[float32], blurRadius: float32) =
  var (r, rx, ry) = (rectArr(rect), radX(radii), radY(radii))
  ctx.ck fdc_draw_backdrop_blur(ctx.h, r[0].addr, rx[0].addr, ry[0].addr, blurRadius)

method beginMask*(ctx: CudaContext, clipRect: Rect, radii: CornerRadii2D[float32]) =
  var (r, rx, ry) = (rectArr(clipRect), radX(radii), radY(radii))
  ctx.ck fdc_begin_mask(ctx.h, r[0].addr, rx[0].addr, ry[0].addr)
method endMask*(ctx: CudaContext) = ctx.ck fdc_end_mask(ctx.h)
method popMask*(ctx: CudaContext) = ctx.ck fdc_pop_mask(ctx.h)
method beginRectMask*(ctx: CudaContext, maskRect: Rect, radii: CornerRadii2D[float32]) =
  var (r, rx, ry) = (rectArr(maskRect), radX(radii), radY(radii))
  ctx.ck fdc_begin_rect_mask(ctx.h, r[0].addr, rx[0].addr, ry[0].addr)
method popRectMask*(ctx: CudaContext) = ctx.ck fdc_pop_rect_mask(ctx.h)

proc close*(ctx: CudaContext) =
  if pointer(ctx.h) != nil:
    fdc_destroy(ctx.h)
    ctx.h = FdcCtx(nil)

# ---------------------------------------------------------------------------------------------------------------------
# Optional bulk entry points (INTEGRATION.md sections 3, 3b).  Types mirror the PODs of figdraw_cuda.h one to one.
type
  FdcCall* {.importc: "fdc_call", header: "figdraw_cuda.h", bycopy.} = object
    op*: uint32
    u*: array[9, uint32]
    f*: array[22, cfloat]
  FdcFig* {.importc: "fdc_fig", header: "figdraw_cuda.h", incompleteStruct.} = object
  FdcGlyph* {.importc: "fdc_glyph", header: "figdraw_cuda.h", incompleteStruct.} = object
  FdcDrawOp* {.importc: "fdc_draw_op", header: "figdraw_cuda.h", incompleteStruct.} = object
  FdcTextRect* {.importc: "fdc_text_rect", header: "figdraw_cuda.h", incompleteStruct.} = object
  FdcRenderList* {.importc: "fdc_render_list", header: "figdraw_cuda.h", bycopy.} = object
    nodes*: ptr FdcFig
    n_nodes*: uint32
    root_ids*: ptr int32
    n_roots*: uint32
  FdcScene* {.importc: "fdc_scene", header: "figdraw_cuda.h", bycopy.} = object
    lists*: ptr FdcRenderList
    n_lists*: uint32
    glyphs*: ptr FdcGlyph
    text_rects*: ptr FdcTextRect
    ops*: ptr FdcDrawOp
    points*: ptr cfloat

proc fdc_submit_calls(ctx: FdcCtx, calls: ptr FdcCall, n: csize_t): cint {.importc, header: "figdraw_cuda.h".}
proc fdc_submit_draws(ctx: FdcCtx, draws: ptr FdcCall, n: csize_t): cint {.importc, header: "figdraw_cuda.h".}
proc fdc_render_frame(ctx: FdcCtx, scene: ptr FdcScene, uiScale, frameW, frameH: cfloat, clearMain: cint,
                      clearRgba: ptr cfloat): cint {.importc, header: "figdraw_cuda.h".}

proc submitCalls*(ctx: CudaContext, calls: openArray[FdcCall]) =
  ## A recorded display list (one record per backend call) replayed as if the methods above had been called.
  if calls.len > 0: ctx.ck fdc_submit_calls(ctx.h, calls[0].unsafeAddr, calls.len.csize_t)

proc renderFrameNative*(ctx: CudaContext, scene: var FdcScene, frameSize: Vec2, clearMain: bool, clearColor: Color) =
  ## renderFrame (figrender.nim:1960-2002) with the node DFS done inside the library: `scene.lists` point at POD copies of
  ## `Renders.layers[*].nodes` (fdc_fig), in table order.
  var rgba = [clearColor.r.cfloat, clearColor.g.cfloat, clearColor.b.cfloat, clearColor.a.cfloat]
  ctx.ck fdc_render_frame(ctx.h, scene.addr, figUiScale().cfloat, frameSize.x.cfloat, frameSize.y.cfloat, clearMain.cint,
                          rgba[0].addr)


# ---- round-2 additions: optional fast paths a maintainer can adopt one by one ------------------------------------------

proc exportFramebuffer*(ctx: CudaContext, size: Vec2): tuple[fd: cint, bytes: int] =
  ## Present without `readPixels`: import `fd` once with Vulkan (OPAQUE_FD) / GL (EXT_memory_object_fd) / CUDA and sample
  ## the RGBA8 rows (pitch = width * 4) after every `endFrame`.  Replaces the glReadPixels path of glcontext.nim:2094-2135
  ## for a presenter on the same machine.
  var bytes: csize_t
  ctx.ck fdc_export_framebuffer(ctx.h, size.x.cint, size.y.cint, result.fd.addr, bytes.addr)
  result.bytes = bytes.int

proc replayFrame*(ctx: CudaContext) =
  ## Re-render the recorded frame unchanged (one CUDA-graph launch): what `renderFrame` amounts to when the `Renders`
  ## did not change since the previous frame.
  ctx.ck fdc_replay_frame(ctx.h)

proc generateGlyphs*(ctx: CudaContext, jobs: openArray[FdcGlyphJob], segs: openArray[FdcOutlineSeg], lcd: bool) =
  ## GPU replacement for pixie_raster.nim:45-95 (`generateGlyphImage` + `putImage`): outlines in, atlas texels out.
  ## The caller fills `segs` from `typeface.getGlyphPath(rune)` scaled to the bitmap (y down).  Parity unpinned: the
  ## anti-aliasing is exact-area, not pixie's.
  var rebuilt: cint
  ctx.ck fdc_rasterize_glyphs(ctx.h, jobs[0].unsafeAddr, jobs.len.csize_t, segs[0].unsafeAddr, segs.len.csize_t,
                              lcd.cint, rebuilt.addr)
  if rebuilt != 0:
    ctx.entries.clear()
    ctx.atlasEntryMeta.clear()
    ctx.noteAtlasRebuilt()


# ---- the rest of include/figdraw_cuda.h: plumbing a host may want, declared so that the binding is complete --------------
type
  FdcRect64* {.importc: "fdc_rect64", header: "figdraw_cuda.h", bycopy.} = object   ## 64-byte rounded-rect record
  FdcFrameStats* {.importc: "fdc_frame_stats", header: "figdraw_cuda.h", bycopy.} = object
  FdcAtlasUsage* {.importc: "fdc_atlas_usage", header: "figdraw_cuda.h", bycopy.} = object
  FdcFlattenEnv* {.importc: "fdc_flatten_env", header: "figdraw_cuda.h", bycopy.} = object

proc fdc_abi_version(): cint {.importc, header: hdr.}
proc fdc_read_pixels_async(ctx: FdcCtx, x, y, w, h: cint, outRgba: ptr uint8): cint {.importc, header: hdr.}
proc fdc_get_transform(ctx: FdcCtx, outMat4: ptr cfloat): cint {.importc, header: hdr.}
proc fdc_has_image(ctx: FdcCtx, key: uint64): cint {.importc, header: hdr.}
proc fdc_get_image_rect(ctx: FdcCtx, key: uint64, outRect: ptr cfloat): cint {.importc, header: hdr.}
proc fdc_get_atlas_usage(ctx: FdcCtx, outUsage: ptr FdcAtlasUsage): cint {.importc, header: hdr.}
proc fdc_get_frame_stats(ctx: FdcCtx, outStats: ptr FdcFrameStats): cint {.importc, header: hdr.}
# compact records
proc fdc_pack_rect64(call: ptr FdcCall, outRect: ptr FdcRect64): cint {.importc, header: hdr.}
proc fdc_expand_rect64(rect: ptr FdcRect64, outCall: ptr FdcCall) {.importc, header: hdr.}
proc fdc_submit_rects64(ctx: FdcCtx, rects: ptr FdcRect64, n: csize_t): cint {.importc, header: hdr.}
# native front-end without a context (tests: byte-identity with the per-call front-end)
proc fdc_flatten_renders(scene: ptr FdcScene, env: ptr FdcFlattenEnv, outCalls: ptr FdcCall, cap: csize_t,
                         nOut: ptr csize_t): cint {.importc, header: hdr.}
# framebuffer plumbing: caller-owned memory, bands, peers over CUDA IPC, gather mode
proc fdc_bind_framebuffer(ctx: FdcCtx, deviceRgba8: pointer): cint {.importc, header: hdr.}
proc fdc_framebuffer_ptr(ctx: FdcCtx): pointer {.importc, header: hdr.}
proc fdc_band_rows(ctx: FdcCtx, y0, y1: ptr cint): cint {.importc, header: hdr.}
proc fdc_stream(ctx: FdcCtx): pointer {.importc, header: hdr.}
proc fdc_set_peer_framebuffers(ctx: FdcCtx, devicePtrs: ptr pointer, n: cint): cint {.importc, header: hdr.}
proc fdc_set_frame_barrier(ctx: FdcCtx, enabled: cint): cint {.importc, header: hdr.}
proc fdc_set_peer_gather(ctx: FdcCtx, mode, subBands: cint): cint {.importc, header: hdr.}
proc fdc_reserve_framebuffer(ctx: FdcCtx, width, rows: cint): cint {.importc, header: hdr.}
proc fdc_framebuffer_ipc_handle(ctx: FdcCtx, outHandle: ptr uint8): cint {.importc, header: hdr.}
proc fdc_open_peer_framebuffer(ctx: FdcCtx, handle: ptr uint8, outDevicePtr: ptr pointer): cint {.importc, header: hdr.}
# debugging
proc fdc_debug_bins(ctx: FdcCtx, segment: cint, tileOffsets: ptr uint32, offsetsCap: csize_t, entries: ptr uint32,
                    entriesCap: csize_t, nOffsets, nEntries: ptr csize_t): cint {.importc, header: hdr.}
proc fdc_debug_limit_lists(ctx: FdcCtx, coarseEntries, tileEntries: uint32): cint {.importc, header: hdr.}
proc fdc_debug_shade_stats(ctx: FdcCtx, outStats: ptr uint64): cint {.importc, header: hdr.}

proc rebalanceBands*(ctx: CudaContext, bounds: openArray[cint]) =
  ## Bands chosen by the host (the same boundaries on every rank): tile rows, `bounds[0] = 0`, `bounds[^1] = ceil(H / 16)`.
  ## The profile to split comes from `fdc_get_tile_row_costs`, summed over the ranks.
  ctx.ck fdc_set_band_tile_rows(ctx.h, bounds[0].unsafeAddr, bounds.len.cint)
