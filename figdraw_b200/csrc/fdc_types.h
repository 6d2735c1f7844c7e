// Shared host/device types of the figdraw CUDA backend (sm_100a).
//
// Data layout in HBM for one frame (see DESIGN.md "Data layout"):
//   RawDraw[n]   the draw records exactly as they crossed the C ABI (fdc_call, 128 B) -- what GL's ten vertex
//                attribute arrays carried, once per quad instead of four times.
//   RunState[r]  per-run state (transform, AA, mask depth, rect mask, clip parent); a run is a maximal range
//                of draws issued under the same backend state.  Draw i finds its run by binary search.
//   Prim[n]      128-byte shading record produced on the device by prim_setup_kernel.
//   coarse/tile lists: u32 primitive indices, order preserving.
#pragma once
#include <stdint.h>

#include "../../include/figdraw_cuda.h"

namespace fdc {

constexpr int kTileW = 16;         // pixels
constexpr int kTileH = 16;
constexpr int kCoarse = 8;         // coarse bin = kCoarse x kCoarse tiles (128 x 128 px)
#ifndef FDC_CHUNK
#define FDC_CHUNK 512
#endif
constexpr int kChunk = FDC_CHUNK;  // primitives per coarse-binning chunk (one CTA; 32 per warp)
constexpr int kMaxMaskDepth = 15;  // texture-mask nesting (GL: unbounded): levels 1..8 live in two registers per pixel, 9..15 in
                                   // shared memory (the depth field of a tile entry has 4 bits)
constexpr int kAtlasMargin = 4;    // glcontext.nim:257
constexpr int kMaxAtlasLevels = 14;
constexpr int kPrimFastBytes = 80; // q0..q4
constexpr uint64_t kRectKey = 0x7265637472656374ull;  // stands for hash("rect"), glcontext.nim:966

// Prim.mode_flags layout
constexpr uint32_t PF_MODE_MASK = 0x1Fu;        // SdfMode 0..20
constexpr uint32_t PF_ELLIPTICAL = 1u << 5;     // sdfMode + 128 in the reference
constexpr uint32_t PF_FILLMODE_SHIFT = 6;       // 3 bits: 0 vertex colours, 1..4 linear3 X/Y/TLBR/BLTR
constexpr uint32_t PF_FILLMODE_MASK = 7u << 6;
constexpr uint32_t PF_MASK_WRITE = 1u << 9;     // drawn between beginMask/endMask: blends into mask level `depth`
constexpr uint32_t PF_MASK_BEGIN = 1u << 10;    // first primitive of a mask level: level is cleared to 0 first
constexpr uint32_t PF_GENERAL = 1u << 11;       // not an axis-aligned quad: use QuadGeom
constexpr uint32_t PF_SOLID = 1u << 12;         // four equal vertex colours
constexpr uint32_t PF_OCCLUDER = 1u << 13;      // opaque ClipAA fill: inner rect (ix0..iy1) has alpha exactly 1
constexpr uint32_t PF_RECTMASK = 1u << 14;      // fast rect mask applies (index in aux)
constexpr uint32_t PF_SUBPIXEL = 1u << 15;      // atlas: subpixel shift enabled
constexpr uint32_t PF_DEPTH_SHIFT = 16;         // 4 bits: texture-mask level read (content) or written (mask)
constexpr uint32_t PF_DEPTH_MASK = 0xFu << 16;
constexpr uint32_t PF_FAST = 1u << 20;          // axis-aligned circular-corner ClipAA / AnnularAA / DropShadow content quad
constexpr uint32_t PF_INNER = 1u << 21;         // inner rect (ix0..iy1) valid: coverage is exactly 1 there ...
constexpr uint32_t PF_INNER_EMPTY = 1u << 22;   // ... or exactly 0 (interior of an AnnularAA stroke)
constexpr uint32_t PF_VISIT_FULL = 1u << 23;    // shade kernel only: the warp's block lies inside the inner rect
constexpr uint32_t PF_MASK_WIDE = 1u << 24;     // first primitive of a mask level that holds SEVERAL draws: binned over the parent's
                                                // whole clip box so the level is cleared everywhere (GL clears the full mask texture,
                                                // glcontext.nim:1901-1902); slow-path primitives keep their own bbox in ix0..iy1
constexpr uint32_t PF_CLEAR_ONLY = 1u << 25;    // PF_MASK_WIDE primitive whose own quad is empty: clears the level, draws nothing
constexpr uint32_t PF_EMPTY = 1u << 31;         // dropped (early-out or empty clipped bbox)

// Tile-list entry (8 bytes): .x = primitive index (bit 31: unused), .y = everything the shade kernel needs to decide
// what each of the tile's 8 warps (8x4-pixel blocks; block = row*2 + col) does with the primitive, precomputed once
// per (tile, primitive) by fine_bin_kernel instead of 8 times per pair by the shading warps.
// Compile-time switches for the optional fast paths (A/B measurements; all on by default).
#ifndef FDC_FAST_TEX
#define FDC_FAST_TEX 1
#endif
#ifndef FDC_FAST_MASK
#define FDC_FAST_MASK 1
#endif

constexpr uint32_t TE_OV_SHIFT = 0;      // bits 0..7 : the primitive's clipped bbox overlaps block b
constexpr uint32_t TE_FULL_SHIFT = 8;    // bits 8..15: block b lies inside the inner rect (coverage exactly 1)
constexpr uint32_t TE_FAST = 1u << 16;   // PF_FAST
constexpr uint32_t TE_SOLID = 1u << 17;  // PF_SOLID
constexpr uint32_t TE_GRAD3 = 1u << 18;  // 3-stop fill (fill mode != 0)
constexpr uint32_t TE_KIND_SHIFT = 19;   // 2 bits: 0 ClipAA, 1 AnnularAA, 2 DropShadow; with TE_TEX: 0 atlas, 1 MSDF, 2 MTSDF (fast primitives)
constexpr uint32_t TE_OCCLUDER = 1u << 21;
constexpr uint32_t TE_TEX = 1u << 22;    // fast primitive that samples the atlas (kind selects atlas / MSDF / MTSDF; TE_GRAD3 = annular stroke)
constexpr uint32_t TE_MASKW = 1u << 23;  // fast ClipAA mask write (PF_MASK_WRITE): updates mask level DEPTH instead of the pixel
constexpr uint32_t TE_DEPTH_SHIFT = 24;  // 4 bits: texture-mask level read by content (written by TE_MASKW)
constexpr uint32_t TE_MASKB = 1u << 28;  // PF_MASK_BEGIN: clear the level first
constexpr uint32_t TE_RECTMASK = 1u << 29;  // content under a first-level analytic rect mask (PF_RECTMASK)
// Device counters of the binning pipeline (BinBuffers::counters).  Words 0..7 are per SEGMENT (launch_binning zeroes
// them), words 8..15 per FRAME (zeroed once by execute_frame): an overflow in any segment stays visible to the host.
enum : int {
  kCntCursor = 0,          // tile-list cursor: entries reserved so far (= size the segment needs when it overflows)
  kCntOverflow = 1,        // bit 0 coarse list, bit 1 tile list overflowed in this segment
  kCntCoarseTotal = 2,     // entries of the coarse list
  kCntFullTiles = 4,       // tiles of this segment that need the shade kernel's full loop
  kCntStickyOverflow = 8,  // OR of kCntOverflow over the frame's segments
  kCntMaxCoarse = 9,       // largest coarse list any segment needs
  kCntMaxTile = 10,        // largest tile list any segment needs
  kCntSumEntries = 11,     // tile entries of the whole frame (statistics)
  kNumCounters = 16
};

// tile_count[tile]: bits 0..23 the number of entries; bit 31 set by the fine binner when the tile holds anything but
// unmasked PF_FAST content (a general-path primitive, a mask write, content under a texture mask or a rect mask) --
// such tiles are shaded by the full loop, all others by the call-free lean loop.
constexpr uint32_t kTileCountMask = 0x00FFFFFFu;
constexpr uint32_t kTileNeedsFullPath = 1u << 31;
struct alignas(8) TileEntry {
  uint32_t pid, info;
};

// 128-byte shading record, eight 16-byte quads q0..q7.  q0..q4 (80 bytes) are everything the shade kernel's fast path
// reads; they are the part each warp stages into shared memory with one bulk (TMA) copy per primitive.
struct alignas(16) Prim {
  // q0: SDF modes:   SDF-space position from the pixel index: p.x = x*u0 + du ; -p.y = y*v0 + dv
  //     atlas modes: texel mapping tu = s*du + u0, tv = t*dv + v0 (level-0 texels, already minus 0.5)
  float u0, du, v0, dv;
  // q1: sdfParams (quadHalf.xy, shapeHalf.xy | inset offset | bezier p0 | (atlasSize, strokeW))
  float qhx, qhy, p2, p3;
  // q2: sdfRadii (TR,BR,TL,BL | packed elliptical | bezier p1,p2)
  float r0, r1, r2, r3;
  // q3: factor, spread (or midPos), aa factor, k: shadows -0.5*log2(e)/sigma^2; atlas lambda (LOD); msdf screenPxRange
  float factor, spread, aa, k;
  // q4: vertex colours BL, BR, TR, TL packed RGBA8; PF_SOLID: the colour as four floats 0..255 (bit patterns)
  uint32_t c[4];
  // q5: 3-stop colours and the inner rect [ix0,ix1) x [iy0,iy1) (pixels)
  uint32_t c_mid, c_stop;
  int16_t ix0, iy0, ix1, iy1;
  // q6: clipped bin bbox [bx0,bx1) x [by0,by1) (pixels), flags, aux: rect mask index+1 (low 16)
  int16_t bx0, by0, bx1, by1;
  uint32_t mode_flags, aux;
  // q7: quad-local (s,t) in [0,1] from the integer pixel index: s = x*su + ou, t = y*sv + ov (axis aligned).
  //     PF_GENERAL: su's bits hold the QuadGeom index.
  float su, ou, sv, ov;
};
static_assert(sizeof(Prim) == 128, "Prim must be 128 bytes");

// What the binning kernels need of a primitive, 32 bytes in an array of its own: they stream / gather these instead of
// two 16-byte pieces out of every 128-byte Prim (4x fewer sectors for the coarse passes, one sector per gathered entry
// in the fine pass).
struct alignas(32) PrimBin {
  int16_t bx0, by0, bx1, by1;  // = Prim q6: clipped bin bbox
  uint32_t mode_flags, aux;
  int16_t ix0, iy0, ix1, iy1;  // = Prim inner rect (valid with PF_INNER)
  uint32_t pad_[2];
};
static_assert(sizeof(PrimBin) == 32, "PrimBin must be 32 bytes");

// Gradient colours of a PF_FAST primitive as floats (0..255), evaluated straight from the pixel index.
//   3-stop (fill mode 1..4): tt = sat(x*ta + y*tb + tc); colour = tt <= mid ? a0 + d0*tt : a1 + d1*tt
//   vertex colours (affine): colour = a0 + d0*x + a1*y
struct alignas(16) PrimExt {
  float ta, tb, tc, mid;
  float a0[4], d0[4], a1[4], d1[4];
};
static_assert(sizeof(PrimExt) == 80, "PrimExt must be 80 bytes");

// Geometry of a general (rotated / arbitrary) quad: ceil'd integer vertices BL, BR, TR, TL.
struct alignas(16) QuadGeom {
  int32_t vx[4], vy[4];
};

// Per-run backend state, resolved on the host (cheap, sequential) and consumed by prim_setup_kernel.
struct alignas(16) RunState {
  uint32_t first_draw;  // index of the first draw of this run (ascending)
  uint32_t xform;       // index into the transform table
  float aa;
  float subpixel_shift;  // already clamped; < 0 when positioning is disabled
  uint32_t flags;        // PF_MASK_WRITE | PF_MASK_BEGIN | depth bits
  uint32_t rectmask;     // 0 = none, else index+1 into the rect mask table
  int32_t clip_draw;     // draw index of the mask primitive whose bbox clips this run's draws, or -1
  uint32_t call_index;   // backend-call ordinal of the first draw (draws of a run are consecutive calls)
  uint32_t compact;      // 1: the run's records are fdc_rect64 at rects64[src_off + (draw - first_draw)]
  uint32_t src_off;
  uint32_t pad_[2];
};
static_assert(sizeof(RunState) == 48, "RunState must be 48 bytes");

// fdc_rect64 -> the 32 words of the fdc_call it was packed from (op, u[9], f[22]).  Shared by the host helper and the
// setup kernel, so what the tests check on the CPU is what runs on the device.
#ifdef __CUDACC__
#define FDC_HD __host__ __device__
#else
#define FDC_HD
#endif
FDC_HD inline void expand_rect64_words(const fdc_rect64& r, uint32_t* w) {
  union { float f; uint32_t u; } cv;
  for (int k = 0; k < 32; k++) w[k] = 0u;
  w[0] = FDC_OP_ROUNDED_RECT;
  const uint32_t kind = (r.packed >> 8) & 3u;
  w[1] = r.packed & 255u;          // u[0] mode
  w[2] = kind;                     // u[1] fill kind
  w[3] = (r.packed >> 10) & 3u;    // u[2] axis
  w[4] = r.c[0]; w[5] = r.c[1]; w[6] = r.c[2];
  for (int k = 0; k < 4; k++) {
    cv.f = r.rect[k]; w[10 + k] = cv.u;
    cv.f = r.radii[k]; w[14 + k] = cv.u; w[18 + k] = cv.u;
  }
  cv.f = r.factor; w[22] = cv.u;
  cv.f = r.spread; w[23] = cv.u;
  cv.f = r.shape_size[0]; w[24] = cv.u;
  cv.f = r.shape_size[1]; w[25] = cv.u;
  float mid = 0.5f;
  if (kind == (uint32_t)FDC_FILL_LINEAR3) {
    mid = (float)((r.packed >> 16) & 255u) / 255.0f;
    mid = mid < 0.01f ? 0.01f : (mid > 0.99f ? 0.99f : mid);
  }
  cv.f = mid; w[26] = cv.u;        // f[16] mid_pos
}

// 2-D affine part of the transform (vmath Mat4 columns 0, 1, 3; rows x, y).
struct alignas(8) Xform {
  float m00, m10, m30, m01, m11, m31;  // x' = (m00*x + m10*y) + m30 ; y' = (m01*x + m11*y) + m31
};

// Fast rect mask, atlas_rect_mask.frag:222-237 / glcontext.nim:831-850.
struct alignas(16) RectMaskRec {
  float cx, cy, hx, hy;          // params
  float r0, r1, r2, r3;          // radii (TR,BR,TL,BL or packed)
  float ax, ay, az, elliptical;  // matX.xyz, matY.w
  float bx, by, bz, pad;         // matY.xyz
};

// Atlas entry table mirrored on the device (open addressing, linear probing) so image draws resolve their
// atlas rect in prim_setup_kernel instead of on the host.
struct alignas(16) AtlasEntry {
  uint64_t key;
  uint32_t used, pad;
  float x, y, w, h;  // normalised rect (x,y,w,h)/atlasSize, as `entries[key]` in the reference
};
static_assert(sizeof(AtlasEntry) == 32, "AtlasEntry must be 32 bytes");

struct AtlasView {
  const uint8_t* level[kMaxAtlasLevels];  // RGBA8, level l is (size >> l)^2
  const AtlasEntry* table;
  uint32_t table_mask;  // capacity - 1 (power of two)
  int size;
  int n_levels;
  int pixelate;  // magnification filter GL_NEAREST (`newContext(pixelate = true)`, glcontext.nim:165-168)
};

struct FrameView {
  int W, H;              // frame size in pixels
  int tiles_x, tiles_y;  // full frame tile grid
  int ty0, ty1;          // tile rows this rank owns (band)
  int cbx, cby;          // coarse bin grid covering the band
  int cty0;              // first tile row of coarse row 0 (= ty0 rounded down to a multiple of kCoarse)
  int band_y0, band_y1;  // pixel rows of the band
};

__host__ __device__ inline uint32_t atlas_hash(uint64_t key) {
  key ^= key >> 33;
  key *= 0xff51afd7ed558ccdull;
  key ^= key >> 33;
  key *= 0xc4ceb9fe1a85ec53ull;
  key ^= key >> 33;
  return (uint32_t)key;
}

}  // namespace fdc
