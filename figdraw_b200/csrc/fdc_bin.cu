// Primitive setup and order-preserving primitive-to-tile binning (sm_100a).
//
// prim_setup_kernel  -- the device restatement of the reference's quad emission: `drawRoundedRectSdfOpenGl`
//   (glcontext.nim:1449-1559), `drawUvRect*` (:1022-1095, :1169-1302), `drawQuadraticBezierSdfOpenGl` (:1619-1711),
//   `drawFilledQuad` (:963-982), `roundedRadiiVec` (:745-817), `encodeSdfMode` (:1002-1008) and
//   `gradientColors` (figbackend.nim:129-183).  One thread per draw record; the host only appends records.
//   Everything that feeds `ceil` is computed with explicitly rounded float32 ops in the reference's operation
//   order so the integer quad corners -- and therefore the bin lists -- are bit-exact against the oracle.
//
// Binning -- two levels, both tile-major so list order == emission order without sorting or atomics on order:
//   coarse: the frame (band) is cut into 128x128-px bins; primitives into chunks of 512.  ONE kernel: a CTA enumerates
//           the (primitive, bin) pairs of its chunk in primitive order, counts them per bin, reserves a region of the
//           coarse list and scatters; a table says where each (bin, chunk) segment sits (coarse_pairs_kernel).
//   fine:   one CTA per coarse bin stages its entries 1024 at a time (chunk by chunk through the table); each warp takes
//           groups of 32 entries, two transposes give lane t the entries of tiles t and 32+t.  Counts -> CTA scan -> one
//           atomicAdd reserves the bin's slice of the tile list (slice placement is the only non-deterministic thing
//           and is not observable) -> the lanes write their tiles' 8-byte TileEntry records in bit order.
#include <cuda_runtime.h>

#include "fdc_kernels.h"

namespace fdc {

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// `ctx.mat * vec2`, glcontext.nim:905-906, vmath order: (m00*x + m10*y) + m30, every op rounded once.
__device__ __forceinline__ float2 xf_apply(const Xform& m, float x, float y) {
  float2 r;
  r.x = fadd(fadd(fmul(m.m00, x), fmul(m.m10, y)), m.m30);
  r.y = fadd(fadd(fmul(m.m01, x), fmul(m.m11, y)), m.m31);
  return r;
}

__device__ __forceinline__ float nim_round(float x) { return roundf(x); }  // half away from zero
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// clampRadius, glcontext.nim:745-749
__device__ __forceinline__ float clamp_radius(float r, float maxr) {
  if (r <= 0.0f) return 0.0f;
  return nim_round(fmaxf(1.0f, fminf(r, maxr)));
}

// roundedRadiiVec, glcontext.nim:751-817.  rx/ry: TL,TR,BL,BR.  out: TR,BR,TL,BL.
__device__ bool rounded_radii_vec(const float* rx, const float* ry, float hx, float hy, float out[4]) {
  const int order[4] = {1, 3, 0, 2};
  bool circ = true;
#pragma unroll
  for (int i = 0; i < 4; i++) circ = circ && (rx[i] == ry[i]);
  float mr = fminf(hx, hy);
  if (circ) {
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = clamp_radius(rx[order[k]], mr);
    return false;
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int c = order[k];
    float cx = clamp_radius(rx[c], hx), cy = clamp_radius(ry[c], hy);
    if (rx[c] == ry[c]) {
      out[k] = -(clamp_radius(rx[c], mr) + 1.0f);
    } else if (cx == cy) {
      out[k] = -(cx + 1.0f);
    } else {
      float qx = nim_round(fmul(clampf(fdiv(cx, fmaxf(hx, 0.000001f)), 0.0f, 1.0f), 4095.0f));
      float qy = nim_round(fmul(clampf(fdiv(cy, fmaxf(hy, 0.000001f)), 0.0f, 1.0f), 4095.0f));
      out[k] = fadd(qx, fmul(qy, 4096.0f));
    }
  }
  return true;
}

// lerpColor / sampleColor / gradientColors, figbackend.nim:129-183
__device__ uint32_t lerp_color(uint32_t a, uint32_t b, float t) {
  float ct = clampf(t, 0.0f, 1.0f), inv = __fsub_rn(1.0f, ct);
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float av = (float)((a >> (8 * k)) & 255u), bv = (float)((b >> (8 * k)) & 255u);
    float v = nim_round(fadd(fmul(av, inv), fmul(bv, ct)));
    r |= ((uint32_t)(int)v & 255u) << (8 * k);
  }
  return r;
}
__device__ uint32_t sample_color(int kind, const uint32_t* c, float mid, float t) {
  if (kind == FDC_FILL_COLOR) return c[0];
  if (kind == FDC_FILL_LINEAR2) return lerp_color(c[0], c[1], t);
  float ct = clampf(t, 0.0f, 1.0f);
  if (ct <= mid) return lerp_color(c[0], c[1], fdiv(ct, mid));
  return lerp_color(c[1], c[2], fdiv(__fsub_rn(ct, mid), __fsub_rn(1.0f, mid)));
}
__device__ void gradient_colors(int kind, int axis, const uint32_t* c, float mid, uint32_t out[4]) {
  const float ts[4][4] = {{0.0f, 1.0f, 1.0f, 0.0f}, {1.0f, 1.0f, 0.0f, 0.0f}, {0.5f, 1.0f, 0.5f, 0.0f}, {0.0f, 0.5f, 1.0f, 0.5f}};
  if (kind == FDC_FILL_COLORS4) {
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = c[k];
    return;
  }
  if (kind == FDC_FILL_COLOR) axis = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = sample_color(kind, c, mid, ts[axis & 3][k]);
}

__device__ __forceinline__ int find_run_in(const RunState* runs, int lo, int hi, uint32_t draw) {
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (runs[mid].first_draw <= draw) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__device__ __forceinline__ int find_run(const RunState* runs, int n_runs, uint32_t draw) {
  int lo = 0, hi = n_runs - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (runs[mid].first_draw <= draw) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ bool atlas_lookup(const AtlasView& at, uint64_t key, float rect[4]) {
  if (!at.table) return false;
  uint32_t h = atlas_hash(key) & at.table_mask;
  for (uint32_t probe = 0; probe <= at.table_mask; probe++) {
    const AtlasEntry& e = at.table[(h + probe) & at.table_mask];
    if (!e.used) return false;
    if (e.used == 1 && e.key == key) {
      rect[0] = e.x; rect[1] = e.y; rect[2] = e.w; rect[3] = e.h;
      return true;
    }
  }
  return false;
}

struct QuadPos {
  float x[4], y[4];  // BL, BR, TR, TL after ceil
};

__device__ __forceinline__ void quad_from_rect(const Xform& m, float atx, float aty, float tox, float toy, QuadPos& q) {
  float2 p0 = xf_apply(m, atx, toy), p1 = xf_apply(m, tox, toy), p2 = xf_apply(m, tox, aty), p3 = xf_apply(m, atx, aty);
  q.x[0] = ceilf(p0.x); q.y[0] = ceilf(p0.y);
  q.x[1] = ceilf(p1.x); q.y[1] = ceilf(p1.y);
  q.x[2] = ceilf(p2.x); q.y[2] = ceilf(p2.y);
  q.x[3] = ceilf(p3.x); q.y[3] = ceilf(p3.y);
}

__device__ __forceinline__ int clamp_i(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int f2i_sat(float v) {  // float -> int with saturation well inside int16 range
  v = fminf(fmaxf(v, -30000.0f), 30000.0f);
  return (int)v;
}

// Pixel bbox [x0,x1) x [y0,y1) of a quad's ceil'd corners.
__device__ __forceinline__ void quad_bbox(const QuadPos& q, int& x0, int& y0, int& x1, int& y1) {
  x0 = f2i_sat(fminf(fminf(q.x[0], q.x[1]), fminf(q.x[2], q.x[3])));
  x1 = f2i_sat(fmaxf(fmaxf(q.x[0], q.x[1]), fmaxf(q.x[2], q.x[3])));
  y0 = f2i_sat(fminf(fminf(q.y[0], q.y[1]), fminf(q.y[2], q.y[3])));
  y1 = f2i_sat(fmaxf(fmaxf(q.y[0], q.y[1]), fmaxf(q.y[2], q.y[3])));
}

// Clip chain: intersect with the bbox of every enclosing texture-mask primitive (they are ROUNDED_RECT records).
__device__ void apply_clip_chain(const SetupArgs& a, int clip_draw, int& x0, int& y0, int& x1, int& y1) {
  for (int guard = 0; clip_draw != -1 && guard < 2 * kMaxMaskDepth; guard++) {
    if (clip_draw < -1) { x1 = x0; return; }  // -2: the enclosing mask level is empty, nothing under it is visible
    const fdc_call& d = a.draws[clip_draw];
    const RunState& rs = a.runs[find_run(a.runs, a.n_runs, (uint32_t)clip_draw)];
    QuadPos q;
    quad_from_rect(a.xforms[rs.xform], d.f[0], d.f[1], fadd(d.f[0], d.f[2]), fadd(d.f[1], d.f[3]), q);
    int cx0, cy0, cx1, cy1;
    quad_bbox(q, cx0, cy0, cx1, cy1);
    x0 = max(x0, cx0); y0 = max(y0, cy0); x1 = min(x1, cx1); y1 = min(y1, cy1);
    clip_draw = rs.clip_draw;
  }
}

// One staged draw record (fdc_call layout: op, u[9], f[22]) in shared memory.
struct DrawView {
  const uint32_t* w;
  __device__ __forceinline__ uint32_t op() const { return w[0]; }
  __device__ __forceinline__ uint32_t u(int k) const { return w[1 + k]; }
  __device__ __forceinline__ float f(int k) const { return __uint_as_float(w[10 + k]); }
  __device__ __forceinline__ const uint32_t* colors() const { return w + 4; }                              // u[3..6]
  __device__ __forceinline__ const float* radii_x() const { return reinterpret_cast<const float*>(w + 14); }  // f[4..7]
  __device__ __forceinline__ const float* radii_y() const { return reinterpret_cast<const float*>(w + 18); }  // f[8..11]
};

// The same view of a 64-byte compact record (fdc_rect64), straight from registers: what expand_rect64_words would have
// written, field by field (every index below is a literal at the call sites, so the switches fold away), without the
// round trip through the staging slot.  A compact record is always a rounded rect: the other cases of setup_record's
// switch are not even compiled into this instantiation.
struct RectView {
  fdc_rect64 r;
  float mid;
  __device__ __forceinline__ explicit RectView(const fdc_rect64& rec) : r(rec), mid(0.5f) {
    if (((r.packed >> 8) & 3u) == (uint32_t)FDC_FILL_LINEAR3) {
      mid = (float)((r.packed >> 16) & 255u) / 255.0f;
      mid = mid < 0.01f ? 0.01f : (mid > 0.99f ? 0.99f : mid);
    }
  }
  __device__ __forceinline__ uint32_t op() const { return FDC_OP_ROUNDED_RECT; }
  __device__ __forceinline__ uint32_t u(int k) const {
    switch (k) {
      case 0: return r.packed & 255u;
      case 1: return (r.packed >> 8) & 3u;
      case 2: return (r.packed >> 10) & 3u;
      case 3: return r.c[0];
      case 4: return r.c[1];
      case 5: return r.c[2];
      default: return 0u;
    }
  }
  __device__ __forceinline__ float f(int k) const {
    switch (k) {
      case 0: case 1: case 2: case 3: return r.rect[k];
      case 4: case 5: case 6: case 7: return r.radii[k - 4];
      case 8: case 9: case 10: case 11: return r.radii[k - 8];
      case 12: return r.factor;
      case 13: return r.spread;
      case 14: return r.shape_size[0];
      case 15: return r.shape_size[1];
      case 16: return mid;
      default: return 0.0f;
    }
  }
  __device__ __forceinline__ const uint32_t* colors() const { return r.c; }
  __device__ __forceinline__ const float* radii_x() const { return r.radii; }
  __device__ __forceinline__ const float* radii_y() const { return r.radii; }
};

// Full setup of draw record `i` of the segment; `rec` is its staging slot in shared memory (33 words).
template <class Rec>
__device__ __forceinline__ void setup_record_from(const SetupArgs& a, uint32_t i, const Rec& d, const RunState& rs);

__device__ void setup_record(const SetupArgs& a, uint32_t i, uint32_t* rec, int run_lo, int run_hi) {
  const uint32_t di = a.first + i;
  const int ri = find_run_in(a.runs, run_lo, run_hi, di);  // the CTA's records usually share one run: no search at all
  const RunState rs = a.runs[ri];
  if (rs.compact) {
    // the run arrived as 64-byte fdc_rect64 records
    const RectView d(a.rects64[rs.src_off + (di - rs.first_draw)]);
    setup_record_from(a, i, d, rs);
  } else {
    const DrawView d{rec};
    setup_record_from(a, i, d, rs);
  }
}

template <class Rec>
__device__ __forceinline__ void setup_record_from(const SetupArgs& a, uint32_t i, const Rec& d, const RunState& rs) {
  const uint32_t di = a.first + i;
  const Xform xf = a.xforms[rs.xform];
  a.prim_call[i] = rs.call_index + (di - rs.first_draw);

  Prim p;
  memset(&p, 0, sizeof(p));
  uint32_t flags = rs.flags & (PF_MASK_WRITE | PF_MASK_BEGIN | PF_MASK_WIDE | PF_DEPTH_MASK);
  bool empty = false;
  QuadPos q;
  float atx = 0, aty = 0, tox = 0, toy = 0;
  bool have_rect_quad = true;
  int mode = 0, fill_mode = 0;
  uint32_t cols[4] = {0, 0, 0, 0};
  float uax = 0.0f, uay = 0.0f, utx = 1.0f, uty = 1.0f;  // normalised atlas uv corners (atlas modes)
  bool atlas_mode = false;

  switch (d.op()) {
    case FDC_OP_ROUNDED_RECT: {
      float w = d.f(2), h = d.f(3);
      if (w <= 0.0f || h <= 0.0f) { empty = true; break; }
      mode = (int)d.u(0);
      int fkind = (int)d.u(1), axis = (int)d.u(2);
      if (fkind == FDC_FILL_LINEAR3 && (mode == FDC_SDF_CLIP_AA || mode == FDC_SDF_ANNULAR || mode == FDC_SDF_ANNULAR_AA)) {
        fill_mode = 1 + (axis & 3);
        cols[0] = cols[1] = cols[2] = cols[3] = d.u(3);
        p.c_mid = d.u(4);
        p.c_stop = d.u(5);
      } else {
        gradient_colors(fkind, axis, d.colors(), d.f(16), cols);
      }
      float qhx = fmul(w, 0.5f), qhy = fmul(h, 0.5f);
      bool inset = (mode == FDC_SDF_INSET_SHADOW);
      float ssx = d.f(14), ssy = d.f(15);
      float rsx = (ssx > 0.0f && ssy > 0.0f) ? ssx : w, rsy = (ssx > 0.0f && ssy > 0.0f) ? ssy : h;
      float shx = inset ? qhx : fmul(rsx, 0.5f), shy = inset ? qhy : fmul(rsy, 0.5f);
      p.qhx = qhx; p.qhy = qhy;
      p.p2 = inset ? ssx : shx;
      p.p3 = inset ? ssy : shy;
      float rr[4];
      if (rounded_radii_vec(d.radii_x(), d.radii_y(), shx, shy, rr)) flags |= PF_ELLIPTICAL;
      p.r0 = rr[0]; p.r1 = rr[1]; p.r2 = rr[2]; p.r3 = rr[3];
      p.factor = d.f(12);
      p.spread = fill_mode == 0 ? d.f(13) : clampf(d.f(16), 0.01f, 0.99f);
      atx = d.f(0); aty = d.f(1); tox = fadd(d.f(0), w); toy = fadd(d.f(1), h);
      break;
    }
    case FDC_OP_IMAGE: {
      float r[4];
      uint64_t key = (uint64_t)d.u(0) | ((uint64_t)d.u(1) << 32);
      if (!atlas_lookup(a.atlas, key, r)) { empty = true; break; }
      float as = (float)a.atlas.size;
      float sw = d.f(2), sh = d.f(3);
      if (!(sw > 0.0f && sh > 0.0f)) { sw = fmul(r[2], as); sh = fmul(r[3], as); }
      uax = r[0]; utx = fadd(r[0], r[2]);
      uay = r[1]; uty = fadd(r[1], r[3]);
      if (d.u(7)) { float tmp = uay; uay = uty; uty = tmp; }
#pragma unroll
      for (int k = 0; k < 4; k++) cols[k] = d.u(3 + k);
      mode = FDC_SDF_ATLAS;
      atlas_mode = true;
      atx = d.f(0); aty = d.f(1); tox = fadd(d.f(0), sw); toy = fadd(d.f(1), sh);
      break;
    }
    case FDC_OP_MSDF: {
      float r[4];
      uint64_t key = (uint64_t)d.u(0) | ((uint64_t)d.u(1) << 32);
      if (!atlas_lookup(a.atlas, key, r)) { empty = true; break; }
      float stroke_w = fmaxf(0.0f, d.f(6));
      bool mtsdf = d.u(2) != 0;
      mode = stroke_w > 0.0f ? (mtsdf ? FDC_SDF_MTSDF_ANNULAR : FDC_SDF_MSDF_ANNULAR) : (mtsdf ? FDC_SDF_MTSDF : FDC_SDF_MSDF);
      uax = r[0]; utx = fadd(r[0], r[2]);
      uay = r[1]; uty = fadd(r[1], r[3]);
      if (d.u(7)) { float tmp = uay; uay = uty; uty = tmp; }
      cols[0] = cols[1] = cols[2] = cols[3] = d.u(3);
      p.qhx = (float)a.atlas.size; p.qhy = stroke_w;
      p.factor = d.f(4); p.spread = d.f(5);
      atlas_mode = true;
      atx = d.f(0); aty = d.f(1); tox = fadd(d.f(0), d.f(2)); toy = fadd(d.f(1), d.f(3));
      break;
    }
    case FDC_OP_BEZIER: {
      if (d.f(2) <= 0.0f || d.f(3) <= 0.0f || d.f(10) <= 0.0f) { empty = true; break; }
      int fkind = (int)d.u(1), axis = (int)d.u(2);
      if (fkind == FDC_FILL_LINEAR3) {
        fill_mode = 1 + (axis & 3);
        cols[0] = cols[1] = cols[2] = cols[3] = d.u(3);
        p.c_mid = d.u(4);
        p.c_stop = d.u(5);
      } else {
        gradient_colors(fkind, axis, d.colors(), d.f(16), cols);
      }
      p.qhx = fmul(d.f(2), 0.5f); p.qhy = fmul(d.f(3), 0.5f); p.p2 = d.f(4); p.p3 = d.f(5);
      p.r0 = d.f(6); p.r1 = d.f(7); p.r2 = d.f(8); p.r3 = d.f(9);
      p.factor = d.f(10);
      p.spread = fill_mode == 0 ? 0.0f : clampf(d.f(16), 0.01f, 0.99f);
      int cap = (int)d.u(0);
      mode = cap == FDC_CAP_BUTT ? FDC_SDF_BEZIER_STROKE_BUTT_AA
                                 : (cap == FDC_CAP_SQUARE ? FDC_SDF_BEZIER_STROKE_SQUARE_AA : FDC_SDF_BEZIER_STROKE_AA);
      atx = d.f(0); aty = d.f(1); tox = fadd(d.f(0), d.f(2)); toy = fadd(d.f(1), d.f(3));
      break;
    }
    case FDC_OP_FILLED_QUAD: {
      float r[4];
      if (!atlas_lookup(a.atlas, kRectKey, r)) { empty = true; break; }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float2 v = xf_apply(xf, d.f(2 * k), d.f(2 * k + 1));
        q.x[k] = ceilf(v.x); q.y[k] = ceilf(v.y);
        cols[k] = d.u(3 + k);
      }
      have_rect_quad = false;
      uax = utx = fadd(r[0], fdiv(r[2], 2.0f));
      uay = uty = fadd(r[1], fdiv(r[3], 2.0f));
      mode = FDC_SDF_ATLAS;
      atlas_mode = true;
      break;
    }
    case FDC_OP_RECT: {
      float r[4];
      if (!atlas_lookup(a.atlas, kRectKey, r)) { empty = true; break; }
      uax = utx = fadd(r[0], fdiv(r[2], 2.0f));
      uay = uty = fadd(r[1], fdiv(r[3], 2.0f));
      cols[0] = cols[1] = cols[2] = cols[3] = d.u(3);
      mode = FDC_SDF_ATLAS;
      atlas_mode = true;
      atx = d.f(0); aty = d.f(1); tox = fadd(d.f(0), d.f(2)); toy = fadd(d.f(1), d.f(3));
      break;
    }
    default: empty = true; break;
  }

  int bx0 = 0, by0 = 0, bx1 = 0, by1 = 0;
  if (!empty) {
    if (have_rect_quad) quad_from_rect(xf, atx, aty, tox, toy, q);
    quad_bbox(q, bx0, by0, bx1, by1);
    bool aligned = have_rect_quad && q.x[0] == q.x[3] && q.x[1] == q.x[2] && q.y[0] == q.y[1] && q.y[2] == q.y[3];
    float X0 = q.x[3], Y0 = q.y[3], X1 = q.x[1], Y1 = q.y[1];
    if (aligned) {
      if (X0 == X1 || Y0 == Y1) empty = true;
      else {
        // s = (x + .5 - X0) / (X1 - X0)
        float iw = 1.0f / (X1 - X0), ih = 1.0f / (Y1 - Y0);
        p.su = iw; p.ou = (0.5f - X0) * iw;
        p.sv = ih; p.ov = (0.5f - Y0) * ih;
      }
    } else {
      flags |= PF_GENERAL;
      QuadGeom g;
#pragma unroll
      for (int k = 0; k < 4; k++) { g.vx[k] = (int)fminf(fmaxf(q.x[k], -1.0e6f), 1.0e6f); g.vy[k] = (int)fminf(fmaxf(q.y[k], -1.0e6f), 1.0e6f); }
      a.geoms[i] = g;
      p.su = __uint_as_float(i);
    }
    // clip: frame, band, enclosing texture masks
    int cx0 = bx0, cy0 = by0, cx1 = bx1, cy1 = by1;
    apply_clip_chain(a, rs.clip_draw, cx0, cy0, cx1, cy1);
    cx0 = max(cx0, 0); cx1 = min(cx1, a.frame.W);
    cy0 = max(cy0, a.frame.band_y0); cy1 = min(cy1, a.frame.band_y1);
    if (cx0 >= cx1 || cy0 >= cy1) empty = true;
    p.bx0 = (int16_t)cx0; p.by0 = (int16_t)cy0; p.bx1 = (int16_t)cx1; p.by1 = (int16_t)cy1;
    if (flags & PF_MASK_WIDE) {
      // own clipped bbox for the slow path's inside test (shade_prim); the bin bbox is widened below
      p.ix0 = (int16_t)cx0; p.iy0 = (int16_t)cy0; p.ix1 = (int16_t)cx1; p.iy1 = (int16_t)cy1;
    }

    if (!empty) {
      p.aa = rs.aa;
      bool solid = cols[0] == cols[1] && cols[1] == cols[2] && cols[2] == cols[3];
      if (solid) flags |= PF_SOLID;
#pragma unroll
      for (int k = 0; k < 4; k++) p.c[k] = cols[k];
      if (rs.rectmask && !(flags & PF_MASK_WRITE)) { flags |= PF_RECTMASK; p.aux = rs.rectmask & 0xFFFFu; }
      if (mode == FDC_SDF_DROP_SHADOW || mode == FDC_SDF_DROP_SHADOW_AA || mode == FDC_SDF_INSET_SHADOW) {
        float sigma = fmaxf(0.5f * p.factor, 0.5f);
        p.k = -0.5f * 1.4426950408889634f / (sigma * sigma);  // exp(-.5 z^2) = exp2(k * sd^2)
      }
      if (atlas_mode) {
        float as = (float)a.atlas.size;
        p.u0 = uax * as - 0.5f; p.du = (utx - uax) * as;
        p.v0 = uay * as - 0.5f; p.dv = (uty - uay) * as;
        if (mode == FDC_SDF_ATLAS && rs.subpixel_shift >= 0.0f) {
          flags |= PF_SUBPIXEL;
          p.u0 -= rs.subpixel_shift;  // atlasUv.x -= shift * atlasTexelSize.x  (atlas.frag:286-288)
        }
        if (aligned) {
          float ddx = fabsf(p.du * p.su), ddy = fabsf(p.dv * p.sv);  // texels per pixel
          if (mode == FDC_SDF_ATLAS) {
            float rho = fmaxf(ddx, ddy);
            p.k = rho > 0.0f ? log2f(rho) : -1000.0f;
          } else {
            // msdfScreenPxRange, atlas.frag:45-49
            float px_range = p.factor;
            p.k = fmaxf(0.5f * (px_range / ddx + px_range / ddy), 1.0f);
          }
        }
      }
      // SDF modes: direct pixel -> SDF-space mapping.  p.x = (s - .5) * 2 * qhx with s = x*su + ou.
      const bool sdf_rect = d.op() == FDC_OP_ROUNDED_RECT;
      if (sdf_rect && aligned) {
        p.u0 = 2.0f * p.qhx * p.su; p.du = (2.0f * p.ou - 1.0f) * p.qhx;
        p.v0 = -2.0f * p.qhy * p.sv; p.dv = -(2.0f * p.ov - 1.0f) * p.qhy;
      }
      const bool circ = !(flags & PF_ELLIPTICAL);
      const bool content = !(flags & PF_MASK_WRITE);
      // Atlas quads (glyphs, images, drawRect) magnified or 1:1 -- one bilinear fetch of level 0 -- take the fast path too.
      // MSDF / MTSDF quads always sample level 0 (textureLod) and carry one solid colour.
      const bool atlas_fast = FDC_FAST_TEX && aligned && ((mode == FDC_SDF_ATLAS && p.k <= 0.0f) || (mode >= FDC_SDF_MSDF && mode <= FDC_SDF_MTSDF_ANNULAR));
      // ClipAA mask writes (beginMask's clip shape) update a mask level instead of the pixel: same SDF, fast path too.
      const bool mask_fast = FDC_FAST_MASK && (flags & PF_MASK_WRITE) && sdf_rect && circ && mode == FDC_SDF_CLIP_AA && solid && fill_mode == 0;
      if (aligned && (content || mask_fast) && (FDC_FAST_MASK || !(flags & PF_RECTMASK)) &&
          ((atlas_fast && content) || (sdf_rect && circ &&
                          (mode == FDC_SDF_CLIP_AA || mode == FDC_SDF_ANNULAR_AA || mode == FDC_SDF_DROP_SHADOW)))) {
        // gradient colours as float coefficients of the pixel index (PrimExt)
        bool fast = true;
        if (fill_mode != 0) {
          PrimExt e;
          const float hs = 0.5f;
          if (fill_mode == 1) { e.ta = p.su; e.tb = 0.0f; e.tc = p.ou; }
          else if (fill_mode == 2) { e.ta = 0.0f; e.tb = p.sv; e.tc = p.ov; }
          else if (fill_mode == 3) { e.ta = hs * p.su; e.tb = hs * p.sv; e.tc = hs * (p.ou + p.ov); }
          else { e.ta = hs * p.su; e.tb = -hs * p.sv; e.tc = hs * (p.ou + 1.0f - p.ov); }
          const float mid = p.spread;  // already clamped to [0.01, 0.99]
          e.mid = mid;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const float c0 = (float)((cols[0] >> (8 * k)) & 255u), c1 = (float)((p.c_mid >> (8 * k)) & 255u),
                        c2 = (float)((p.c_stop >> (8 * k)) & 255u);
            e.a0[k] = c0; e.d0[k] = (c1 - c0) / mid;
            e.d1[k] = (c2 - c1) / (1.0f - mid); e.a1[k] = c1 - e.d1[k] * mid;
          }
          a.exts[i] = e;
        } else if (!solid) {
          // BL,BR,TR,TL at (s,t) = (0,1),(1,1),(1,0),(0,0): one affine function iff BL + TR == TL + BR per channel
          PrimExt e;
          e.ta = e.tb = e.tc = e.mid = 0.0f;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int bl = (cols[0] >> (8 * k)) & 255, br = (cols[1] >> (8 * k)) & 255, tr = (cols[2] >> (8 * k)) & 255,
                      tl = (cols[3] >> (8 * k)) & 255;
            if (bl + tr != tl + br) fast = false;
            const float ds = (float)(tr - tl), dt = (float)(bl - tl);
            e.a0[k] = (float)tl + ds * p.ou + dt * p.ov; e.d0[k] = ds * p.su; e.a1[k] = dt * p.sv; e.d1[k] = 0.0f;
          }
          if (fast) a.exts[i] = e;
        }
        if (fast) flags |= PF_FAST;
      }
      // Inner rect: the pixels where coverage is exactly 1 (ClipAA, DropShadow) or exactly 0 (inside an AnnularAA
      // stroke).  In the cross region |p| <= b - rmax the rounded-box SDF is d = max(|px|-bx, |py|-by).
      if (sdf_rect && aligned && circ && (content || mask_fast) && X1 > X0 && Y1 > Y0 && rs.aa > 0.0f &&
          (mode == FDC_SDF_CLIP_AA || mode == FDC_SDF_ANNULAR_AA || mode == FDC_SDF_DROP_SHADOW)) {
        // {d <= -D} of a rounded box is the box shrunk by D with corner radius max(r - D, 0); the largest axis-aligned
        // rect inside it is inset by D + (1 - 1/sqrt2) * max(r - D, 0) from the original edges (rmax is conservative).
        const float rmax = fmaxf(fmaxf(p.r0, p.r1), fmaxf(p.r2, p.r3));
        float D;
        bool ok = true;
        if (mode == FDC_SDF_CLIP_AA) D = 0.5f / rs.aa;                                   // coverage == 1
        else if (mode == FDC_SDF_DROP_SHADOW) D = -(fill_mode == 0 ? p.spread : 0.0f);   // sd = d - spread <= 0
        else { ok = p.factor >= 0.0f; D = p.factor + 0.5f / rs.aa; }                     // inside the stroke: coverage == 0
        float m = D + 0.29290f * fmaxf(rmax - D, 0.0f) + 0.02f;
        const float gx = 2.0f * p.qhx * p.su, gy = 2.0f * p.qhy * p.sv;  // d p / d pixel
        const float lox = X0 - 0.5f + (p.qhx - p.p2 + m) / gx, hix = X0 - 0.5f + (p.qhx + p.p2 - m) / gx;
        const float loy = Y0 - 0.5f + (p.qhy - p.p3 + m) / gy, hiy = Y0 - 0.5f + (p.qhy + p.p3 - m) / gy;
        const int ix0 = max((int)ceilf(lox + 1e-3f), cx0), ix1 = min((int)floorf(hix - 1e-3f) + 1, cx1);
        const int iy0 = max((int)ceilf(loy + 1e-3f), cy0), iy1 = min((int)floorf(hiy - 1e-3f) + 1, cy1);
        if (ok && ix0 < ix1 && iy0 < iy1) {
          flags |= PF_INNER;
          if (mode == FDC_SDF_ANNULAR_AA) flags |= PF_INNER_EMPTY;
          p.ix0 = (int16_t)ix0; p.iy0 = (int16_t)iy0; p.ix1 = (int16_t)ix1; p.iy1 = (int16_t)iy1;
          // occluder: opaque unmasked ClipAA fill -- nothing painted before it survives inside the inner rect
          // (PF_FAST only: the fast path stores the colour directly for such visits, independent of the destination)
          if (mode == FDC_SDF_CLIP_AA && (flags & PF_FAST) && (flags & PF_DEPTH_MASK) == 0 && !(flags & (PF_RECTMASK | PF_MASK_WRITE))) {
            uint32_t amin = min(min(cols[0] >> 24, cols[1] >> 24), min(cols[2] >> 24, cols[3] >> 24));
            if (fill_mode != 0) amin = min(amin, min(p.c_mid >> 24, p.c_stop >> 24));
            if (amin == 255u) flags |= PF_OCCLUDER;
          }
        }
      }
      if (solid && fill_mode == 0) {
        // the colour as floats 0..255 so the shade kernel never converts
        p.c[0] = __float_as_uint((float)(cols[0] & 255u));
        p.c[1] = __float_as_uint((float)((cols[0] >> 8) & 255u));
        p.c[2] = __float_as_uint((float)((cols[0] >> 16) & 255u));
        p.c[3] = __float_as_uint((float)(cols[0] >> 24));
      } else {
        flags &= ~PF_SOLID;
      }
    }
  }
  // First draw of a mask level that holds several draws: GL cleared the whole mask texture at beginMask
  // (glcontext.nim:1901-1902) and content under such a level is not clipped to one primitive's bbox, so the level
  // must read 0 wherever none of its draws lands -- bin this primitive over the parent's whole clip box.
  int wx0 = 0, wy0 = 0, wx1 = 0, wy1 = 0;
  const bool wide = (rs.flags & PF_MASK_WIDE) != 0;
  if (wide) {
    wx0 = 0; wx1 = a.frame.W; wy0 = a.frame.band_y0; wy1 = a.frame.band_y1;
    apply_clip_chain(a, rs.clip_draw, wx0, wy0, wx1, wy1);
  }
  if (empty && wide && wx0 < wx1 && wy0 < wy1) {
    memset(&p, 0, sizeof(p));
    flags = (rs.flags & (PF_MASK_WRITE | PF_MASK_BEGIN | PF_MASK_WIDE | PF_DEPTH_MASK)) | PF_CLEAR_ONLY | (uint32_t)FDC_SDF_CLIP_AA;
    p.bx0 = (int16_t)wx0; p.by0 = (int16_t)wy0; p.bx1 = (int16_t)wx1; p.by1 = (int16_t)wy1;
  } else if (empty) {
    flags = PF_EMPTY;
    p.bx0 = p.by0 = p.bx1 = p.by1 = 0;
  } else {
    flags |= (uint32_t)mode & PF_MODE_MASK;
    flags |= ((uint32_t)fill_mode << PF_FILLMODE_SHIFT) & PF_FILLMODE_MASK;
    if (wide) {
      if (flags & PF_INNER) {
        // fast mask write: ix0..iy1 hold the inner rect and the fast path tests the quad in SDF space -- keep both
      }
      p.bx0 = (int16_t)wx0; p.by0 = (int16_t)wy0; p.bx1 = (int16_t)wx1; p.by1 = (int16_t)wy1;
    }
  }
  p.mode_flags = flags;
  a.prims[i] = p;
  PrimBin pb;
  pb.bx0 = p.bx0; pb.by0 = p.by0; pb.bx1 = p.bx1; pb.by1 = p.by1;
  pb.mode_flags = flags; pb.aux = p.aux;
  pb.ix0 = p.ix0; pb.iy0 = p.iy0; pb.ix1 = p.ix1; pb.iy1 = p.iy1;
  pb.pad_[0] = pb.pad_[1] = 0u;
  a.prim_bins[i] = pb;
}

// Tile-band partitions: most records of a frame land outside a rank's band (7/8 of them on 8 GPUs).  A rounded rect is
// outside exactly when the full setup would find its clipped bbox empty; this restates just that part -- the same
// transform, ceil and bbox arithmetic -- so the rest of the setup is skipped for it.
__device__ bool outside_band(const SetupArgs& a, uint32_t i, const uint32_t* rec, uint32_t* call_index, int run_lo, int run_hi) {
  const uint32_t di = a.first + i;
  const RunState rs = a.runs[find_run_in(a.runs, run_lo, run_hi, di)];
  *call_index = rs.call_index + (di - rs.first_draw);
  if (rs.flags & PF_MASK_WIDE) return false;  // carries a clear over the parent's clip box whatever its own quad is
  float x, y, w, h;
  if (rs.compact) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(&a.rects64[rs.src_off + (di - rs.first_draw)]));
    x = r.x; y = r.y; w = r.z; h = r.w;
  } else {
    const DrawView d{rec};
    if (d.op() != FDC_OP_ROUNDED_RECT) return false;
    x = d.f(0); y = d.f(1); w = d.f(2); h = d.f(3);
  }
  if (w <= 0.0f || h <= 0.0f) return true;
  QuadPos q;
  quad_from_rect(a.xforms[rs.xform], x, y, fadd(x, w), fadd(y, h), q);
  int bx0, by0, bx1, by1;
  quad_bbox(q, bx0, by0, bx1, by1);
  return max(bx0, 0) >= min(bx1, a.frame.W) || max(by0, a.frame.band_y0) >= min(by1, a.frame.band_y1);
}

#ifndef FDC_SETUP_MIN_BLOCKS
#define FDC_SETUP_MIN_BLOCKS 1
#endif
template <bool kBanded>
__global__ void __launch_bounds__(128, FDC_SETUP_MIN_BLOCKS) prim_setup_kernel(SetupArgs a) {
  // The CTA's 128 records (16 KB) come in with coalesced 16-byte loads and are read back from shared memory at a
  // 33-word stride (no bank conflicts): one record per thread straight from global memory was a 128-byte-stride
  // gather that left the kernel waiting on loads (issue slot utilisation 0.17, profiles/r01_binning.md).
  __shared__ uint32_t s_draw[128][33];
  __shared__ uint32_t s_keep[128];
  __shared__ uint32_t s_nkeep;
  const uint32_t first = blockIdx.x * 128u;
  const uint32_t n_here = min(128u, a.count - first);
  // the binning kernels that follow start from zeroed counters (this replaces a memset node per segment)
  if (blockIdx.x == 0 && (int)threadIdx.x < a.zero_counters) a.counters[threadIdx.x] = 0u;
  if (blockIdx.x == 0 && a.zero_counters == (int)kNumCounters)
    for (int r = threadIdx.x; r < a.n_row_cost; r += 128) a.row_cost[r] = 0u;
  // Runs of the CTA's first and last record (most CTAs sit inside one run), found by every thread for itself: the
  // loads are uniform, so this costs one broadcast per step and saves the barrier a single searching thread needed.
  const int run_lo = find_run(a.runs, a.n_runs, a.first + first);
  const int run_hi = find_run_in(a.runs, run_lo, a.n_runs - 1, a.first + first + n_here - 1u);
  bool all_compact = true;
  for (int r = run_lo; r <= run_hi; r++) all_compact = all_compact && a.runs[r].compact != 0;
  if (kBanded) {
    if (threadIdx.x == 0) s_nkeep = 0;
    __syncthreads();
  }
  if (!all_compact) {  // compact runs have no fdc_call records to stage (their slots in `draws` are unused)
    const uint4* src = reinterpret_cast<const uint4*>(a.draws + a.first + first);
    for (uint32_t k = threadIdx.x; k < n_here * 8u; k += 128u) {
      const uint4 v = __ldg(src + k);
      uint32_t* dst = &s_draw[k >> 3][(k & 7u) * 4u];
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __syncthreads();
  }
  if (!kBanded) {
    if (threadIdx.x < n_here) setup_record(a, first + threadIdx.x, s_draw[threadIdx.x], run_lo, run_hi);
    return;
  }
  // Band partition: weed out the records that land outside the band first (they only get their PF_EMPTY marker), then
  // the CTA's threads share the survivors -- dense warps instead of a few live lanes in every warp.
  bool keep = false;
  if (threadIdx.x < n_here) {
    uint32_t call_index;
    keep = !outside_band(a, first + threadIdx.x, s_draw[threadIdx.x], &call_index, run_lo, run_hi);
    if (!keep) {
      const uint32_t i = first + threadIdx.x;
      a.prim_call[i] = call_index;
      reinterpret_cast<int4*>(&a.prims[i])[6] = make_int4(0, 0, (int)PF_EMPTY, 0);
      int4* pb = reinterpret_cast<int4*>(&a.prim_bins[i]);
      pb[0] = make_int4(0, 0, (int)PF_EMPTY, 0);
      pb[1] = make_int4(0, 0, 0, 0);
    }
  }
  const uint32_t km = __ballot_sync(0xFFFFFFFFu, keep);
  uint32_t wbase = 0;
  if ((threadIdx.x & 31) == 0 && km) wbase = atomicAdd(&s_nkeep, (uint32_t)__popc(km));
  wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
  if (keep) s_keep[wbase + __popc(km & ((1u << (threadIdx.x & 31)) - 1u))] = threadIdx.x;
  __syncthreads();
  const uint32_t n_keep = s_nkeep;
  for (uint32_t k = threadIdx.x; k < n_keep; k += 128u) {
    const uint32_t j = s_keep[k];
    setup_record(a, first + j, s_draw[j], run_lo, run_hi);
  }
}

void launch_prim_setup(const SetupArgs& a, cudaStream_t stream) {
  if (a.count == 0) return;
  const bool banded = a.frame.band_y0 > 0 || a.frame.band_y1 < a.frame.H;
  if (banded) prim_setup_kernel<true><<<(a.count + 127) / 128, 128, 0, stream>>>(a);
  else prim_setup_kernel<false><<<(a.count + 127) / 128, 128, 0, stream>>>(a);
}

// ------------------------------------------------------------------------------------------------ coarse binning
// Coarse rect of a primitive in band-local coarse-bin coordinates, inclusive, packed x0 | y0<<8 | x1<<16 | y1<<24.
// Empty primitives get x0 = 255 > x1 = 0 so they never match.
__device__ __forceinline__ uint32_t coarse_rect(const PrimBin* prims, uint32_t idx, uint32_t n, const FrameView& f) {
  if (idx >= n) return 0x000000FFu;
  const int4 q6 = __ldg(reinterpret_cast<const int4*>(&prims[idx]));
  if ((uint32_t)q6.z & PF_EMPTY) return 0x000000FFu;
  int bx0 = (int16_t)(q6.x & 0xFFFF), by0 = (int16_t)(q6.x >> 16), bx1 = (int16_t)(q6.y & 0xFFFF), by1 = (int16_t)(q6.y >> 16);
  const int cpx = kTileW * kCoarse, cpy = kTileH * kCoarse;
  int oy = f.cty0 * kTileH;
  uint32_t x0 = (uint32_t)(bx0 / cpx), x1 = (uint32_t)((bx1 - 1) / cpx);
  uint32_t y0 = (uint32_t)((by0 - oy) / cpy), y1 = (uint32_t)((by1 - 1 - oy) / cpy);
  return x0 | (y0 << 8) | (x1 << 16) | (y1 << 24);
}

__device__ __forceinline__ bool rect_hits(uint32_t r, uint32_t bx, uint32_t by) {
  return (r & 255u) <= bx && bx <= ((r >> 16) & 255u) && ((r >> 8) & 255u) <= by && by <= (r >> 24);
}

// 32x32 bit-matrix transpose across a warp: lane i passes row i (bit k = A[i][k]) and receives column i (bit k = A[k][i]).
// Five butterfly steps swapping off-diagonal blocks of size 16, 8, 4, 2, 1.  (Used by the fine binner.)
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, x, j);
    x = (lane & j) ? (((y >> j) & m) | (x & ~m)) : ((x & m) | ((y & m) << j));
  }
  return x;
}

// ---- coarse level, r02: (primitive, bin) PAIRS instead of a bitmap walk.
// A primitive covers a small rectangle of 128x128-px bins (1.8 on average at 4K).  The r01 kernels marked bins in a
// bitmap and transposed every marked 32-bin word across the warp -- ~100 warp instructions per primitive, independent
// of how few pairs there were.  Now a CTA enumerates exactly the pairs of its chunk of primitives, in primitive order:
// an exclusive scan of the per-primitive pair counts, then every warp takes an equal, contiguous share of the pair
// sequence 32 pairs at a time (a binary search finds each pair's owner), so a full-frame primitive with hundreds of
// pairs is spread over all warps instead of serialising one.  Counting is a shared-memory histogram; the stable rank
// of a pair inside a round is popc(match_any(bin) & lower lanes); rounds and warps are ordered by running offsets
// kept per (bin, warp).  Cost follows the number of pairs -- a band of an 8-GPU partition has 1/8 of them.
constexpr int kSlots = 1024;  // coarse bins one CTA handles: a range of whole bin rows
constexpr int kWarps = kChunk / 32;

struct ChunkPairs {
  uint32_t* excl;    // [kChunk] exclusive scan of the pair counts
  uint32_t* rect;    // [kChunk] x0 | y0 << 8 | w << 16 of each primitive's bin rectangle clipped to the CTA's row range
  uint32_t total;    // pairs of the chunk
  uint32_t lo, hi;   // this warp's share [lo, hi)
};
// All threads of the CTA call this; ends with a __syncthreads().
__device__ __forceinline__ ChunkPairs chunk_pairs(uint32_t rect, int row0, int row1, uint32_t* s_excl, uint32_t* s_rect,
                                                  uint32_t* s_wtot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = rect & 255u, y0 = (rect >> 8) & 255u, x1 = (rect >> 16) & 255u, y1 = rect >> 24;
  const int ya = max(y0, row0), yb = min(y1, row1 - 1);
  const int w = x1 - x0 + 1, h = yb - ya + 1;
  const uint32_t k = (w > 0 && h > 0) ? (uint32_t)(w * h) : 0u;  // empty primitives carry x0 = 255 > x1 = 0
  s_rect[threadIdx.x] = (uint32_t)x0 | ((uint32_t)(ya & 255) << 8) | ((uint32_t)(w & 0xFFFF) << 16);
  uint32_t incl = k;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int j = 0; j < kWarps; j++) {
    const uint32_t t = s_wtot[j];
    if (j < warp) wbase += t;
    total += t;
  }
  s_excl[threadIdx.x] = wbase + incl - k;
  __syncthreads();
  ChunkPairs cp;
  cp.excl = s_excl;
  cp.rect = s_rect;
  cp.total = total;
  const uint32_t rounds = (total + 31u) >> 5, per_warp = (rounds + kWarps - 1) / kWarps;
  cp.lo = min(total, (uint32_t)warp * per_warp * 32u);
  cp.hi = min(total, cp.lo + per_warp * 32u);
  return cp;
}
// Pair number p of the chunk (p < total): its owner (index inside the chunk) and its bin.
__device__ __forceinline__ void chunk_pair(const ChunkPairs& cp, uint32_t p, int& owner, int& bx, int& by) {
  int o = 0;
#pragma unroll
  for (int step = kChunk / 2; step >= 1; step >>= 1)
    if (cp.excl[o + step] <= p) o += step;  // largest index with excl <= p; ties: the highest wins, skipping empty ones
  owner = o;
  const uint32_t r = cp.rect[o];
  const int j = (int)(p - cp.excl[o]);
  const int w = max((int)(r >> 16), 1);
  const int q = (int)(((float)j + 0.5f) * (1.0f / (float)w));  // j / w, exact for j < 2^16
  bx = (int)(r & 255u) + (j - q * w);
  by = (int)((r >> 8) & 255u) + q;
}

// One kernel: count, place and scatter in the same CTA.  (Round 1 and the first half of round 2 had three -- count,
// a grid-wide scan, scatter -- because every bin's list was contiguous.)  Nothing requires that: the fine binner reads a
// bin's entries in CHUNK order, so a bin's list may as well be one segment per chunk, anywhere in the coarse list, as long
// as a table says where.  A CTA counts its chunk's pairs per bin (shared-memory histogram, one 16-bit counter per (bin,
// warp)), turns the counts into running offsets over the warps, scans its own bins, reserves its region of the list with
// one atomicAdd (placement is not observable, order is: segments are read in chunk order and written in primitive
// order), writes seg[bin][chunk] = (start, count) and scatters: position = segment start + pairs of earlier warps +
// pairs of earlier rounds of this warp + rank inside the round.  The fine binner maps "entry k of the bin" to (chunk,
// offset) with a binary search over the scanned counts of its bin's row of the table.
__global__ void __launch_bounds__(kChunk) coarse_pairs_kernel(const PrimBin* __restrict__ prims, uint32_t n, FrameView f, int rows_per_cta,
                                                              uint2* __restrict__ seg, uint32_t* __restrict__ coarse_list,
                                                              uint32_t coarse_cap, uint32_t* __restrict__ counters) {
  // per (bin, warp): first the pair count, then (after the prefix over warps) the running offset inside the chunk's
  // segment of the bin.  16 bits each (a chunk puts at most kChunk pairs into a bin), two warps per word; [word][bin] so
  // that consecutive threads hit consecutive banks.
  __shared__ uint32_t s_pos[kWarps / 2][kSlots];
  __shared__ uint32_t s_gbase[kSlots];
  __shared__ uint32_t s_excl[kChunk], s_rect[kChunk], s_wtot[kWarps];
  __shared__ uint32_t s_region;
  const uint32_t chunk = blockIdx.x;
  const int row0 = blockIdx.y * rows_per_cta, row1 = min(row0 + rows_per_cta, f.cby);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int used = (row1 - row0) * f.cbx;
  const uint32_t rect = coarse_rect(prims, chunk * kChunk + threadIdx.x, n, f);
  for (int k = 0; k < kWarps / 2; k++)
    for (int sl = threadIdx.x; sl < used; sl += blockDim.x) s_pos[k][sl] = 0;
  const ChunkPairs cp = chunk_pairs(rect, row0, row1, s_excl, s_rect, s_wtot);
  const int word = warp >> 1, shift = (warp & 1) * 16;
  for (uint32_t base = cp.lo; base < cp.hi; base += 32) {
    const uint32_t p = base + lane;
    if (p < cp.hi) {
      int owner, bx, by;
      chunk_pair(cp, p, owner, bx, by);
      atomicAdd(&s_pos[word][(by - row0) * f.cbx + bx], 1u << shift);
    }
  }
  if (threadIdx.x == 0) {
    const uint32_t at = cp.total ? atomicAdd(&counters[kCntCoarseTotal], cp.total) : 0u;
    if (cp.total) atomicMax(&counters[kCntMaxCoarse], at + cp.total);  // the last reservation's end = pairs of the segment
    if (at + cp.total > coarse_cap) { atomicOr(&counters[kCntOverflow], 1u); atomicOr(&counters[kCntStickyOverflow], 1u); }
    s_region = at;
  }
  __syncthreads();
  for (int sl = threadIdx.x; sl < used; sl += blockDim.x) {  // counts -> exclusive prefix over the warps; bin total
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < kWarps / 2; k++) {
      const uint32_t v = s_pos[k][sl];
      const uint32_t lo = v & 0xFFFFu, hi = v >> 16;
      s_pos[k][sl] = run | ((run + lo) << 16);
      run += lo + hi;
    }
    s_gbase[sl] = run;
  }
  __syncthreads();
  // exclusive scan over the CTA's bins: thread t owns slots 2t and 2t + 1 (kSlots = 2 * kChunk)
  static_assert(kSlots == 2 * kChunk, "one pair of bin slots per thread");
  const int sl0 = 2 * (int)threadIdx.x;
  const uint32_t c0 = sl0 < used ? s_gbase[sl0] : 0u, c1 = sl0 + 1 < used ? s_gbase[sl0 + 1] : 0u;
  uint32_t incl = c0 + c1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int j = 0; j < kWarps; j++)
    if (j < warp) wbase += s_wtot[j];
  const uint32_t region = s_region;
  const bool fits = region + cp.total <= coarse_cap;
  const uint32_t e0 = region + wbase + incl - c0 - c1, e1 = e0 + c0;
  const size_t n_chunks = gridDim.x;
  if (sl0 < used) {
    s_gbase[sl0] = e0;
    seg[(size_t)(row0 * f.cbx + sl0) * n_chunks + chunk] = make_uint2(e0, fits ? c0 : 0u);
  }
  if (sl0 + 1 < used) {
    s_gbase[sl0 + 1] = e1;
    seg[(size_t)(row0 * f.cbx + sl0 + 1) * n_chunks + chunk] = make_uint2(e1, fits ? c1 : 0u);
  }
  __syncthreads();
  if (!fits) return;  // overflow: host regrows and re-runs the frame
  const uint32_t first = chunk * kChunk;
  for (uint32_t base = cp.lo; base < cp.hi; base += 32) {
    const uint32_t p = base + lane;
    const bool active = p < cp.hi;
    int owner = 0, bx = 0, by = row0;
    if (active) chunk_pair(cp, p, owner, bx, by);
    const int sl = (by - row0) * f.cbx + bx;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, active ? (uint32_t)sl : (0x80000000u | (uint32_t)lane));
    if (active) {
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      const uint32_t off = (s_pos[word][sl] >> shift) & 0xFFFFu;
      coarse_list[s_gbase[sl] + off + rank] = first + (uint32_t)owner;
    }
    __syncwarp();
    if (active && (peers & ((1u << lane) - 1u)) == 0u) atomicAdd(&s_pos[word][sl], (uint32_t)__popc(peers) << shift);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ fine binning
// One CTA per coarse bin.  The bin's list is staged in shared memory 1024 entries at a time (each thread gathers the
// bbox, flags and inner rect of 4 entries), then warp w owns tile row w: per 32 staged entries one ballot per tile
// column gives counts and stable ranks.  Counts -> CTA scan -> one atomicAdd reserves the bin's slice of the tile list
// -> second walk scatters 8-byte TileEntry records that already carry, for each of the tile's eight 8x4 blocks,
// "bbox overlaps" and "inside the inner rect" bits, so the shading warps never touch the primitive for culling.
constexpr int kStage = 1024;  // = 4 entries per thread of a 256-thread CTA
static_assert(kTileW == 16 && kTileH == 16 && kCoarse == 8, "fine_bin_kernel shifts assume 16x16 tiles, 8x8 tiles per bin");
static_assert(TE_OV_SHIFT == 0 && TE_FULL_SHIFT == 8, "fine_bin_kernel builds the overlap and inner-rect bytes as one 16-bit value");

// bits [a..b] of a 32-bit word (empty when b < a)
__device__ __forceinline__ uint32_t bit_range(int a, int b) {
  if (b < a) return 0u;
  const uint32_t hi = b >= 31 ? 0xFFFFFFFFu : ((2u << b) - 1u);
  return hi & ~((1u << a) - 1u);
}
// Blocks of `bs` pixels starting at p0: which of them does [lo,hi) overlap / fully cover (a block's extent is clipped to
// `limit`, the frame edge, when deciding "covered")?
__device__ __forceinline__ uint32_t blocks_overlapped(int lo, int hi, int p0, int shift, int nblk) {
  return bit_range(max((lo - p0) >> shift, 0), min((hi - 1 - p0) >> shift, nblk - 1));
}
__device__ __forceinline__ uint32_t blocks_covered(int lo, int hi, int p0, int shift, int nblk, int limit) {
  const int bs = 1 << shift;
  const int a = max((lo - p0 + bs - 1) >> shift, 0);
  const int b = hi >= limit ? min((limit - 1 - p0) >> shift, nblk - 1) : min(((hi - p0) >> shift) - 1, nblk - 1);
  return bit_range(a, b);
}
// 4 row bits -> each duplicated into a pair (block = row*2 + col)
__device__ __forceinline__ uint32_t spread_rows(uint32_t r) {
  uint32_t x = (r | (r << 2)) & 0x33u;
  x = (x | (x << 1)) & 0x55u;
  return x * 3u;
}
__device__ __forceinline__ uint32_t spread_cols(uint32_t c) { return (c & 1u) * 0x55u | (c >> 1) * 0xAAu; }

constexpr uint32_t kInfoEmptyInner = 1u << 30;  // staging-only markers, stripped before the entry is written
constexpr uint32_t kInfoBegin = 1u << 31;

// tile-row hit bits (8) from 32 block-row bits: bit r = "nibble r is non-zero"
__device__ __forceinline__ uint32_t nibbles_nonzero(uint32_t x) {
  x |= x >> 1;
  x |= x >> 2;
  x &= 0x11111111u;                       // bit 4r
  x = (x | (x >> 3)) & 0x03030303u;       // bits 8j, 8j+1
  x = (x | (x >> 6)) & 0x000F000Fu;       // bits 16j .. 16j+3
  return (x | (x >> 12)) & 0xFFu;
}
// tile-column hit bits (8) from 16 block-column bits: bit c = "pair c is non-zero"
__device__ __forceinline__ uint32_t pairs_nonzero(uint32_t x) {
  x = (x | (x >> 1)) & 0x5555u;
  x = (x | (x >> 1)) & 0x3333u;
  x = (x | (x >> 2)) & 0x0F0Fu;
  return (x | (x >> 4)) & 0xFFu;
}
// bit i of a nibble -> bit 8i
__device__ __forceinline__ uint32_t spread_nibble_to_bytes(uint32_t r) {
  return (r & 1u) | ((r & 2u) << 7) | ((r & 4u) << 14) | ((r & 8u) << 21);
}

// Everything the shade kernel needs per (tile, primitive), derived once per staged coarse entry.
// four nibbles of a 16-bit value -> the low nibbles of four bytes
__device__ __forceinline__ uint32_t nibbles_to_bytes(uint32_t x) {
  x = (x | (x << 8)) & 0x00FF00FFu;
  return (x | (x << 4)) & 0x0F0F0F0Fu;
}
// four bit pairs of an 8-bit value -> bits 0..1 of four bytes
__device__ __forceinline__ uint32_t pairs_to_bytes(uint32_t x) {
  x = (x | (x << 12)) & 0x000F000Fu;
  return (x | (x << 6)) & 0x03030303u;
}

__device__ __forceinline__ void stage_entry(uint32_t k, uint32_t pid, const int4& q6, const int2& ir, int px0, int py0, const FrameView& f,
                                            uint32_t* s_pid, uint32_t* s_rm0, uint32_t* s_rm1, uint32_t* s_cm0, uint32_t* s_cm1,
                                            uint32_t* s_info, uint32_t* s_lo, uint32_t* s_hi) {
  const uint32_t fl = (uint32_t)q6.z;
  const int bx0 = (int16_t)(q6.x & 0xFFFF), by0 = (int16_t)(q6.x >> 16), bx1 = (int16_t)(q6.y & 0xFFFF), by1 = (int16_t)(q6.y >> 16);
  uint32_t cols = blocks_overlapped(bx0, bx1, px0, 3, 16);
  uint32_t rows_ov = blocks_overlapped(by0, by1, py0, 2, 32), rows_full = 0;
  if (fl & PF_INNER) {
    const int ix0 = (int16_t)(ir.x & 0xFFFF), iy0 = (int16_t)(ir.x >> 16), ix1 = (int16_t)(ir.y & 0xFFFF), iy1 = (int16_t)(ir.y >> 16);
    cols |= blocks_covered(ix0, ix1, px0, 3, 16, f.W) << 16;
    rows_full = blocks_covered(iy0, iy1, py0, 2, 32, f.H);
  }
  const uint32_t mode = fl & PF_MODE_MASK;
  uint32_t kind = mode == FDC_SDF_CLIP_AA ? 0u : (mode == FDC_SDF_ANNULAR_AA ? 1u : 2u), tex = 0u;
  if (mode == FDC_SDF_ATLAS) { tex = TE_TEX; kind = 0u; }
  else if (mode == FDC_SDF_MSDF) { tex = TE_TEX; kind = 1u; }
  else if (mode == FDC_SDF_MTSDF) { tex = TE_TEX; kind = 2u; }
  else if (mode == FDC_SDF_MSDF_ANNULAR) { tex = TE_TEX | TE_GRAD3; kind = 1u; }
  else if (mode == FDC_SDF_MTSDF_ANNULAR) { tex = TE_TEX | TE_GRAD3; kind = 2u; }
  uint32_t info = ((fl & PF_FAST) ? TE_FAST : 0u) | ((fl & PF_SOLID) ? TE_SOLID : 0u) | tex |
                  ((fl & PF_FILLMODE_MASK) ? TE_GRAD3 : 0u) | (kind << TE_KIND_SHIFT) |
                  ((fl & PF_OCCLUDER) ? TE_OCCLUDER : 0u) | (((fl & PF_DEPTH_MASK) >> PF_DEPTH_SHIFT) << TE_DEPTH_SHIFT);
  if (fl & PF_MASK_WRITE) info |= TE_MASKW;
  if (fl & PF_MASK_BEGIN) info |= TE_MASKB;
  if (fl & PF_RECTMASK) info |= TE_RECTMASK;
  if (fl & PF_INNER_EMPTY) info |= kInfoEmptyInner;
  if (fl & PF_MASK_BEGIN) info |= kInfoBegin;
  const uint32_t tr = nibbles_nonzero(rows_ov), tc = pairs_nonzero(cols & 0xFFFFu);
  s_pid[k] = pid;
  // What the write loop needs per (entry, tile), pre-packed so that a tile's share is one byte of one word:
  //   row words (tile rows 0-3 / 4-7): byte i = overlap bits of the tile row's four block rows | inner-rect bits << 4
  //   column words (tile columns 0-3 / 4-7): byte i = overlap bits of the tile column's two block columns | inner bits << 4
  s_rm0[k] = nibbles_to_bytes(rows_ov & 0xFFFFu) | (nibbles_to_bytes(rows_full & 0xFFFFu) << 4);
  s_rm1[k] = nibbles_to_bytes(rows_ov >> 16) | (nibbles_to_bytes(rows_full >> 16) << 4);
  s_cm0[k] = pairs_to_bytes(cols & 0xFFu) | (pairs_to_bytes((cols >> 16) & 0xFFu) << 4);
  s_cm1[k] = pairs_to_bytes((cols >> 8) & 0xFFu) | (pairs_to_bytes(cols >> 24) << 4);
  s_info[k] = info;
  s_lo[k] = tc * spread_nibble_to_bytes(tr & 15u);
  s_hi[k] = tc * spread_nibble_to_bytes(tr >> 4);
}

constexpr int kGroups = kStage / 32;
constexpr int kSegBlock = 1024;  // chunks whose segments of a bin are addressed at a time (= 512 k primitives)
static_assert(kGroups * 64 == 2 * kSegBlock, "segment table block and s_gpos share storage");
#ifndef FDC_DIRECT_FINE_LIMIT
#define FDC_DIRECT_FINE_LIMIT 262144
#endif
#ifndef FDC_FINE_SPLIT_BIG
#define FDC_FINE_SPLIT_BIG 0
#endif
#ifndef FDC_FINE_SPLIT_BINS
#define FDC_FINE_SPLIT_BINS 296
#endif
constexpr int kFineSplitBins = FDC_FINE_SPLIT_BINS;  // at most this many bins: two CTAs per bin (148 SMs x 4 CTAs = 592 slots)
constexpr size_t kDirectFineLimit = FDC_DIRECT_FINE_LIMIT;  // primitives x coarse bins below which the coarse pass is skipped

// One CTA per coarse bin (8x8 tiles).  Staged entries are handled 32 at a time ("groups"), one group per warp: every
// lane turns its entry into a 64-bit tile-hit mask (tile rows x tile columns), two 32x32 bit transposes across the
// warp turn "tiles per entry" into "entries per tile", and lane t then owns tiles t and 32+t of the bin: the count is
// a popc, the emission order is the order of the set bits.  (Before: warp w owned tile row w and every warp walked
// every entry with one ballot per tile column -- 8x the instructions for the same lists; profiles/r01_binning.md.)
#ifndef FDC_FINE_MIN_BLOCKS
#define FDC_FINE_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(256, FDC_FINE_MIN_BLOCKS) fine_bin_kernel(const PrimBin* __restrict__ prims, FrameView f,
                                                       const uint2* __restrict__ seg,
                                                       int n_chunks, const uint32_t* __restrict__ coarse_list, uint32_t coarse_cap,
                                                       uint32_t* __restrict__ tile_start, uint32_t* __restrict__ tile_count,
                                                       TileEntry* __restrict__ tile_list, uint32_t tile_cap,
                                                       uint32_t* __restrict__ counters, uint32_t* __restrict__ row_cost,
                                                       uint32_t n_direct, int split_big) {
  // n_direct != 0: small scene, no coarse pass was run -- every bin stages primitives 0..n_direct-1 themselves (those
  // that miss the bin get an empty tile mask); otherwise the bin's coarse list.
  // per staged coarse entry, computed once: 16 block columns (8 px) and 32 block rows (4 px) of this 128x128-px bin
  __shared__ uint32_t s_pid[kStage];
  __shared__ uint32_t s_rm0[kStage], s_rm1[kStage];  // block-row bits per tile row (stage_entry)
  __shared__ uint32_t s_cm0[kStage], s_cm1[kStage];  // block-column bits per tile column
  __shared__ uint32_t s_info[kStage];
  // tile-hit masks: tile rows 0-3 / 4-7, bit = row*8 + column.  Staged per entry; the counting step replaces every group
  // of 32 words by its transpose (word t of the group = the group's entries that hit tile t), which the write loop reads.
  __shared__ uint32_t s_lo[kStage], s_hi[kStage];
  __shared__ uint8_t s_gcnt[kGroups][64];          // entries of group g in tile t
  // s_gpos: where group g's entries of tile t go in the tile list (written after a stage is counted, read by the write
  // loop).  The same 8 KB hold, while a stage is being gathered, the bin's row of the segment table for up to kSegBlock
  // chunks: scanned counts and segment starts (coarse_pairs_kernel).
  __shared__ union {
    uint32_t gpos[kGroups][64];
    struct { uint32_t excl[kSegBlock]; uint32_t start[kSegBlock]; } seg;
  } s_u;
  __shared__ uint32_t s_wsum[8];
  __shared__ uint32_t s_run[64];                   // pass 0: running tile counts; pass 1: write cursors
  __shared__ uint32_t s_base[64];
  __shared__ uint32_t s_cls[64];                   // != 0: the tile holds something the shade kernel's lean loop cannot take
  __shared__ uint32_t s_alloc;
  if (!n_direct && counters[kCntCoarseTotal] > coarse_cap) return;
  const int b = blockIdx.x;
  const int cbx_i = b % f.cbx, cby_i = b / f.cbx;
  // Where the bin's entries come from: primitives 0..n_direct-1 themselves, or one segment of the coarse list per chunk
  // (coarse_pairs_kernel), addressed kSegBlock chunks at a time.
  const bool use_seg = !n_direct;
  const uint2* seg_row = use_seg ? seg + (size_t)b * (size_t)n_chunks : nullptr;
  const int n_blocks = use_seg ? (n_chunks + kSegBlock - 1) / kSegBlock : 1;
  uint32_t blk_total = n_direct;  // entries of the current block of chunks (segments: known once the block is loaded)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_x0 = cbx_i * kCoarse, tile_y0 = f.cty0 + cby_i * kCoarse;
  const int px0 = tile_x0 * kTileW, py0 = tile_y0 * kTileH;
  bool single = !use_seg && blk_total <= (uint32_t)kStage;  // the whole bin fits one stage: pass 1 reuses what pass 0 staged
  if (threadIdx.x < 64) { s_run[threadIdx.x] = 0; s_cls[threadIdx.x] = 0; }
  // Few bins (a band of a multi-GPU partition, a small frame) leave most SMs idle with one CTA per bin: the launch then
  // has two CTAs per bin, each building the lists of one half of the bin's tiles (tile rows 0-3 / 4-7).
  // With many bins (split_big, grid.y = 2 as well) only the bins that need more than one stage are shared by two CTAs
  // -- they are the launch's critical path, each stage is gathered twice -- and the second CTA of any other bin leaves.
  bool do_lo = gridDim.y == 1 || blockIdx.y == 0, do_hi = gridDim.y == 1 || blockIdx.y == 1;
  if (split_big && !use_seg && single) {  // (direct mode: the whole "list" is known up front)
    if (blockIdx.y == 1) return;
    do_lo = do_hi = true;
  }

  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) {
      __syncthreads();
      if (warp == 0) {
        // exclusive scan of the 64 tile counts: two per lane
        const uint32_t c0 = s_run[2 * lane], c1 = s_run[2 * lane + 1];
        uint32_t incl = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
          if (lane >= o) incl += t;
        }
        s_base[2 * lane] = incl - c0 - c1;
        s_base[2 * lane + 1] = incl - c1;
        if (lane == 31) {
          const uint32_t tot = incl;
          const uint32_t at = atomicAdd(&counters[kCntCursor], tot);
          atomicMax(&counters[kCntMaxTile], at + tot);  // per-frame maximum of the tile-list size a segment needs
          atomicAdd(&counters[kCntSumEntries], tot);    // tile entries of the whole frame (statistics)
          if (at + tot > tile_cap) { atomicOr(&counters[kCntOverflow], 2u); atomicOr(&counters[kCntStickyOverflow], 2u); s_alloc = 0xFFFFFFFFu; }
          else s_alloc = at;
        }
      }
      __syncthreads();
      const uint32_t alloc = s_alloc;
      if (threadIdx.x < 64) {
        const int t = threadIdx.x;
        const int tx = tile_x0 + (t & 7), ty = tile_y0 + (t >> 3);
        const bool in_band = tx < f.tiles_x && ty >= f.ty0 && ty < f.ty1 && (t < 32 ? do_lo : do_hi);
        // how many tiles need the shade kernel's full loop (its launch returns at once when there are none)
        const uint32_t fullm = __ballot_sync(0xFFFFFFFFu, in_band && alloc != 0xFFFFFFFFu && s_cls[t] != 0u && s_run[t] != 0u);
        if ((t & 31) == 0 && fullm) atomicAdd(&counters[kCntFullTiles], (uint32_t)__popc(fullm));
        if (row_cost) {  // tile entries per tile row of the frame: what a host balances the bands of a partition with
          uint32_t rsum = (in_band && alloc != 0xFFFFFFFFu) ? s_run[t] : 0u;
          rsum += __shfl_xor_sync(0xFFFFFFFFu, rsum, 1);
          rsum += __shfl_xor_sync(0xFFFFFFFFu, rsum, 2);
          rsum += __shfl_xor_sync(0xFFFFFFFFu, rsum, 4);
          if ((t & 7) == 0 && rsum) atomicAdd(&row_cost[ty], rsum);
        }
        if (in_band) {
          tile_start[ty * f.tiles_x + tx] = alloc == 0xFFFFFFFFu ? 0u : alloc + s_base[t];
#ifdef FDC_FINE_EXP  // timing experiments only (tools/variants_cfg.sh): the tiles stay empty, the lists are not (all) written
          tile_count[ty * f.tiles_x + tx] = 0u;
#else
          tile_count[ty * f.tiles_x + tx] = alloc == 0xFFFFFFFFu ? 0u : (s_run[t] | (s_cls[t] ? kTileNeedsFullPath : 0u));
#endif
        }
      }
      if (alloc == 0xFFFFFFFFu) return;
#if defined(FDC_FINE_EXP) && FDC_FINE_EXP == 2
      return;
#endif
      __syncthreads();
      if (threadIdx.x < 64) s_run[threadIdx.x] = alloc + s_base[threadIdx.x];
    }
    for (int cb = 0; cb < n_blocks; cb++)
    for (uint32_t s0 = 0; s0 == 0 || s0 < blk_total; s0 += kStage) {
      if (pass == 0 || !single) {
        __syncthreads();  // previous stage fully consumed
        if (use_seg && (pass == 1 || s0 == 0)) {
          // (pass 1 overwrote the table with write positions: every stage loads it again)
          const int c0 = cb * kSegBlock, n_seg = min(kSegBlock, n_chunks - c0);
#pragma unroll
          for (int j = 0; j < kSegBlock / 256; j++) {
            const int idx = (int)threadIdx.x + j * 256;
            const uint2 v = idx < n_seg ? __ldg(&seg_row[c0 + idx]) : make_uint2(0u, 0u);
            s_u.seg.start[idx] = v.x;
            s_u.seg.excl[idx] = v.y;
          }
          __syncthreads();
          uint32_t cnt[kSegBlock / 256], sum = 0;
#pragma unroll
          for (int j = 0; j < kSegBlock / 256; j++) { cnt[j] = s_u.seg.excl[threadIdx.x * (kSegBlock / 256) + j]; sum += cnt[j]; }
          uint32_t incl = sum;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
          }
          if (lane == 31) s_wsum[warp] = incl;
          __syncthreads();
          uint32_t run = incl - sum, total = 0;
#pragma unroll
          for (int w = 0; w < 8; w++) {
            const uint32_t t = s_wsum[w];
            if (w < warp) run += t;
            total += t;
          }
#pragma unroll
          for (int j = 0; j < kSegBlock / 256; j++) { s_u.seg.excl[threadIdx.x * (kSegBlock / 256) + j] = run; run += cnt[j]; }
          blk_total = total;
          if (pass == 0 && cb == 0 && s0 == 0) {
            single = n_blocks == 1 && total <= (uint32_t)kStage;
            if (split_big && single) {
              if (blockIdx.y == 1) return;
              do_lo = do_hi = true;
            }
          }
          __syncthreads();
        }
      }
      if (s0 >= blk_total) break;  // (an empty block)
      const uint32_t ns = min((uint32_t)kStage, blk_total - s0);
      const int n_groups = (int)((ns + 31u) >> 5);
      if (pass == 0 || !single) {
        // Four entries per thread with the three dependent gathers (list -> bbox -> inner rect) issued level by level,
        // so a stage costs three memory latencies instead of twelve.
        {
          uint32_t pid[4];
          int4 q6[4];
          int2 ir[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t k = threadIdx.x + j * 256u;
            if (k >= ns) pid[j] = 0xFFFFFFFFu;
            else if (n_direct) pid[j] = s0 + k;
            else {
              // entry q of the block -> its chunk: the last chunk whose scanned count is <= q (empty chunks tie with
              // their successor and lose)
              const uint32_t q = s0 + k;
              int o = 0;
#pragma unroll
              for (int step = kSegBlock / 2; step >= 1; step >>= 1)
                if (s_u.seg.excl[o + step] <= q) o += step;
              pid[j] = __ldg(&coarse_list[s_u.seg.start[o] + (q - s_u.seg.excl[o])]);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; j++)
            q6[j] = pid[j] != 0xFFFFFFFFu ? __ldg(reinterpret_cast<const int4*>(&prims[pid[j]])) : make_int4(0, 0, 0, 0);
#pragma unroll
          for (int j = 0; j < 4; j++)
            ir[j] = (pid[j] != 0xFFFFFFFFu && ((uint32_t)q6[j].z & PF_INNER)) ? __ldg(reinterpret_cast<const int2*>(&prims[pid[j]]) + 2)
                                                                              : make_int2(0, 0);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t k = threadIdx.x + j * 256u;
            if (k >= ns) break;
            stage_entry(k, pid[j], q6[j], ir[j], px0, py0, f, s_pid, s_rm0, s_rm1, s_cm0, s_cm1, s_info, s_lo, s_hi);
          }
        }
        __syncthreads();
        // entries per (group, tile)
        for (int g = warp; g < n_groups; g += 8) {
          const uint32_t k = (uint32_t)g * 32u + lane;
          const uint32_t lo = k < ns ? s_lo[k] : 0u, hi = k < ns ? s_hi[k] : 0u;
          const uint32_t t_lo = do_lo ? transpose32(lo, lane) : 0u, t_hi = do_hi ? transpose32(hi, lane) : 0u;
          s_gcnt[g][lane] = (uint8_t)__popc(t_lo);
          s_gcnt[g][32 + lane] = (uint8_t)__popc(t_hi);
          s_lo[k] = t_lo;  // (own slot: read above by this thread only)
          s_hi[k] = t_hi;
          // tile class: anything but unmasked fast content sends the whole tile to the shade kernel's full loop
          const uint32_t inf = k < ns ? s_info[k] : TE_FAST;
          const bool not_lean = !(inf & TE_FAST) || (inf & ((15u << TE_DEPTH_SHIFT) | TE_RECTMASK | TE_MASKW | TE_MASKB)) != 0u;
          const uint32_t nlm = __ballot_sync(0xFFFFFFFFu, not_lean);
          if (t_lo & nlm) atomicOr(&s_cls[lane], 1u);
          if (t_hi & nlm) atomicOr(&s_cls[32 + lane], 1u);
        }
        __syncthreads();
      }
      // per tile: running count (pass 0) or the groups' write positions (pass 1)
      if (threadIdx.x < 64) {
        const int t = threadIdx.x;
        uint32_t run = s_run[t];
        for (int g = 0; g < n_groups; g++) {
          if (pass == 1) s_u.gpos[g][t] = run;
          run += s_gcnt[g][t];
        }
        s_run[t] = run;
      }
      if (pass == 0) continue;
      __syncthreads();
      // The write loop.  Lane t owns tiles t and 32 + t; its share of an entry is one byte of a row word and one byte
      // of a column word: overlap and inner-rect bits travel together as two bytes of one register (TileEntry bits 0..15).
      const uint32_t sel_r = 0x4440u | (uint32_t)((lane >> 3) & 3), sel_c = 0x4440u | (uint32_t)(lane & 3);
      const uint32_t* s_cm = (lane & 4) ? s_cm1 : s_cm0;
#ifndef FDC_FINE_PAIR
#define FDC_FINE_PAIR 1
#endif
      for (int g = warp; g < n_groups; g += 8) {
        const uint32_t k = (uint32_t)g * 32u + lane;
#pragma unroll
        for (int half = 0; half < 2; half++) {
          if (half == 0 ? !do_lo : !do_hi) continue;
          uint32_t m = half ? s_hi[k] : s_lo[k];  // entries of this group that hit my tile (transposed while counting)
          const uint32_t* s_rm = half ? s_rm1 : s_rm0;
          uint32_t pos = s_u.gpos[g][half * 32 + lane];
#if defined(FDC_FINE_EXP) && FDC_FINE_EXP == 1
          m = 0;
#endif
#if FDC_FINE_PAIR
          uint2 held = make_uint2(0u, 0u);  // an entry at an even position waiting for its neighbour: the two leave as 16 bytes
          bool holding = false;
#endif
          while (m) {
            const int e = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t ke = (uint32_t)g * 32u + (uint32_t)e;
            const uint32_t info = s_info[ke];
            uint32_t rb = __byte_perm(s_rm[ke], 0u, sel_r);          // my tile row: overlap nibble | inner nibble << 4
            rb = (rb | (rb << 4)) & 0x0F0Fu;                          // -> one nibble per byte
            rb = (rb | (rb << 2)) & 0x3333u;                          // each row bit doubled (block = row * 2 + column),
            rb = ((rb | (rb << 1)) & 0x5555u) * 3u;                   //   both bytes at once
            const uint32_t cb = __byte_perm(0xFFAA5500u, 0u, __byte_perm(s_cm[ke], 0u, sel_c));  // column pairs -> 0x55 / 0xAA patterns
            uint32_t of = rb & cb;                                    // bits 0..7 overlap, 8..15 inside the inner rect
            if (info & kInfoEmptyInner) of = of & ~(of >> 8) & 0xFFu;  // inside an AnnularAA stroke: nothing to shade
            if (info & kInfoBegin) of |= 0xFFu;                        // PF_MASK_BEGIN: every block resets the level
            TileEntry te;
            te.pid = s_pid[ke];
            te.info = (info & 0x3FFFFFFFu) | of;
#if defined(FDC_FINE_EXP) && FDC_FINE_EXP == 3  // everything but the stores (the sink keeps the arithmetic alive)
            if (te.info == 0x12345678u && te.pid == 0x9ABCDEFu) tile_list[pos] = te;
            pos++;
#elif FDC_FINE_PAIR
            // half as many store transactions: 8-byte entries written one by one are one partial 32-byte sector each
            if (holding) {
              *reinterpret_cast<uint4*>(&tile_list[pos - 1]) = make_uint4(held.x, held.y, te.pid, te.info);
              holding = false;
            } else if (!(pos & 1u) && m) {
              held = make_uint2(te.pid, te.info);
              holding = true;
            } else {
              tile_list[pos] = te;
            }
            pos++;
#else
            tile_list[pos++] = te;
#endif
          }
        }
      }
    }
  }
}

__global__ void fill_u32_kernel(uint32_t* dst, uint32_t v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}
void launch_fill_u32(uint32_t* dst, uint32_t value, size_t n, cudaStream_t stream) {
  if (n == 0) return;
  fill_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dst, value, n);
}

void launch_binning(const PrimBin* prims, uint32_t n_prims, const FrameView& f, const BinBuffers& b, cudaStream_t stream,
                    int* n_launches) {
  const int n_bins = f.cbx * f.cby;
  const int n_chunks = (int)((n_prims + kChunk - 1) / kChunk);
  // tiles of the band start empty; per-segment counters: tile cursor, overflow flags, coarse total, scan ticket.
  // The per-FRAME words (kCntStickyOverflow ...) are zeroed once per frame by the caller, so an overflow in any segment
  // is still visible after later segments reset the per-segment words.
  if (n_prims == 0 || n_bins == 0) {
    cudaMemsetAsync(b.tile_count + (size_t)f.ty0 * f.tiles_x, 0, sizeof(uint32_t) * (size_t)(f.ty1 - f.ty0) * f.tiles_x, stream);
    cudaMemsetAsync(b.counters, 0, sizeof(uint32_t) * kCntStickyOverflow, stream);
    return;
  }
  // Otherwise nothing to clear: the setup kernel of this segment zeroed the per-segment counters, and the fine binner
  // writes start and count of EVERY tile of the band, empty ones included.  (After a list overflow it leaves them
  // stale -- and every later kernel of the frame returns at once on the sticky flag.)
  if ((size_t)n_prims * (size_t)n_bins <= (size_t)kDirectFineLimit) {
    // Small scene: the launch of coarse binning costs more than letting every bin look at every primitive.
    fine_bin_kernel<<<dim3(n_bins, n_bins <= kFineSplitBins ? 2 : 1), 256, 0, stream>>>(prims, f, nullptr, 0, b.coarse_list, b.coarse_cap,
                                                                                    b.tile_start, b.tile_count, b.tile_list, b.tile_cap,
                                                                                    b.counters, b.row_cost, n_prims, 0);
    if (n_launches) *n_launches += 1;
    return;
  }
  const int rows_per_cta = max(1, kSlots / max(f.cbx, 1));  // whole bin rows, at most kSlots bins per CTA
  dim3 grid(n_chunks, (f.cby + rows_per_cta - 1) / rows_per_cta);
  uint2* seg = reinterpret_cast<uint2*>(b.seg_table);  // [bin][chunk] (start, count)
  coarse_pairs_kernel<<<grid, kChunk, 0, stream>>>(prims, n_prims, f, rows_per_cta, seg, b.coarse_list, b.coarse_cap, b.counters);
  fine_bin_kernel<<<dim3(n_bins, FDC_FINE_SPLIT_BIG || n_bins <= kFineSplitBins ? 2 : 1), 256, 0, stream>>>(
      prims, f, seg, n_chunks, b.coarse_list, b.coarse_cap, b.tile_start, b.tile_count, b.tile_list, b.tile_cap, b.counters, b.row_cost,
      0u, FDC_FINE_SPLIT_BIG && n_bins > kFineSplitBins);
  if (n_launches) *n_launches += 2;
}

}  // namespace fdc
