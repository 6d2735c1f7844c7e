// Native scene flattening: the reference's front-end DFS restated as one host pass over POD `Fig`
// records that appends `fdc_call` records -- no per-call dynamic dispatch, no FFI per backend call.
//
// Follows src/figdraw/figrender.nim: renderFrame :1960-2002, renderRoot :1946-1958, render :1756-1839 (stage order:
// rotation, nkTransform, drop shadows, clip mask, rect mask, the node itself, inner shadows, children, cleanups in
// reverse), renderDropShadows :654-689, renderInnerShadows :716-744, renderRoundedShapeScaledCorners :806-873,
// renderText :417-497 (glyph loop), renderImage / renderMsdfImage / renderMtsdfImage / renderBackdropBlur :1673-1754,
// renderDrawable :1653-1667 with line :946-995, circle :1122-1136, rect :1138-1142, ellipse :1617-1635 and the
// 3-control quadratic Bezier :1330-1370; fills: figrender.nim:580-647 and toBackendFill figbackend.nim:109-127.
//
// All arithmetic is float32 in the reference's operation order (the file is compiled with -ffp-contract=off); the
// records are byte-identical to what the per-call recorder produces for the same scene (tests/test_flatten.py).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "fdc_flatten.h"

namespace fdc {

namespace {

struct BFill {  // BackendFill, figbackend.nim:96-107
  uint32_t kind = FDC_FILL_COLOR, axis = 0;
  uint32_t c[4] = {0, 0, 0, 0};
  float mid_pos = 0.5f;
};

struct Radii {
  float x[4], y[4];
};

inline float nim_round(float x) { return roundf(x); }  // half away from zero
inline float clampf(float v, float lo, float hi) { return std::min(std::max(v, lo), hi); }
inline uint32_t chan(uint32_t c, int k) { return (c >> (8 * k)) & 255u; }

uint32_t lerp_color(uint32_t a, uint32_t b, float t) {  // figrender.nim:580-590
  t = clampf(t, 0.0f, 1.0f);
  const float inv = 1.0f - t;
  uint32_t out = 0;
  for (int k = 0; k < 4; k++) {
    const float x = (float)chan(a, k) * inv;
    const float y = (float)chan(b, k) * t;
    const int v = (int)nim_round(x + y);
    out |= ((uint32_t)v & 255u) << (8 * k);
  }
  return out;
}

uint32_t fill_alpha_max(const fdc_node_fill& f) {
  if (f.kind == 0) return f.c[0] >> 24;
  if (f.kind == 1) return std::max(f.c[0] >> 24, f.c[2] >> 24);
  return std::max(std::max(f.c[0] >> 24, f.c[1] >> 24), f.c[2] >> 24);
}

float fill_mid(const fdc_node_fill& f) { return clampf((float)f.mid_pos / 255.0f, 0.01f, 0.99f); }

uint32_t sample_gradient(const fdc_node_fill& f, float t) {  // figrender.nim:600-615
  if (f.kind == 0) return f.c[0];
  if (f.kind == 1) return lerp_color(f.c[0], f.c[2], t);
  const float ct = clampf(t, 0.0f, 1.0f);
  const float mid = fill_mid(f);
  if (ct <= mid) return lerp_color(f.c[0], f.c[1], ct / mid);
  return lerp_color(f.c[1], f.c[2], (ct - mid) / (1.0f - mid));
}

uint32_t fill_center_color(const fdc_node_fill& f) { return sample_gradient(f, 0.5f); }

void gradient_colors(const fdc_node_fill& f, uint32_t out[4]) {  // figrender.nim:623-647, order BL, BR, TR, TL
  static const float ts[4][4] = {{0.0f, 1.0f, 1.0f, 0.0f}, {1.0f, 1.0f, 0.0f, 0.0f}, {0.5f, 1.0f, 0.5f, 0.0f}, {0.0f, 0.5f, 1.0f, 0.5f}};
  const int ax = f.kind != 0 ? (f.axis & 3) : 0;
  for (int k = 0; k < 4; k++) out[k] = sample_gradient(f, ts[ax][k]);
}

BFill to_backend_fill(const fdc_node_fill& f) {  // figbackend.nim:109-127
  BFill b;
  if (f.kind == 0) {
    b.kind = FDC_FILL_COLOR;
    b.c[0] = f.c[0];
  } else if (f.kind == 1) {
    b.kind = FDC_FILL_LINEAR2;
    b.axis = f.axis;
    b.c[0] = f.c[0];
    b.c[1] = f.c[2];
  } else {
    b.kind = FDC_FILL_LINEAR3;
    b.axis = f.axis;
    b.c[0] = f.c[0];
    b.c[1] = f.c[1];
    b.c[2] = f.c[2];
    b.mid_pos = fill_mid(f);
  }
  return b;
}

BFill solid_fill(uint32_t color) {
  BFill b;
  b.c[0] = color;
  return b;
}

fdc_node_stroke no_stroke() {
  fdc_node_stroke s;
  memset(&s, 0, sizeof(s));
  s.fill.mid_pos = 128;
  return s;
}

float radius_corner(float radius) {  // figrender.nim:560-571 (uint16 corner)
  if (radius <= 0.0f) return 0.0f;
  if (radius >= 65535.0f) return 65535.0f;
  return (float)(int)nim_round(radius);
}

struct V2 {
  float x, y;
};
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 operator*(V2 a, float k) { return {a.x * k, a.y * k}; }
inline float vlen(V2 v) { return sqrtf(v.x * v.x + v.y * v.y); }
inline V2 normalized_or(V2 v, V2 fallback) {  // figrender.nim:911-916
  const float len = vlen(v);
  if (len <= 0.000001f) return fallback;
  return {v.x / len, v.y / len};
}
inline V2 normal_left(V2 d) { return {-d.y, d.x}; }
inline float cross2(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
// Transcendentals in double, rounded to float32: what the Python mirror does (the reference calls cosf/sinf/acosf).
inline float cos32(float a) { return (float)cos((double)a); }
inline float sin32(float a) { return (float)sin((double)a); }

struct Span {  // DrawableQuadraticSpan, figrender.nim:1203-1214
  V2 p0, p1, p2;
  V2 start_tangent() const { return normalized_or(p1 - p0, normalized_or(p2 - p0, V2{1.0f, 0.0f})); }
  V2 end_tangent() const { return normalized_or(p2 - p1, normalized_or(p2 - p0, V2{1.0f, 0.0f})); }
};

constexpr float kAdaptiveTolerancePx = 0.5f;
constexpr int kMaxAdaptiveSteps = 192;  // max(DefaultDrawableBezierSteps * 4, 64), fignodes.nim:100
constexpr int kMaxAdaptiveDepth = 8;

struct Flattener {
  fdc_call* out;       // caller's buffer: records past `cap` are counted, not stored
  size_t cap, n = 0;
  fdc_call spill;      // where a record past the capacity is "written"
  const fdc_glyph* glyphs;
  const fdc_text_rect* text_rects;
  const fdc_draw_op* ops;
  const float* points;
  const fdc_flatten_env& env;
  float ui;
  float aa;
  bool subpixel;
  const char* error = nullptr;

  Flattener(fdc_call* o, size_t c, const fdc_scene& sc, const fdc_flatten_env& e)
      : out(o), cap(c), glyphs(sc.glyphs), text_rects(sc.text_rects), ops(sc.ops), points(sc.points), env(e), ui(e.ui_scale),
        aa(e.aa_factor), subpixel(e.subpixel_enabled != 0) {}

  // ---- the backend calls, recorded
  fdc_call& rec(uint32_t op) {
    fdc_call& c = n < cap ? out[n] : spill;
    n++;
    memset(&c, 0, sizeof(c));
    c.op = op;
    return c;
  }
  static void put_fill(fdc_call& c, const BFill& f) {
    c.u[1] = f.kind;
    c.u[2] = f.axis;
    for (int k = 0; k < 4; k++) c.u[3 + k] = f.c[k];
    c.f[16] = f.mid_pos;
  }
  static void put_rect_radii(fdc_call& c, const float r[4], const Radii& rad) {
    for (int k = 0; k < 4; k++) {
      c.f[k] = r[k];
      c.f[4 + k] = rad.x[k];
      c.f[8 + k] = rad.y[k];
    }
  }
  void save() { rec(FDC_OP_SAVE_TRANSFORM); }
  void restore() { rec(FDC_OP_RESTORE_TRANSFORM); }
  void translate(float x, float y) {
    fdc_call& c = rec(FDC_OP_TRANSLATE);
    c.f[0] = x;
    c.f[1] = y;
  }
  void rotate(float a) { rec(FDC_OP_ROTATE).f[0] = a; }
  void scale(float x, float y) {
    fdc_call& c = rec(FDC_OP_SCALE);
    c.f[0] = x;
    c.f[1] = y;
  }
  void set_aa(float v) {
    if (aa == v) return;
    aa = v;
    rec(FDC_OP_SET_AA).f[0] = v;
  }
  void set_subpixel_shift(float shift) {
    if (!subpixel) return;
    fdc_call& c = rec(FDC_OP_SET_SUBPIXEL);
    c.u[0] = 1;
    c.f[0] = shift;
  }
  bool has_image(uint64_t key) const {
    return env.image_keys && std::binary_search(env.image_keys, env.image_keys + env.n_image_keys, key);
  }
  void rounded_rect(const float r[4], const BFill& fill, const Radii& rad, int mode, float factor, float spread, float sx, float sy) {
    fdc_call& c = rec(FDC_OP_ROUNDED_RECT);
    put_rect_radii(c, r, rad);
    c.f[12] = factor;
    c.f[13] = spread;
    c.f[14] = sx;
    c.f[15] = sy;
    c.u[0] = (uint32_t)mode;
    put_fill(c, fill);
  }
  void draw_image(uint64_t key, float x, float y, const uint32_t colors[4], float w, float h, bool flip) {
    fdc_call& c = rec(FDC_OP_IMAGE);
    c.u[0] = (uint32_t)(key & 0xFFFFFFFFull);
    c.u[1] = (uint32_t)(key >> 32);
    for (int k = 0; k < 4; k++) c.u[3 + k] = colors[k];
    c.u[7] = flip ? 1u : 0u;
    c.f[0] = x; c.f[1] = y; c.f[2] = w; c.f[3] = h;
  }

  // ---- helpers
  void scaled_box(const float b[4], float o[4]) const {
    for (int k = 0; k < 4; k++) o[k] = b[k] * ui;
  }
  Radii scaled_corners(const float x[4], const float y[4]) const {
    Radii r;
    for (int k = 0; k < 4; k++) {
      r.x[k] = x[k] * ui;
      r.y[k] = y[k] * ui;
    }
    return r;
  }
  Radii node_corners(const fdc_fig& n) const {
    return scaled_corners(n.corners, (n.flags & FDC_NF_ELLIPTICAL_CORNERS) ? n.corner_radii_y : n.corners);
  }
  static bool shadow_active(const fdc_node_shadow& s, uint32_t style) {
    if (s.style != style) return false;
    if (s.blur <= 0.0f && s.spread <= 0.0f) return false;
    return fill_alpha_max(s.fill) != 0;
  }

  // ---- figrender.nim:654-689
  void drop_shadows(const fdc_fig& n) {
    for (int i = 0; i < 4; i++) {
      const fdc_node_shadow& s = n.u.rect.shadows[i];
      if (!shadow_active(s, 1)) continue;
      float box[4];
      scaled_box(n.screen_box, box);
      const float sx = s.x * ui, sy = s.y * ui, blur = s.blur * ui, spread = s.spread * ui;
      const float blur_pad = nim_round(1.5f * blur);
      const float pad = std::max(nim_round(spread) + blur_pad, 0.0f);
      const float srx = box[0] + sx, sry = box[1] + sy, srw = box[2] + 0.0f, srh = box[3] + 0.0f;
      const float quad[4] = {srx - pad, sry - pad, srw + 2.0f * pad, srh + 2.0f * pad};
      rounded_rect(quad, to_backend_fill(s.fill), node_corners(n), FDC_SDF_DROP_SHADOW, blur, spread, srw, srh);
    }
  }
  // ---- figrender.nim:716-744
  bool has_inner_shadow(const fdc_fig& n) const {
    for (int i = 0; i < 4; i++)
      if (shadow_active(n.u.rect.shadows[i], 2)) return true;
    return false;
  }
  void inner_shadows(const fdc_fig& n) {
    for (int i = 0; i < 4; i++) {
      const fdc_node_shadow& s = n.u.rect.shadows[i];
      if (!shadow_active(s, 2)) continue;
      float box[4];
      scaled_box(n.screen_box, box);
      rounded_rect(box, to_backend_fill(s.fill), node_corners(n), FDC_SDF_INSET_SHADOW, s.blur * ui, s.spread * ui, s.x * ui, s.y * ui);
    }
  }
  // ---- figrender.nim:806-873 (SDF branch)
  void rounded_shape_scaled(const float shape_box[4], const fdc_node_fill& fill, const fdc_node_stroke& stroke, const Radii& corners) {
    float box[4];
    scaled_box(shape_box, box);
    const bool gradient = (fill.kind == 1 || fill.kind == 2) && fill_alpha_max(fill) > 0;
    if (gradient) rounded_rect(box, to_backend_fill(fill), corners, FDC_SDF_CLIP_AA, 4.0f, 0.0f, 0.0f, 0.0f);
    else if (fill_alpha_max(fill) > 0) rounded_rect(box, solid_fill(fill_center_color(fill)), corners, FDC_SDF_CLIP_AA, 4.0f, 0.0f, 0.0f, 0.0f);
    if (fill_alpha_max(stroke.fill) > 0 && stroke.weight > 0.0f)
      rounded_rect(box, to_backend_fill(stroke.fill), corners, FDC_SDF_ANNULAR_AA, stroke.weight * ui, 0.0f, 0.0f, 0.0f);
  }
  void rounded_shape(const float box[4], const fdc_node_fill& fill, const fdc_node_stroke& stroke, const float cx[4], const float cy[4]) {
    rounded_shape_scaled(box, fill, stroke, scaled_corners(cx, cy));
  }
  void boxes(const fdc_fig& n, const fdc_node_stroke& stroke) {
    rounded_shape(n.screen_box, n.fill, stroke, n.corners, (n.flags & FDC_NF_ELLIPTICAL_CORNERS) ? n.corner_radii_y : n.corners);
  }

  // ---- drawables
  void stroke_cap(float cx, float cy, float radius, const fdc_node_fill& fill) {
    if (radius <= 0.0f || fill_alpha_max(fill) == 0) return;
    const float d = radius * 2.0f;
    const float box[4] = {cx - radius, cy - radius, d, d};
    const float rc = radius_corner(radius);
    const float c4[4] = {rc, rc, rc, rc};
    rounded_shape(box, fill, no_stroke(), c4, c4);
  }
  // figrender.nim:946-995: a rotated zero-radius box (+ round caps)
  void drawable_line(float ox, float oy, const float a[2], const float b[2], const fdc_node_stroke& stroke) {
    const float weight = std::max(0.0f, stroke.weight);
    if (weight <= 0.0f || fill_alpha_max(stroke.fill) == 0) return;
    const float ax = ox + a[0], ay = oy + a[1], bx = ox + b[0], by = oy + b[1];
    const float dx = bx - ax, dy = by - ay;
    const float length = sqrtf(dx * dx + dy * dy);
    if (length <= 0.0f) return;
    const int cap = stroke.cap == 0 ? 2 : stroke.cap;  // scAuto -> scButt
    const float cap_radius = weight * 0.5f;
    const float dirx = dx / length, diry = dy / length;
    float dax = ax, day = ay, dbx = bx, dby = by, draw_length = length;
    if (cap == 3) {  // scSquare
      dax = ax - dirx * cap_radius; day = ay - diry * cap_radius;
      dbx = bx + dirx * cap_radius; dby = by + diry * cap_radius;
      draw_length = length + weight;
    }
    const float cx = (dax + dbx) / 2.0f, cy = (day + dby) / 2.0f;
    const float box[4] = {cx - draw_length / 2.0f, cy - weight / 2.0f, draw_length, weight};
    float sb[4];
    scaled_box(box, sb);
    const float px = sb[0] + sb[2] / 2.0f, py = sb[1] + sb[3] / 2.0f;
    const float angle = (float)atan2((double)dy, (double)dx);
    save();
    translate(px, py);
    rotate(angle);
    translate(-px, -py);
    const float zero[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    rounded_shape(box, stroke.fill, no_stroke(), zero, zero);
    restore();
    if (cap == 1) {  // scRound
      stroke_cap(ax, ay, cap_radius, stroke.fill);
      stroke_cap(bx, by, cap_radius, stroke.fill);
    }
  }
  static void quadratic_point(const float p0[2], const float p1[2], const float p2[2], float t, float q[2]) {
    const float inv = 1.0f - t;
    for (int k = 0; k < 2; k++) q[k] = p0[k] * (inv * inv) + p1[k] * (2.0f * inv * t) + p2[k] * (t * t);
  }
  static void quadratic_bounds(const float p0[2], const float p1[2], const float p2[2], float padding, float box[4]) {
    float mn[2] = {std::min(p0[0], p2[0]), std::min(p0[1], p2[1])};
    float mx[2] = {std::max(p0[0], p2[0]), std::max(p0[1], p2[1])};
    for (int k = 0; k < 2; k++) {
      const float denom = p0[k] - 2.0f * p1[k] + p2[k];
      if (fabs((double)denom) > 0.000001) {
        const float t = (p0[k] - p1[k]) / denom;
        if ((double)t > 0.0 && (double)t < 1.0) {
          float q[2];
          quadratic_point(p0, p1, p2, t, q);
          for (int j = 0; j < 2; j++) {
            mn[j] = std::min(mn[j], q[j]);
            mx[j] = std::max(mx[j], q[j]);
          }
        }
      }
    }
    box[0] = mn[0] - padding; box[1] = mn[1] - padding;
    box[2] = mx[0] - mn[0] + padding * 2.0f; box[3] = mx[1] - mn[1] + padding * 2.0f;
  }
  // figrender.nim:1327-1366.  cap 0 (scAuto) resolves through the stroke.
  void quadratic_bezier(float ox, float oy, V2 q0, V2 q1, V2 q2, const fdc_node_stroke& stroke, int cap_in) {
    const int cap = cap_in != 0 ? cap_in : (stroke.cap == 0 ? 1 : stroke.cap);  // scAuto -> scRound
    const float p0[2] = {q0.x, q0.y}, p1[2] = {q1.x, q1.y}, p2[2] = {q2.x, q2.y};
    const float cr = (p1[0] - p0[0]) * (p2[1] - p1[1]) - (p1[1] - p0[1]) * (p2[0] - p1[0]);
    if (fabs((double)cr) <= 0.0001) {
      fdc_node_stroke s2 = stroke;
      s2.cap = (uint8_t)cap;
      drawable_line(ox, oy, p0, p2, s2);
      return;
    }
    const float weight = std::max(0.0f, stroke.weight);
    const float padding = weight * 0.5f + 2.0f / ui;
    const float a[2] = {ox + p0[0], oy + p0[1]}, b[2] = {ox + p1[0], oy + p1[1]}, c[2] = {ox + p2[0], oy + p2[1]};
    float box[4];
    quadratic_bounds(a, b, c, padding, box);
    if (box[2] <= 0.0f || box[3] <= 0.0f) return;
    const float cx = box[0] + box[2] * 0.5f, cy = box[1] + box[3] * 0.5f;
    fdc_call& r = rec(FDC_OP_BEZIER);
    float sb[4];
    scaled_box(box, sb);
    for (int k = 0; k < 4; k++) r.f[k] = sb[k];
    r.f[4] = (a[0] - cx) * ui; r.f[5] = (a[1] - cy) * ui;
    r.f[6] = (b[0] - cx) * ui; r.f[7] = (b[1] - cy) * ui;
    r.f[8] = (c[0] - cx) * ui; r.f[9] = (c[1] - cy) * ui;
    r.f[10] = weight * ui;
    r.u[0] = (uint32_t)cap;
    put_fill(r, to_backend_fill(stroke.fill));
  }
  void line_v(float ox, float oy, V2 a, V2 b, const fdc_node_stroke& stroke) {
    const float pa[2] = {a.x, a.y}, pb[2] = {b.x, b.y};
    drawable_line(ox, oy, pa, pb, stroke);
  }
  // figrender.nim:1010-1039
  void endpoint_cap(float ox, float oy, V2 point, V2 tangent, float radius, const fdc_node_stroke& stroke, int cap, bool is_start) {
    if (radius <= 0.0f || fill_alpha_max(stroke.fill) == 0) return;
    if (cap == 1) {
      stroke_cap(ox + point.x, oy + point.y, radius, stroke.fill);
    } else if (cap == 3) {
      const V2 dir = normalized_or(tangent, V2{1.0f, 0.0f});
      const V2 a = is_start ? point - dir * radius : point;
      const V2 b = is_start ? point : point + dir * radius;
      fdc_node_stroke s2 = stroke;
      s2.cap = 2;  // scButt
      line_v(ox, oy, a, b, s2);
    }
  }
  // figrender.nim:1049-1057
  void filled_quad(const V2 v[4], const fdc_node_fill& fill) {
    if (fill_alpha_max(fill) == 0) return;
    const uint32_t color = fill_center_color(fill);
    fdc_call& c = rec(FDC_OP_FILLED_QUAD);
    for (int k = 0; k < 4; k++) {
      c.f[2 * k] = v[k].x * ui;
      c.f[2 * k + 1] = v[k].y * ui;
      c.u[3 + k] = color;
    }
  }
  // figrender.nim:1059-1109
  void stroke_join(float ox, float oy, V2 point, V2 incoming_tangent, V2 outgoing_tangent, float radius, const fdc_node_fill& fill, int join) {
    if (radius <= 0.0f || fill_alpha_max(fill) == 0) return;
    if (join == 1) {  // sjRound
      stroke_cap(ox + point.x, oy + point.y, radius, fill);
      return;
    }
    if (join != 2 && join != 3) return;  // sjBevel, sjMiter
    const V2 incoming = normalized_or(incoming_tangent, V2{1.0f, 0.0f});
    const V2 outgoing = normalized_or(outgoing_tangent, incoming);
    const float turn = cross2(incoming, outgoing);
    if (fabs((double)turn) <= 0.0001) return;
    const float side = turn > 0.0f ? -1.0f : 1.0f;
    const V2 in_outer = point + normal_left(incoming) * (radius * side);
    const V2 out_outer = point + normal_left(outgoing) * (radius * side);
    const V2 origin{ox, oy};
    if (join == 3) {
      const float denom = cross2(incoming, outgoing);  // lineIntersection(p = in_outer, r = incoming, q = out_outer, s = outgoing)
      if (!(fabs((double)denom) <= 0.000001)) {
        const float t = cross2(out_outer - in_outer, outgoing) / denom;
        const V2 miter = in_outer + incoming * t;
        if (vlen(miter - point) <= radius * 4.0f) {
          const V2 q[4] = {origin + point, origin + in_outer, origin + miter, origin + out_outer};
          filled_quad(q, fill);
          return;
        }
      }
    }
    const V2 q[4] = {origin + point, origin + in_outer, origin + out_outer, origin + out_outer};
    filled_quad(q, fill);
  }
  // De Casteljau, figrender.nim:1134-1147
  V2 bezier_point(const float* ctrl, uint32_t n, float t) const {
    if (n == 0) return V2{0.0f, 0.0f};
    V2 work[64];
    std::vector<V2> big;
    V2* w = work;
    if (n > 64) { big.resize(n); w = big.data(); }
    for (uint32_t i = 0; i < n; i++) w[i] = V2{ctrl[2 * i], ctrl[2 * i + 1]};
    for (uint32_t count = n; count > 1; count--)
      for (uint32_t i = 0; i + 1 < count; i++) w[i] = w[i] * (1.0f - t) + w[i + 1] * t;
    return w[0];
  }
  static V2 quadratic_point_v(V2 p0, V2 p1, V2 p2, float t) {
    const float inv = 1.0f - t;
    return p0 * (inv * inv) + p1 * (2.0f * inv * t) + p2 * (t * t);
  }
  static int explicit_steps(uint32_t steps, int32_t node_steps) {  // figrender.nim:1195-1201
    if (steps != 0) return std::max(1, (int)steps);
    if (node_steps != 0) return std::max(1, (int)node_steps);
    return 0;
  }
  Span bezier_span(const float* ctrl, uint32_t n, float t0, float t2) const {  // figrender.nim:1230-1239
    const float tm = (t0 + t2) * 0.5f;
    const V2 p0 = bezier_point(ctrl, n, t0), pm = bezier_point(ctrl, n, tm), p2 = bezier_point(ctrl, n, t2);
    return Span{p0, pm * 2.0f - (p0 + p2) * 0.5f, p2};
  }
  float approx_error_px(const float* ctrl, uint32_t n, const Span& s, float t0, float t2) const {  // :1248-1256
    float result = 0.0f;
    const float locals[2] = {0.25f, 0.75f};
    for (float local_t : locals) {
      const float t = t0 + (t2 - t0) * local_t;
      const V2 actual = bezier_point(ctrl, n, t);
      const V2 approx = quadratic_point_v(s.p0, s.p1, s.p2, local_t);
      result = std::max(result, vlen((actual - approx) * ui));
    }
    return result;
  }
  void adaptive_spans(const float* ctrl, uint32_t n, float t0, float t2, int depth, std::vector<Span>& spans) const {  // :1258-1272
    const Span s = bezier_span(ctrl, n, t0, t2);
    const float error = approx_error_px(ctrl, n, s, t0, t2);
    if (error <= kAdaptiveTolerancePx || depth >= kMaxAdaptiveDepth || (int)spans.size() >= kMaxAdaptiveSteps - 1) {
      spans.push_back(s);
    } else {
      const float tm = (t0 + t2) * 0.5f;
      adaptive_spans(ctrl, n, t0, tm, depth + 1, spans);
      adaptive_spans(ctrl, n, tm, t2, depth + 1, spans);
    }
  }
  // shared body of renderDrawableBezierQuadratics (:1414-1457) and renderDrawableArcQuadratics (:1551-1593)
  void quadratic_spans(float ox, float oy, const std::vector<Span>& spans, const fdc_node_stroke& stroke) {
    const int cap = stroke.cap == 0 ? 1 : stroke.cap, join = stroke.join == 0 ? 1 : stroke.join;
    const bool simple_round = cap == 1 && join == 1;
    const int span_cap = simple_round ? 1 : 2;
    const float cap_radius = std::max(0.0f, stroke.weight) / 2.0f;
    for (size_t step = 0; step < spans.size(); step++) {
      const Span& s = spans[step];
      quadratic_bezier(ox, oy, s.p0, s.p1, s.p2, stroke, span_cap);
      if (!simple_round) {
        if (step == 0) endpoint_cap(ox, oy, s.p0, s.start_tangent(), cap_radius, stroke, cap, true);
        else stroke_join(ox, oy, s.p0, spans[step - 1].end_tangent(), s.start_tangent(), cap_radius, stroke.fill, join);
        if (step + 1 == spans.size()) endpoint_cap(ox, oy, s.p2, s.end_tangent(), cap_radius, stroke, cap, false);
      }
    }
  }
  static float distance_to_line(V2 p, V2 a, V2 b) {  // figrender.nim:1219-1225
    const V2 ab = b - a;
    const float denom = ab.x * ab.x + ab.y * ab.y;
    if (denom <= 0.000001f) return vlen(p - a);
    const V2 pa = p - a;
    const float h = clampf((pa.x * ab.x + pa.y * ab.y) / denom, 0.0f, 1.0f);
    return vlen(p - (a + ab * h));
  }
  void adaptive_segment_points(const float* ctrl, uint32_t n, float t0, float t2, int depth, std::vector<V2>& pts) const {  // :1283-1297
    const V2 p0 = bezier_point(ctrl, n, t0), p2 = bezier_point(ctrl, n, t2);
    const float tm = (t0 + t2) * 0.5f;
    const V2 pm = bezier_point(ctrl, n, tm);
    const float error = distance_to_line(pm * ui, p0 * ui, p2 * ui);
    if (error <= kAdaptiveTolerancePx || depth >= kMaxAdaptiveDepth || (int)pts.size() >= kMaxAdaptiveSteps) {
      pts.push_back(p2);
    } else {
      adaptive_segment_points(ctrl, n, t0, tm, depth + 1, pts);
      adaptive_segment_points(ctrl, n, tm, t2, depth + 1, pts);
    }
  }
  // figrender.nim:1368-1412: polyline with endpoint caps and joins (what a 2-control Bezier takes)
  void bezier_segments(float ox, float oy, const float* ctrl, uint32_t n, const fdc_node_stroke& stroke, int fixed_steps) {
    std::vector<V2> pts;
    pts.push_back(bezier_point(ctrl, n, 0.0f));
    if (fixed_steps > 0) {
      for (int step = 1; step <= fixed_steps; step++) pts.push_back(bezier_point(ctrl, n, (float)step / (float)fixed_steps));
    } else {
      adaptive_segment_points(ctrl, n, 0.0f, 1.0f, 0, pts);
    }
    if (pts.size() < 2) return;
    const int cap = stroke.cap == 0 ? 1 : stroke.cap, join = stroke.join == 0 ? 1 : stroke.join;
    const float cap_radius = std::max(0.0f, stroke.weight) / 2.0f;
    fdc_node_stroke seg = stroke;
    seg.cap = 2;  // scButt
    V2 previous = pts[0], previous_tangent{1.0f, 0.0f};
    for (size_t step = 1; step < pts.size(); step++) {
      const V2 current = pts[step], tangent = current - previous;
      line_v(ox, oy, previous, current, seg);
      if (step == 1) endpoint_cap(ox, oy, previous, tangent, cap_radius, stroke, cap, true);
      else stroke_join(ox, oy, previous, previous_tangent, tangent, cap_radius, stroke.fill, join);
      if (step + 1 == pts.size()) endpoint_cap(ox, oy, current, tangent, cap_radius, stroke, cap, false);
      previous = current;
      previous_tangent = tangent;
    }
  }
  // figrender.nim:1459-1486 (SDF build)
  void bezier(float ox, float oy, const fdc_draw_op& op, const fdc_node_stroke& stroke, int32_t node_steps) {
    const uint32_t n = op.n_points;
    if (n < 2) return;
    if (stroke.weight <= 0.0f || fill_alpha_max(stroke.fill) == 0) return;
    const float* ctrl = points + 2 * (size_t)op.first_point;
    if (n == 3) {
      quadratic_bezier(ox, oy, V2{ctrl[0], ctrl[1]}, V2{ctrl[2], ctrl[3]}, V2{ctrl[4], ctrl[5]}, stroke, stroke.cap == 0 ? 1 : stroke.cap);
    } else if (n > 3) {
      const int fixed = explicit_steps(op.steps, node_steps);
      std::vector<Span> spans;
      if (fixed > 0) {
        for (int step = 0; step < fixed; step++) spans.push_back(bezier_span(ctrl, n, (float)step / (float)fixed, (float)(step + 1) / (float)fixed));
      } else {
        adaptive_spans(ctrl, n, 0.0f, 1.0f, 0, spans);
      }
      quadratic_spans(ox, oy, spans, stroke);
    } else {
      bezier_segments(ox, oy, ctrl, n, stroke, explicit_steps(op.steps, node_steps));
    }
  }
  int adaptive_arc_steps(float radius, float sweep) const {  // figrender.nim:1307-1318
    const float radius_px = std::max(0.0f, radius * ui);
    const float abs_sweep = fabsf(sweep);
    if (radius_px <= 0.0f || abs_sweep <= 0.0f) return 1;
    const float cos_limit = clampf(1.0f - kAdaptiveTolerancePx / radius_px, -1.0f, 1.0f);
    const float max_angle = std::max(0.01f, 2.0f * (float)acos((double)cos_limit));
    const int n = (int)ceil((double)(abs_sweep / max_angle));
    return std::min(std::max(n, 1), kMaxAdaptiveSteps);
  }
  // figrender.nim:1595-1611 -> :1551-1593, arcQuadraticSpan :1535-1549
  void arc(float ox, float oy, const fdc_draw_op& op, const fdc_node_stroke& stroke, int32_t node_steps) {
    const float radius = std::max(0.0f, op.radius);
    if (radius <= 0.0f || op.sweep_angle == 0.0f) return;
    if (stroke.weight <= 0.0f || fill_alpha_max(stroke.fill) == 0) return;
    int steps = explicit_steps(op.steps, node_steps);
    if (steps == 0) steps = adaptive_arc_steps(op.radius, op.sweep_angle);
    const V2 center{op.center[0], op.center[1]};
    auto arc_point = [&](float angle) { return V2{center.x + cos32(angle) * radius, center.y + sin32(angle) * radius}; };
    std::vector<Span> spans;
    for (int step = 0; step < steps; step++) {
      const float t0 = (float)step / (float)steps, t2 = (float)(step + 1) / (float)steps;
      const float tm = (t0 + t2) * 0.5f;
      const V2 p0 = arc_point(op.start_angle + op.sweep_angle * t0);
      const V2 pm = arc_point(op.start_angle + op.sweep_angle * tm);
      const V2 p2 = arc_point(op.start_angle + op.sweep_angle * t2);
      spans.push_back(Span{p0, pm * 2.0f - (p0 + p2) * 0.5f, p2});
    }
    quadratic_spans(ox, oy, spans, stroke);
  }
  void drawable_ops(const fdc_fig& n) {
    const float ox = n.screen_box[0], oy = n.screen_box[1];
    const fdc_node_stroke& stroke = n.u.drawable.stroke;
    for (uint32_t i = 0; i < n.u.drawable.n_ops && !error; i++) {
      const fdc_draw_op& op = ops[n.u.drawable.first_op + i];
      switch (op.kind) {
        case 0: drawable_line(ox, oy, op.a, op.b, stroke); break;
        case 1: {
          const float radius = std::max(0.0f, op.radius);
          if (radius <= 0.0f) break;
          const float d = radius * 2.0f;
          const float box[4] = {ox + op.center[0] - radius, oy + op.center[1] - radius, d, d};
          const float rc = radius_corner(radius);
          const float c4[4] = {rc, rc, rc, rc};
          rounded_shape(box, n.fill, stroke, c4, c4);
          break;
        }
        case 2: {
          const float box[4] = {ox + op.box[0], oy + op.box[1], op.box[2], op.box[3]};
          rounded_shape(box, n.fill, stroke, op.corners, op.corners);
          break;
        }
        case 5: {
          const float rx = std::max(0.0f, op.ellipse_radii[0]), ry = std::max(0.0f, op.ellipse_radii[1]);
          if (rx <= 0.0f || ry <= 0.0f) break;
          const float box[4] = {ox + op.center[0] - rx, oy + op.center[1] - ry, rx * 2.0f, ry * 2.0f};
          const float cx4[4] = {rx, rx, rx, rx}, cy4[4] = {ry, ry, ry, ry};
          rounded_shape(box, n.fill, stroke, cx4, cy4);
          break;
        }
        case 3: bezier(ox, oy, op, stroke, n.u.drawable.steps); break;
        case 4: arc(ox, oy, op, stroke, n.u.drawable.steps); break;
        default: error = "unknown drawable op kind"; break;
      }
    }
  }
  // figrender.nim:1653-1667
  void drawable(const fdc_fig& n) {
    const float want = n.u.drawable.aa;
    if (want <= 0.0f || aa == want) {
      drawable_ops(n);
      return;
    }
    const float old = aa;
    set_aa(want);
    drawable_ops(n);
    set_aa(old);
  }
  // figrender.nim:417-497: selection rects, underline/strikethrough, then one atlas quad per glyph.  The text layout
  // (pixie arrangement) is upstream: its results arrive as fdc_text_rect / fdc_glyph records.
  void text(const fdc_fig& n) {
    save();
    translate(n.screen_box[0] * ui, n.screen_box[1] * ui);
    if (n.flags & FDC_NF_INVERT_Y) {
      translate(0.0f, n.screen_box[3] * ui);
      scale(1.0f, -1.0f);
    }
    const Radii zero = {{0.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f, 0.0f}};
    const fdc_text_rect* tr = text_rects ? text_rects + n.u.text.first_rect : nullptr;
    if ((n.flags & FDC_NF_SELECT_TEXT) && fill_alpha_max(n.fill) > 0) {
      for (uint32_t i = 0; i < n.u.text.n_selection && tr; i++) {
        const float* r = tr[i].rect;
        if (!(r[3] > 0.0f)) continue;
        const float sel[4] = {r[0], r[1], std::max(r[2], 1.0f), r[3]};
        float sb[4];
        scaled_box(sel, sb);
        rounded_rect(sb, to_backend_fill(n.fill), zero, FDC_SDF_CLIP_AA, 4.0f, 0.0f, 0.0f, 0.0f);
      }
    }
    for (uint32_t i = 0; i < n.u.text.n_decoration && tr; i++) {  // drawTextDecoration :355-368
      const fdc_text_rect& d = tr[n.u.text.n_selection + i];
      if (d.rect[2] <= 0.0f || d.rect[3] <= 0.0f) continue;
      float sb[4];
      scaled_box(d.rect, sb);
      rounded_rect(sb, to_backend_fill(d.fill), zero, FDC_SDF_CLIP_AA, 4.0f, 0.0f, 0.0f, 0.0f);
    }
    for (uint32_t i = 0; i < n.u.text.n_glyphs; i++) {
      const fdc_glyph& g = glyphs[n.u.text.first_glyph + i];
      float gx = g.pos[0], shift = 0.0f;
      if (subpixel) {  // :462-471 (per-glyph variants are a different atlas key: the host's choice)
        const float snapped = floorf(gx);
        shift = std::max(0.0f, std::min(gx - snapped, 0.999f));
        gx = snapped;
      }
      set_subpixel_shift(shift);
      if (!has_image(g.key)) {
        set_subpixel_shift(0.0f);
        continue;
      }
      uint32_t cols[4];
      gradient_colors(g.fill, cols);
      draw_image(g.key, gx, g.pos[1], cols, 0.0f, 0.0f, false);
      if (subpixel) set_subpixel_shift(0.0f);
    }
    set_subpixel_shift(0.0f);
    restore();
  }
  void image(const fdc_fig& n) {
    if (n.u.image.id == 0) return;
    float box[4];
    scaled_box(n.screen_box, box);
    const uint32_t c = fill_center_color(n.u.image.fill);
    const uint32_t cols[4] = {c, c, c, c};
    draw_image(n.u.image.id, box[0], box[1], cols, box[2], box[3], (n.flags & FDC_NF_INVERT_Y) != 0);
  }
  void sdf_image(const fdc_fig& n, bool mtsdf) {
    if (n.u.msdf.id == 0) return;
    float box[4];
    scaled_box(n.screen_box, box);
    fdc_call& c = rec(FDC_OP_MSDF);
    c.u[0] = (uint32_t)(n.u.msdf.id & 0xFFFFFFFFull);
    c.u[1] = (uint32_t)(n.u.msdf.id >> 32);
    c.u[2] = mtsdf ? 1u : 0u;
    c.u[3] = fill_center_color(n.u.msdf.fill);
    c.u[7] = (n.flags & FDC_NF_INVERT_Y) ? 1u : 0u;
    for (int k = 0; k < 4; k++) c.f[k] = box[k];
    c.f[4] = n.u.msdf.px_range > 0.0f ? n.u.msdf.px_range : 4.0f;
    c.f[5] = (n.u.msdf.sd_threshold > 0.0f && n.u.msdf.sd_threshold < 1.0f) ? n.u.msdf.sd_threshold : 0.5f;
    c.f[6] = std::max(0.0f, n.u.msdf.stroke_weight) * ui;
  }
  // figrender.nim:1734-1754
  void backdrop_blur(const fdc_fig& n) {
    if (n.u.backdrop.blur > 0.0f) {
      float box[4];
      scaled_box(n.screen_box, box);
      fdc_call& c = rec(FDC_OP_BACKDROP_BLUR);
      put_rect_radii(c, box, node_corners(n));
      c.f[12] = n.u.backdrop.blur * ui;
    }
    if (fill_alpha_max(n.fill) == 0) return;
    boxes(n, no_stroke());
  }

  // ---- figrender.nim:1756-1839
  void render(const fdc_render_list& L, uint32_t idx, int depth) {
    if (error) return;
    if (depth > 4096) { error = "node tree deeper than 4096"; return; }
    const fdc_fig& n = L.nodes[idx];
    if (n.flags & FDC_NF_DISABLE_RENDER) return;
    float box[4];
    scaled_box(n.screen_box, box);
    enum { kRestore = 1, kPopMask = 2, kPopRectMask = 3 };
    int cleanups[5];
    int n_clean = 0;

    if (n.rotation != 0.0f) {
      save();
      const float cx = box[0] + box[2] / 2.0f, cy = box[1] + box[3] / 2.0f;
      translate(cx, cy);
      rotate(n.rotation / 180.0f * (float)M_PI);
      translate(-cx, -cy);
      cleanups[n_clean++] = kRestore;
    }
    if (n.kind == FDC_NK_TRANSFORM) {
      save();
      const float* tr = n.u.transform.translation;
      if (tr[0] != 0.0f || tr[1] != 0.0f) translate(tr[0] * ui, tr[1] * ui);
      if (n.u.transform.use_matrix) {
        fdc_call& c = rec(FDC_OP_APPLY_TRANSFORM);
        for (int k = 0; k < 16; k++) c.f[k] = n.u.transform.matrix[k];
      }
      cleanups[n_clean++] = kRestore;
    }
    if (n.kind == FDC_NK_RECTANGLE) drop_shadows(n);
    if (n.flags & FDC_NF_CLIP_CONTENT) {
      put_rect_radii(rec(FDC_OP_BEGIN_MASK), box, node_corners(n));
      rec(FDC_OP_END_MASK);
      cleanups[n_clean++] = kPopMask;
    }
    if (n.flags & FDC_NF_RECT_MASK_CONTENT) {
      put_rect_radii(rec(FDC_OP_BEGIN_RECT_MASK), box, node_corners(n));
      cleanups[n_clean++] = kPopRectMask;
    }
    switch (n.kind) {
      case FDC_NK_TEXT: text(n); break;
      case FDC_NK_DRAWABLE: drawable(n); break;
      case FDC_NK_RECTANGLE: boxes(n, n.u.rect.stroke); break;
      case FDC_NK_IMAGE: image(n); break;
      case FDC_NK_MSDF_IMAGE: sdf_image(n, false); break;
      case FDC_NK_MTSDF_IMAGE: sdf_image(n, true); break;
      case FDC_NK_BACKDROP_BLUR: backdrop_blur(n); break;
      default: break;
    }
    if (n.kind == FDC_NK_RECTANGLE && has_inner_shadow(n)) inner_shadows(n);

    // childIndex, fignodes.nim:165-177: the next `child_count` nodes after idx whose parent is idx
    int32_t found = 0;
    for (uint32_t j = idx + 1; found < n.child_count && j < L.n_nodes && !error; j++) {
      if (L.nodes[j].parent == (int32_t)idx) {
        found++;
        render(L, j, depth + 1);
      }
    }
    for (int k = n_clean - 1; k >= 0; k--) {
      if (cleanups[k] == kRestore) restore();
      else if (cleanups[k] == kPopMask) rec(FDC_OP_POP_MASK);
      else rec(FDC_OP_POP_RECT_MASK);
    }
  }
};

}  // namespace

namespace {

struct RootRef {
  uint32_t list, root;
};

// Flattens roots [r0, r1) of `roots` into out[0..cap); returns the record count needed.
size_t flatten_roots(const fdc_scene& scene, const std::vector<RootRef>& roots, size_t r0, size_t r1, const fdc_flatten_env& env,
                     fdc_call* out, size_t cap, const char** error) {
  Flattener F(out, cap, scene, env);
  for (size_t r = r0; r < r1 && !F.error; r++) F.render(scene.lists[roots[r].list], roots[r].root, 0);
  if (F.error) *error = F.error;
  return F.n;
}

}  // namespace

// Roots are independent of each other (a drawable restores the AA factor it changes; nothing else carries state from
// one root to the next), so a large scene is flattened by several threads: one counting pass per chunk of roots gives
// every chunk its output offset, the second pass writes the records in place.
const char* flatten_renders(const fdc_scene& scene, const fdc_flatten_env& env, fdc_call* out, size_t cap, size_t* n_out) {
  const fdc_render_list* lists = scene.lists;
  const uint32_t n_lists = scene.n_lists;
  *n_out = 0;
  if (!(env.ui_scale > 0.0f)) return "ui_scale must be positive";
  std::vector<RootRef> roots;
  for (uint32_t l = 0; l < n_lists; l++)
    for (uint32_t r = 0; r < lists[l].n_roots; r++) {
      const int32_t root = lists[l].root_ids[r];
      if (root < 0 || (uint32_t)root >= lists[l].n_nodes) return "root id out of range";
      roots.push_back({l, (uint32_t)root});
    }
  const char* error = nullptr;
  size_t n = 0;
  {  // renderFrame prologue: saveTransform, scale(pixelScale)
    Flattener F(out, cap, scene, env);
    F.save();
    F.scale(env.pixel_scale, env.pixel_scale);
    n = F.n;
  }
  size_t total_nodes = 0;
  for (uint32_t l = 0; l < n_lists; l++) total_nodes += lists[l].n_nodes;
  unsigned n_threads = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  if (const char* e = getenv("FDC_FLATTEN_THREADS")) n_threads = (unsigned)std::max(1, atoi(e));
  if (total_nodes < 8192 || roots.size() < 4 * n_threads) n_threads = 1;
  if (n_threads == 1) {
    n += flatten_roots(scene, roots, 0, roots.size(), env, n < cap ? out + n : nullptr, n < cap ? cap - n : 0, &error);
  } else {
    std::vector<size_t> count(n_threads, 0), r0(n_threads + 1, 0);
    std::vector<const char*> errs(n_threads, nullptr);
    for (unsigned t = 0; t <= n_threads; t++) r0[t] = roots.size() * t / n_threads;
    auto run = [&](bool write, const std::vector<size_t>& offset) {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < n_threads; t++)
        th.emplace_back([&, t] {
          fdc_call* dst = nullptr;
          size_t room = 0;
          if (write && offset[t] < cap) { dst = out + offset[t]; room = std::min(count[t], cap - offset[t]); }
          const size_t c = flatten_roots(scene, roots, r0[t], r0[t + 1], env, dst, room, &errs[t]);
          if (!write) count[t] = c;
        });
      for (auto& x : th) x.join();
    };
    std::vector<size_t> offset(n_threads, 0);
    run(false, offset);
    for (unsigned t = 0; t < n_threads; t++) {
      if (errs[t] && !error) error = errs[t];
      offset[t] = n;
      n += count[t];
    }
    if (!error && n + 1 <= cap) run(true, offset);
  }
  {  // epilogue: restoreTransform
    Flattener F(n < cap ? out + n : nullptr, n < cap ? cap - n : 0, scene, env);
    F.restore();
    n += F.n;
  }
  *n_out = n;
  return error;
}

}  // namespace fdc

static_assert(sizeof(fdc_node_fill) == 16 && sizeof(fdc_node_shadow) == 36 && sizeof(fdc_node_stroke) == 24, "scene POD layout");
static_assert(sizeof(fdc_fig) == 248 && sizeof(fdc_glyph) == 32 && sizeof(fdc_draw_op) == 92 && sizeof(fdc_text_rect) == 32, "scene POD layout");

extern "C" int fdc_flatten_renders(const fdc_scene* scene, const fdc_flatten_env* env, fdc_call* out, size_t cap, size_t* n_out) {
  if (!scene || (!scene->lists && scene->n_lists) || !env || !n_out || (!out && cap)) return FDC_ERR_INVALID;
  const char* err = fdc::flatten_renders(*scene, *env, out, cap, n_out);
  if (err) return FDC_ERR_INVALID;
  return *n_out > cap ? FDC_ERR_CAPACITY : FDC_OK;
}
