// Per-tile SDF shade-and-blend kernel (sm_100a).
//
// One CTA of 8 warps owns up to 8 consecutive 16x16-pixel tiles = 64 blocks of 8x4 pixels, one pixel per thread; warps
// pull blocks from a shared-memory counter.  For its block a warp walks the tile's ordered list of 8-byte entries 32 at
// a time: every lane tests one entry's precomputed "overlaps my block" bit (ballot), then the warp shades the
// survivors one by one, all lanes on the same primitive, so every branch on mode/flags is warp-uniform and comes from
// the entry's dispatch bits before the primitive record is touched.  The pixel lives in registers as four floats holding exact RGBA8
// values (0..255) and is re-quantised after EVERY blended primitive, which is what GL's UNORM8 render target
// does (SURVEY.md 8a' trap 6).  Texture masks (clip stack) are evaluated analytically and kept per pixel as
// up to 15 UNORM8 levels (eight packed in two registers, the rest in shared memory), including the reference's a*a quirk (mask.frag:233 with
// GL_BLEND still enabled).
//
// The arithmetic restates src/figdraw/opengl/glsl/atlas.frag (main :252-405, sdRoundedBox :51-69,
// sdEllipticalRoundedBox :71-115, sdBezier :121-160, bezierStrokeSd :178-209, shadowProfile :211-216,
// evalFillColor :218-250), atlas_rect_mask.frag:222-237, mask.frag:186-234 and the blend state of
// utils/glutils.nim:150-154, in float32 with fused multiply-adds and MUFU approximations; results are
// compared against the CPU oracle within +-2 LSB (tests/test_gpu_parity.py).
//
// Work that provably cannot change a pixel is skipped (DESIGN.md "work reduction"): primitives before the last
// opaque full-coverage fill of a warp's block, and blends whose source alpha is exactly zero.
#include <cuda_runtime.h>

#include "fdc_kernels.h"

#ifndef FDC_SHADE_WAVES
#define FDC_SHADE_WAVES 6
#endif
#ifndef FDC_SHADE_MIN_BLOCKS
#define FDC_SHADE_MIN_BLOCKS 4
#endif
#ifndef FDC_SPLIT_KERNELS
#define FDC_SPLIT_KERNELS 1  // A/B switch: lean tiles in their own kernel when few tiles need the full loop
#endif
#ifndef FDC_LEAN_MIN_BLOCKS
#define FDC_LEAN_MIN_BLOCKS 6  // resident CTAs per SM the lean kernel is compiled for (register budget 65536 / (256 * n))
#endif
#ifndef FDC_QPTR
#define FDC_QPTR 1   // A/B switch: queue loop steps a pointer (0: index + base)
#endif
#ifndef FDC_LEAN_LOOP
#define FDC_LEAN_LOOP 1  // A/B switch: call-free loop for tiles that hold only unmasked fast primitives
#endif
#ifndef FDC_DEEP_MASK
#define FDC_DEEP_MASK 1  // A/B switch: texture-mask levels 9..15 in shared memory (0: levels 1..8 only, as in r01)
#endif
// (r01: staging primitive records into shared memory with cp.async or per-record cp.async.bulk/TMA + mbarrier was
// built and measured slower than L1-resident uniform loads -- profiles/r01_shade_staging.md, commit 4cc6a20.)

namespace fdc {

namespace {

__device__ __forceinline__ float sat(float x) { return __saturatef(x); }
// Pixel channels are kept BIASED: value + kBias with kBias = 1.5 * 2^23, where one ulp is exactly 1.  Any float add or
// FMA whose result lands in that binade is therefore rounded to an integer by the hardware (round-to-nearest-even),
// which is the UNORM8 store of the blended value for free; the low byte of the bit pattern is the UNORM8 value.
constexpr float kBias = 12582912.0f;
constexpr uint32_t kBiasBits = 0x4B400000u;
__device__ __forceinline__ float rint255(float x) {  // round to nearest (even) integer, |x| < 2^22
  return __fadd_rn(__fadd_rn(x, kBias), -kBias);
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float len2(float x, float y) { return fast_sqrt(fmaf(x, x, y * y)); }

// atlas.frag:51-69
__device__ __forceinline__ float sd_rounded_box(float px, float py, float bx, float by, float r0, float r1, float r2, float r3) {
  const float rr = px > 0.0f ? (py > 0.0f ? r0 : r1) : (py > 0.0f ? r2 : r3);
  const float qx = fabsf(px) - bx + rr, qy = fabsf(py) - by + rr;
  return fminf(fmaxf(qx, qy), 0.0f) + len2(fmaxf(qx, 0.0f), fmaxf(qy, 0.0f)) - rr;
}

// atlas.frag:71-79
__device__ float sd_ellipse(float px, float py, float rx, float ry) {
  const float sx = fmaxf(rx, 0.000001f), sy = fmaxf(ry, 0.000001f);
  const float k0 = len2(px / sx, py / sy);
  if (k0 <= 0.000001f) return -fminf(sx, sy);
  const float k1 = len2(px / (sx * sx), py / (sy * sy));
  return k0 * (k0 - 1.0f) / fmaxf(k1, 0.000001f);
}

// atlas.frag:96-115
__device__ float sd_elliptical_rounded_box(float px, float py, float bx, float by, float r0, float r1, float r2, float r3) {
  const float sel = px > 0.0f ? (py > 0.0f ? r0 : r1) : (py > 0.0f ? r2 : r3);
  if (sel < 0.0f) {
    const float r = -sel - 1.0f;
    return sd_rounded_box(px, py, bx, by, r, r, r, r);
  }
  const float pv = floorf(sel + 0.5f);
  const float hi = floorf(pv / 4096.0f);
  const float rx = (pv - 4096.0f * hi) * bx / 4095.0f, ry = hi * by / 4095.0f;
  if (rx <= 0.0f || ry <= 0.0f) {
    const float qx = fabsf(px) - bx, qy = fabsf(py) - by;
    return fminf(fmaxf(qx, qy), 0.0f) + len2(fmaxf(qx, 0.0f), fmaxf(qy, 0.0f));
  }
  if (rx == ry) return sd_rounded_box(px, py, bx, by, rx, rx, rx, rx);
  const float qx = fabsf(px) - bx + rx, qy = fabsf(py) - by + ry;
  if (qx > 0.0f && qy > 0.0f) return sd_ellipse(qx, qy, rx, ry);
  return fmaxf(qx - rx, qy - ry);
}

__device__ __forceinline__ float signf_(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// atlas.frag:121-160
__device__ float sd_bezier(float posx, float posy, float Ax, float Ay, float Bx, float By, float Cx, float Cy) {
  const float ax = Bx - Ax, ay = By - Ay;
  const float bx = Ax - 2.0f * Bx + Cx, by = Ay - 2.0f * By + Cy;
  const float bb = bx * bx + by * by;
  if (bb <= 0.000001f) {
    const float bax = Cx - Ax, bay = Cy - Ay;
    const float h = fminf(fmaxf(((posx - Ax) * bax + (posy - Ay) * bay) / fmaxf(bax * bax + bay * bay, 0.000001f), 0.0f), 1.0f);
    return len2(posx - (Ax + bax * h), posy - (Ay + bay * h));
  }
  const float cx = ax * 2.0f, cy = ay * 2.0f;
  const float dx = Ax - posx, dy = Ay - posy;
  const float kk = 1.0f / bb;
  const float kx = kk * (ax * bx + ay * by);
  const float ky = kk * (2.0f * (ax * ax + ay * ay) + (dx * bx + dy * by)) / 3.0f;
  const float kz = kk * (dx * ax + dy * ay);
  const float p = ky - kx * kx;
  const float p3 = p * p * p;
  const float q = kx * (2.0f * kx * kx - 3.0f * ky) + kz;
  float h = q * q + 4.0f * p3;
  float res;
  if (h >= 0.0f) {
    h = sqrtf(h);
    const float x0 = (h - q) / 2.0f, x1 = (-h - q) / 2.0f;
    const float r0 = signf_(x0) * powf(fabsf(x0), 1.0f / 3.0f), r1 = signf_(x1) * powf(fabsf(x1), 1.0f / 3.0f);
    const float t = fminf(fmaxf(r0 + r1 - kx, 0.0f), 1.0f);
    const float ex = dx + (cx + bx * t) * t, ey = dy + (cy + by * t) * t;
    res = ex * ex + ey * ey;
  } else {
    const float z = sqrtf(-p);
    const float v = acosf(fminf(fmaxf(q / (p * z * 2.0f), -1.0f), 1.0f)) / 3.0f;
    const float m = cosf(v), n = sinf(v) * 1.732050808f;
    const float t1 = fminf(fmaxf((m + m) * z - kx, 0.0f), 1.0f), t2 = fminf(fmaxf((-n - m) * z - kx, 0.0f), 1.0f);
    const float e1x = dx + (cx + bx * t1) * t1, e1y = dy + (cy + by * t1) * t1;
    const float e2x = dx + (cx + bx * t2) * t2, e2y = dy + (cy + by * t2) * t2;
    res = fminf(e1x * e1x + e1y * e1y, e2x * e2x + e2y * e2y);
  }
  return sqrtf(res);
}

__device__ __forceinline__ void safe_normalize(float vx, float vy, float fx, float fy, float& ox, float& oy) {
  const float len = len2(vx, vy);
  if (len <= 0.000001f) { ox = fx; oy = fy; } else { ox = vx / len; oy = vy / len; }
}

// atlas.frag:178-209
__device__ float bezier_stroke_sd(float dist, float px, float py, float Ax, float Ay, float Bx, float By, float Cx, float Cy,
                                  float halfW, int mode) {
  if (mode == FDC_SDF_BEZIER_STROKE_AA) return dist - halfW;
  float fbx, fby, stx, sty, etx, ety;
  safe_normalize(Cx - Ax, Cy - Ay, 1.0f, 0.0f, fbx, fby);
  safe_normalize(Bx - Ax, By - Ay, fbx, fby, stx, sty);
  safe_normalize(Cx - Bx, Cy - By, fbx, fby, etx, ety);
  const float startProj = (px - Ax) * stx + (py - Ay) * sty;
  const float endProj = (px - Cx) * etx + (py - Cy) * ety;
  const float trim = (mode == FDC_SDF_BEZIER_STROKE_SQUARE_AA) ? halfW : 0.0f;
  float tube = dist;
  if (mode == FDC_SDF_BEZIER_STROKE_SQUARE_AA) {
    if (startProj < 0.0f) tube = fminf(tube, fabsf((px - Ax) * sty - (py - Ay) * stx));
    if (endProj > 0.0f) tube = fminf(tube, fabsf((px - Cx) * ety - (py - Cy) * etx));
  }
  const float cap = fmaxf(-startProj - trim, endProj - trim);
  return fmaxf(tube - halfW, cap);
}

__device__ __forceinline__ float4 unpack255(uint32_t c) {
  return make_float4((float)(c & 255u), (float)((c >> 8) & 255u), (float)((c >> 16) & 255u), (float)(c >> 24));
}

__device__ __forceinline__ int wrap_i(int i, int n) {
  if ((n & (n - 1)) == 0) return i & (n - 1);  // power-of-two atlas (the usual case): no integer division
  int m = i % n;
  return m < 0 ? m + n : m;
}

// GL_LINEAR + GL_REPEAT on one RGBA8 level; (tu,tv) are texel coordinates minus 0.5.  Returns 0..255 per channel.
__device__ float4 tex_bilinear(const uint8_t* __restrict__ img, int size, float tu, float tv) {
  const float fx = floorf(tu), fy = floorf(tv);
  const float ax = tu - fx, ay = tv - fy;
  const int i0 = wrap_i((int)fx, size), i1 = wrap_i((int)fx + 1, size);
  const int j0 = wrap_i((int)fy, size), j1 = wrap_i((int)fy + 1, size);
  const uint32_t* im = reinterpret_cast<const uint32_t*>(img);
  const float4 t00 = unpack255(__ldg(im + (size_t)j0 * size + i0)), t10 = unpack255(__ldg(im + (size_t)j0 * size + i1));
  const float4 t01 = unpack255(__ldg(im + (size_t)j1 * size + i0)), t11 = unpack255(__ldg(im + (size_t)j1 * size + i1));
  float4 r;
  {
    const float top = fmaf(t10.x - t00.x, ax, t00.x), bot = fmaf(t11.x - t01.x, ax, t01.x);
    r.x = fmaf(bot - top, ay, top);
  }
  {
    const float top = fmaf(t10.y - t00.y, ax, t00.y), bot = fmaf(t11.y - t01.y, ax, t01.y);
    r.y = fmaf(bot - top, ay, top);
  }
  {
    const float top = fmaf(t10.z - t00.z, ax, t00.z), bot = fmaf(t11.z - t01.z, ax, t01.z);
    r.z = fmaf(bot - top, ay, top);
  }
  {
    const float top = fmaf(t10.w - t00.w, ax, t00.w), bot = fmaf(t11.w - t01.w, ax, t01.w);
    r.w = fmaf(bot - top, ay, top);
  }
  return r;
}

// GL_NEAREST + GL_REPEAT on level 0 (a `pixelate` context magnifies with it); (tu,tv) are texel coordinates minus 0.5.
__device__ __forceinline__ float4 tex_nearest(const uint8_t* __restrict__ img, int size, float tu, float tv) {
  const int i = wrap_i((int)floorf(tu + 0.5f), size), j = wrap_i((int)floorf(tv + 0.5f), size);
  return unpack255(__ldg(reinterpret_cast<const uint32_t*>(img) + (size_t)j * size + i));
}
// Level 0 through the magnification filter (lambda <= 0 and textureLod(.., 0)).
__device__ __forceinline__ float4 tex_mag(const AtlasView& at, float tu, float tv) {
  return at.pixelate ? tex_nearest(at.level[0], at.size, tu, tv) : tex_bilinear(at.level[0], at.size, tu, tv);
}

// texture(atlasTex, uv): min LINEAR_MIPMAP_LINEAR / mag LINEAR or NEAREST (glcontext.nim:157-169).  lambda = log2(rho).
__device__ float4 atlas_sample(const AtlasView& at, float tu, float tv, float lambda) {
  if (lambda <= 0.0f) return tex_mag(at, tu, tv);
  const int maxl = at.n_levels - 1;
  const float cu = tu + 0.5f, cv = tv + 0.5f;  // level-0 texel-space coordinate
  if (lambda >= (float)maxl) {
    const float sc = 1.0f / (float)(1 << maxl);
    return tex_bilinear(at.level[maxl], at.size >> maxl, cu * sc - 0.5f, cv * sc - 0.5f);
  }
  const int d1 = (int)floorf(lambda);
  const float fr = lambda - (float)d1;
  const float s1 = 1.0f / (float)(1 << d1), s2 = 0.5f * s1;
  const float4 a = tex_bilinear(at.level[d1], at.size >> d1, cu * s1 - 0.5f, cv * s1 - 0.5f);
  const float4 b = tex_bilinear(at.level[d1 + 1], at.size >> (d1 + 1), cu * s2 - 0.5f, cv * s2 - 0.5f);
  return make_float4(fmaf(b.x - a.x, fr, a.x), fmaf(b.y - a.y, fr, a.y), fmaf(b.z - a.z, fr, a.z), fmaf(b.w - a.w, fr, a.w));
}

struct Pixel {
  float r, g, b, a;   // exact UNORM8 values 0..255, biased by kBias
  uint32_t mlo, mhi;  // texture-mask levels 1..8, UNORM8 each
};

// Texture-mask levels 9..15 (GL nests mask textures without a limit, glcontext.nim:171-201; the depth field of a tile
// entry has four bits): two words per thread in shared memory, [word][thread] so a warp's accesses are conflict free.
// Only touched by primitives drawn at those depths; the register path of levels 1..8 pays one predicate.
__shared__ uint32_t s_deep_mask[2][256];

__device__ __forceinline__ float mask_get(const Pixel& px, int level) {  // level 1..15
  const int i = level - 1;
#if FDC_DEEP_MASK
  const uint32_t w = i < 4 ? px.mlo : (i < 8 ? px.mhi : s_deep_mask[(i - 8) >> 2][threadIdx.x]);
#else
  const uint32_t w = i < 4 ? px.mlo : px.mhi;
#endif
  return (float)((w >> ((i & 3) * 8)) & 255u);
}
__device__ __forceinline__ void mask_set(Pixel& px, int level, float v) {
  const int i = level - 1, sh = (i & 3) * 8;
  const uint32_t b = (uint32_t)(int)v & 255u;
  if (i < 4) px.mlo = (px.mlo & ~(255u << sh)) | (b << sh);
#if FDC_DEEP_MASK
  else if (i < 8) px.mhi = (px.mhi & ~(255u << sh)) | (b << sh);
  else {
    uint32_t* w = &s_deep_mask[(i - 8) >> 2][threadIdx.x];
    *w = (*w & ~(255u << sh)) | (b << sh);
  }
#else
  else px.mhi = (px.mhi & ~(255u << sh)) | (b << sh);
#endif
}

// General (rotated) quad: top-left-rule inside test on the two triangles (3,0,1),(2,3,1) and affine (s,t).
// Mirrors the oracle's raster_tri: exact integer edge functions on doubled coordinates.
struct GeneralHit {
  bool inside;
  float s, t;
  float dsdx, dsdy, dtdx, dtdy;
};
__device__ GeneralHit general_quad(const QuadGeom& g, int ix, int iy, int which) {
  GeneralHit h;
  h.inside = false;
  h.s = h.t = 0.0f;
  h.dsdx = h.dsdy = h.dtdx = h.dtdy = 0.0f;
  const float vs[4] = {0.0f, 1.0f, 1.0f, 0.0f}, vt[4] = {1.0f, 1.0f, 0.0f, 0.0f};
  const int tri[2][3] = {{3, 0, 1}, {2, 3, 1}};
  const long long px2 = 2ll * ix + 1, py2 = 2ll * iy + 1;
  {
    const int k = which;
    const int ia = tri[k][0], ib = tri[k][1], ic = tri[k][2];
    const long long ex[3] = {2ll * g.vx[ia], 2ll * g.vx[ib], 2ll * g.vx[ic]};
    const long long ey[3] = {2ll * g.vy[ia], 2ll * g.vy[ib], 2ll * g.vy[ic]};
    const long long area = (ex[1] - ex[0]) * (ey[2] - ey[0]) - (ey[1] - ey[0]) * (ex[2] - ex[0]);
    if (area == 0) return h;
    const long long sgn = area > 0 ? 1 : -1;
    bool in = true;
#pragma unroll
    for (int e = 0; e < 3; e++) {
      const int n = (e + 1) % 3;
      const long long ea = -sgn * (ey[n] - ey[e]), eb = sgn * (ex[n] - ex[e]);
      const long long v = ea * (px2 - ex[e]) + eb * (py2 - ey[e]);
      if (v < 0 || (v == 0 && !(ea > 0 || (ea == 0 && eb > 0)))) in = false;
    }
    if (!in) return h;
    const float Ax = (float)g.vx[ia], Ay = (float)g.vy[ia], Bx = (float)g.vx[ib], By = (float)g.vy[ib];
    const float Cx = (float)g.vx[ic], Cy = (float)g.vy[ic];
    const float inv_area = 1.0f / ((Bx - Ax) * (Cy - Ay) - (By - Ay) * (Cx - Ax));
    const float bcy = (Cy - Ay) * inv_area, bby = (By - Ay) * inv_area, bcx = (Cx - Ax) * inv_area, bbx = (Bx - Ax) * inv_area;
    h.dsdx = (vs[ib] - vs[ia]) * bcy - (vs[ic] - vs[ia]) * bby;
    h.dsdy = (vs[ic] - vs[ia]) * bbx - (vs[ib] - vs[ia]) * bcx;
    h.dtdx = (vt[ib] - vt[ia]) * bcy - (vt[ic] - vt[ia]) * bby;
    h.dtdy = (vt[ic] - vt[ia]) * bbx - (vt[ib] - vt[ia]) * bcx;
    const float rx = (float)ix + 0.5f - Ax, ry = (float)iy + 0.5f - Ay;
    h.s = vs[ia] + (rx * h.dsdx + ry * h.dsdy);
    h.t = vt[ia] + (rx * h.dtdx + ry * h.dtdy);
    h.inside = true;
  }
  return h;
}

// Vertex colours BL,BR,TR,TL interpolated per triangle: (TL,BL,BR) where t >= s, (TR,TL,BR) otherwise.
__device__ __forceinline__ float4 vertex_color(const uint4 c, float s, float t, bool lower) {
  const float4 bl = unpack255(c.x), br = unpack255(c.y), tr = unpack255(c.z), tl = unpack255(c.w);
  float4 o;
  o.x = lower ? fmaf(bl.x - tl.x, t, fmaf(br.x - bl.x, s, tl.x)) : fmaf(br.x - tr.x, t, fmaf(tr.x - tl.x, s, tl.x));
  o.y = lower ? fmaf(bl.y - tl.y, t, fmaf(br.y - bl.y, s, tl.y)) : fmaf(br.y - tr.y, t, fmaf(tr.y - tl.y, s, tl.y));
  o.z = lower ? fmaf(bl.z - tl.z, t, fmaf(br.z - bl.z, s, tl.z)) : fmaf(br.z - tr.z, t, fmaf(tr.z - tl.z, s, tl.z));
  o.w = lower ? fmaf(bl.w - tl.w, t, fmaf(br.w - bl.w, s, tl.w)) : fmaf(br.w - tr.w, t, fmaf(tr.w - tl.w, s, tl.w));
  return o;
}

// evalFillColor, atlas.frag:218-250 (colours in 0..255)
__device__ __forceinline__ float4 linear3_color(uint32_t c0, uint32_t c1, uint32_t c2, int fill_mode, float mid, float s, float t) {
  float tt = fill_mode == 1 ? s : (fill_mode == 2 ? t : (fill_mode == 3 ? 0.5f * (s + t) : 0.5f * (s + (1.0f - t))));
  tt = sat(tt);
  const float4 a = unpack255(c0), m = unpack255(c1), b = unpack255(c2);
  if (tt <= mid) {
    const float k = tt / mid;
    return make_float4(fmaf(m.x - a.x, k, a.x), fmaf(m.y - a.y, k, a.y), fmaf(m.z - a.z, k, a.z), fmaf(m.w - a.w, k, a.w));
  }
  const float k = (tt - mid) / (1.0f - mid);
  return make_float4(fmaf(b.x - m.x, k, m.x), fmaf(b.y - m.y, k, m.y), fmaf(b.z - m.z, k, m.z), fmaf(b.w - m.w, k, m.w));
}

// atlas_rect_mask.frag:222-237
__device__ float rect_mask_alpha(const RectMaskRec& rm, float aa, float px, float py) {
  if (rm.hx < 0.0f || rm.hy < 0.0f) return 1.0f;
  const float lx = fmaf(rm.ax, px, rm.ay * py) + rm.az, ly = fmaf(rm.bx, px, rm.by * py) + rm.bz;
  const float qx = lx - rm.cx, qy = -(ly - rm.cy);
  const float dist = rm.elliptical > 0.5f ? sd_elliptical_rounded_box(qx, qy, rm.hx, rm.hy, rm.r0, rm.r1, rm.r2, rm.r3)
                                          : sd_rounded_box(qx, qy, rm.hx, rm.hy, rm.r0, rm.r1, rm.r2, rm.r3);
  return 1.0f - sat(fmaf(aa, dist, 0.5f));
}

// Out-of-line versions for the fast path: rare, large (elliptical SDF with IEEE divisions / four texel fetches), scalar in,
// scalar out -- inlining them into the visit loop costs the hot SDF path registers.
__device__ __noinline__ float rect_mask_alpha_call(const RectMaskRec* __restrict__ rm, float aa, float px, float py) {
  return rect_mask_alpha(*rm, aa, px, py);
}
__device__ __noinline__ float4 tex_mag_call(const AtlasView* __restrict__ at, float tu, float tv) { return tex_mag(*at, tu, tv); }

__device__ __forceinline__ void blend(Pixel& px, float sr, float sg, float sb, float sa) {
  // rgb = s*sa + d*(1-sa); a = sa + da*(1-sa)  (glBlendFuncSeparate, glutils.nim:150-154), then UNORM8 store:
  // d + (s - d)*sa evaluated by one FMA into the biased binade = rounded to the UNORM8 grid.  sa == 0 is a no-op.
  px.r = fmaf(sr - (px.r - kBias), sa, px.r);
  px.g = fmaf(sg - (px.g - kBias), sa, px.g);
  px.b = fmaf(sb - (px.b - kBias), sa, px.b);
  px.a = fmaf(255.0f - (px.a - kBias), sa, px.a);
}

// Fast path (PF_FAST primitives): axis-aligned quads -- rounded boxes with circular corners in ClipAA / AnnularAA /
// DropShadow mode, atlas / MSDF quads at <= 1 texel per pixel -- as content, as ClipAA mask writes, or under a rect mask.
// `info` is the TileEntry word; `full`: the warp's whole block lies in the primitive's inner rect (coverage exactly 1).
template <bool kMasked, bool kInlineTex>
__device__ __forceinline__ void shade_fast(const float4* __restrict__ S, const PrimExt* __restrict__ E, const AtlasView& at,
                                           const RectMaskRec* __restrict__ rectmasks, uint32_t info, bool full, float fx, float fy,
                                           Pixel& px) {
  // S: q0..q4 of the primitive in global memory; all lanes load the same address (L1-resident, one transaction)
  float4 col;
  if (info & TE_SOLID) {
    col = __ldg(S + 4);
  } else {
    const float4* X = reinterpret_cast<const float4*>(E);
    const float4 a0 = __ldg(X + 1), d0 = __ldg(X + 2), a1 = __ldg(X + 3);
    if ((info & (TE_GRAD3 | TE_TEX)) == TE_GRAD3) {
      const float4 e0 = __ldg(X + 0), d1 = __ldg(X + 4);
      const float tt = sat(fmaf(fx, e0.x, fmaf(fy, e0.y, e0.z)));
      const bool lo = tt <= e0.w;
      col.x = fmaf(lo ? d0.x : d1.x, tt, lo ? a0.x : a1.x);
      col.y = fmaf(lo ? d0.y : d1.y, tt, lo ? a0.y : a1.y);
      col.z = fmaf(lo ? d0.z : d1.z, tt, lo ? a0.z : a1.z);
      col.w = fmaf(lo ? d0.w : d1.w, tt, lo ? a0.w : a1.w);
    } else {
      col.x = fmaf(d0.x, fx, fmaf(a1.x, fy, a0.x));
      col.y = fmaf(d0.y, fx, fmaf(a1.y, fy, a0.y));
      col.z = fmaf(d0.z, fx, fmaf(a1.z, fy, a0.z));
      col.w = fmaf(d0.w, fx, fmaf(a1.w, fy, a0.w));
    }
  }
  if (FDC_FAST_TEX && (info & TE_TEX)) {
    // Atlas-sampling quads.  kind 0: texture(atlas, uv) * vertex colour, magnified or 1:1 (atlas.frag:284-292);
    // kind 1/2: MSDF / MTSDF coverage from textureLod(.., 0) (atlas.frag:294-318), TE_GRAD3 = annular stroke.
    const float4 q0 = __ldg(S + 0), q7 = __ldg(S + 7);  // texel map (u0,du,v0,dv); pixel -> (s,t) map
    const int4 q6 = __ldg(reinterpret_cast<const int4*>(S) + 6);
    const int ix = (int)fx, iy = (int)fy;
    const bool inside = ix >= (int16_t)(q6.x & 0xFFFF) && iy >= (int16_t)(q6.x >> 16) && ix < (int16_t)(q6.y & 0xFFFF) &&
                        iy < (int16_t)(q6.y >> 16);
    const float s = fmaf(fx, q7.x, q7.y), t = fmaf(fy, q7.z, q7.w);
    float4 tex = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside) {
      // lean tiles (no call anywhere in their loop) fetch in line; the full loop keeps the fetch out of line so that it
      // does not cost the rounded-box path registers
      if (kInlineTex) tex = tex_mag(at, fmaf(s, q0.y, q0.x), fmaf(t, q0.w, q0.z));
      else tex = tex_mag_call(&at, fmaf(s, q0.y, q0.x), fmaf(t, q0.w, q0.z));
    }
    const uint32_t kind = (info >> TE_KIND_SHIFT) & 3u;
    float sr = col.x, sg = col.y, sb = col.z, sa;
    if (kind == 0u) {
      sr = tex.x * col.x * (1.0f / 255.0f); sg = tex.y * col.y * (1.0f / 255.0f); sb = tex.z * col.z * (1.0f / 255.0f);
      sa = tex.w * col.w * (1.0f / 255.0f) * (1.0f / 255.0f);
    } else {
      const float4 q1 = __ldg(S + 1), q3 = __ldg(S + 3);  // q1.y stroke weight; q3 = (pxRange, sdThreshold, aa, screenPxRange)
      const float sd = (kind == 2u ? tex.w : fmaxf(fminf(tex.x, tex.y), fminf(fmaxf(tex.x, tex.y), tex.z))) * (1.0f / 255.0f);
      const float spd = q3.w * (sd - q3.y);
      const float cov = (info & TE_GRAD3) ? sat(fmaxf(q1.y, 0.0f) * 0.5f - fabsf(spd) + 0.5f) : sat(spd + 0.5f);
      sa = col.w * (1.0f / 255.0f) * cov;
    }
    if (kMasked && (info & (15u << TE_DEPTH_SHIFT))) sa *= mask_get(px, (int)((info >> TE_DEPTH_SHIFT) & 15u)) * (1.0f / 255.0f);
    if (FDC_FAST_MASK && kMasked && (info & TE_RECTMASK)) {
      const float aa = __ldg(S + 3).z;
      if (inside) sa *= rect_mask_alpha_call(rectmasks + (((uint32_t)q6.w & 0xFFFFu) - 1u), aa, fx + 0.5f, fy + 0.5f);
    }
    blend(px, sr, sg, sb, inside ? sa : 0.0f);
    return;
  }
  if (full && (info & TE_OCCLUDER)) {
    // opaque, coverage exactly 1 on the whole block: dst*(1-1) vanishes, the store is round(src) whatever dst was.
    // (Earlier primitives were skipped for this block, so the result must not depend on dst even in the last ulp.)
    px.r = col.x + kBias; px.g = col.y + kBias; px.b = col.z + kBias; px.a = 255.0f + kBias;
    return;
  }
  if (FDC_FAST_MASK && kMasked && (info & TE_MASKB)) mask_set(px, (int)((info >> TE_DEPTH_SHIFT) & 15u), 0.0f);  // glClear(0) of the level, glcontext.nim:1901-1902
  float sa = col.w * (1.0f / 255.0f);
  bool inside = true;
  if (!full) {
    const float4 q0 = __ldg(S + 0), q1 = __ldg(S + 1), q2 = __ldg(S + 2), q3 = __ldg(S + 3);
    const float ppx = fmaf(fx, q0.x, q0.y), ppy = fmaf(fy, q0.z, q0.w);  // (p.x, -p.y)
    const float apx = fabsf(ppx), apy = fabsf(ppy);
    inside = apx < q1.x && apy < q1.y;  // pixel centre inside the ceil'd quad
    const float rr = ppx > 0.0f ? (ppy > 0.0f ? q2.x : q2.y) : (ppy > 0.0f ? q2.z : q2.w);
    const float qx = apx - q1.z + rr, qy = apy - q1.w + rr;
    const float mx = fmaxf(qx, 0.0f), my = fmaxf(qy, 0.0f);
    const float dist = fminf(fmaxf(qx, qy), 0.0f) + fast_sqrt(fmaf(mx, mx, my * my)) - rr;
    const uint32_t kind = (info >> TE_KIND_SHIFT) & 3u;
    float cov;
    if (kind == 2u) {  // DropShadow: sd > 0 ? exp(-.5 (sd/sigma)^2) : 1
      const float sd = fmaxf(dist - q3.y, 0.0f);
      cov = fast_ex2(q3.w * sd * sd);
    } else if (kind == 0u) {  // ClipAA
      cov = sat(fmaf(-q3.z, dist, 0.5f));
    } else {  // AnnularAA
      const float f = q3.x * 0.5f;
      cov = sat(fmaf(-q3.z, fabsf(dist + f) - f, 0.5f));
    }
    sa = inside ? sa * cov : 0.0f;
  }
  if (FDC_FAST_MASK && kMasked && (info & TE_MASKW)) {
    // mask.frag:219-233 for a ClipAA clip shape: alpha = cov * colour.a * previous level; the R8 target is written
    // with blending still on, so the stored value is a*a + m*(1-a) (SURVEY 8a' trap 1), quantised.
    const int depth = (int)((info >> TE_DEPTH_SHIFT) & 15u);
    float al = sa;
    if (depth > 1) al *= mask_get(px, depth - 1) * (1.0f / 255.0f);
    if (inside) {
      const float m = mask_get(px, depth);
      mask_set(px, depth, rint255(fmaf(al, 255.0f * al, m * (1.0f - al))));
    }
    return;
  }
  if (kMasked && (info & (15u << TE_DEPTH_SHIFT))) sa *= mask_get(px, (int)((info >> TE_DEPTH_SHIFT) & 15u)) * (1.0f / 255.0f);
  if (FDC_FAST_MASK && kMasked && (info & TE_RECTMASK)) {
    const uint32_t aux = (uint32_t)__ldg(reinterpret_cast<const int4*>(S) + 6).w;
    if (sa > 0.0f) sa *= rect_mask_alpha_call(rectmasks + ((aux & 0xFFFFu) - 1u), __ldg(S + 3).z, fx + 0.5f, fy + 0.5f);
  }
  blend(px, col.x, col.y, col.z, sa);
}

__device__ __noinline__ Pixel shade_prim(const ShadeArgs* __restrict__ ap, const Prim* __restrict__ P, int ix, int iy, Pixel px) {
  const ShadeArgs& a = *ap;
  const float4* Q = reinterpret_cast<const float4*>(P);
  const int4 q6 = __ldg(reinterpret_cast<const int4*>(P) + 6);
  const uint32_t flags = (uint32_t)q6.z;
  const int mode = (int)(flags & PF_MODE_MASK);
  const int depth = (int)((flags & PF_DEPTH_MASK) >> PF_DEPTH_SHIFT);
  const bool mask_write = flags & PF_MASK_WRITE;
  if (flags & PF_MASK_BEGIN) mask_set(px, depth, 0.0f);  // glClear(0) of the mask level, glcontext.nim:1901-1902
  if (flags & PF_CLEAR_ONLY) return px;                   // first draw of a multi-draw mask level had an empty quad

  int bx0 = (int16_t)(q6.x & 0xFFFF), by0 = (int16_t)(q6.x >> 16), bx1 = (int16_t)(q6.y & 0xFFFF), by1 = (int16_t)(q6.y >> 16);
  if (flags & PF_MASK_WIDE) {
    // binned over the parent's whole clip box (to clear the level everywhere); its own clipped bbox sits in ix0..iy1
    const int4 q5 = __ldg(reinterpret_cast<const int4*>(P) + 5);
    bx0 = (int16_t)(q5.z & 0xFFFF); by0 = (int16_t)(q5.z >> 16); bx1 = (int16_t)(q5.w & 0xFFFF); by1 = (int16_t)(q5.w >> 16);
  }
  const bool in_box = ix >= bx0 && ix < bx1 && iy >= by0 && iy < by1;
  const float4 q0 = __ldg(Q + 7);  // (su, ou, sv, ov)
  // A rotated / arbitrary quad is two triangles (3,0,1),(2,3,1) drawn one after the other (glcontext.nim:418-429): a
  // pixel inside both (folded quads) is shaded and blended twice, exactly as GL does.
  const int n_tri = (flags & PF_GENERAL) ? 2 : 1;
  for (int tri = 0; tri < n_tri; tri++) {
  bool inside = in_box;
  float s, t;
  float dsdx = q0.x, dsdy = 0.0f, dtdx = 0.0f, dtdy = q0.z;
  if (flags & PF_GENERAL) {
    GeneralHit h;
    h.inside = false;
    if (inside) h = general_quad(a.geoms[__float_as_uint(q0.x)], ix, iy, tri);
    inside = inside && h.inside;
    s = h.s; t = h.t;
    dsdx = h.dsdx; dsdy = h.dsdy; dtdx = h.dtdx; dtdy = h.dtdy;
  } else {
    s = fmaf((float)ix, q0.x, q0.y);
    t = fmaf((float)iy, q0.z, q0.w);
  }
  if (!__any_sync(0xFFFFFFFFu, inside)) continue;

  const float4 q1 = __ldg(Q + 1), q2 = __ldg(Q + 2), q3 = __ldg(Q + 3);
  const uint4 q4 = __ldg(reinterpret_cast<const uint4*>(P) + 4);
  const uint4 q5 = __ldg(reinterpret_cast<const uint4*>(P) + 5);
  const float aa = q3.z;
  const int fill_mode = (int)((flags & PF_FILLMODE_MASK) >> PF_FILLMODE_SHIFT);

  // fill colour (0..255)
  float4 col;
  if (fill_mode != 0) col = linear3_color(q4.x, q5.x, q5.y, fill_mode, q3.y, s, t);
  else if (flags & PF_SOLID) col = __ldg(Q + 4);
  else col = vertex_color(q4, s, t, (flags & PF_GENERAL) ? tri == 0 : t >= s);

  // p = (uv - .5) * 2 * quadHalf ; the SDF is evaluated at (p.x, -p.y)
  const float ppx = (s - 0.5f) * 2.0f * q1.x, ppy = (t - 0.5f) * 2.0f * q1.y;
  float cov = 0.0f;        // coverage alpha
  float sr = col.x, sg = col.y, sb = col.z, salpha = col.w;  // source colour 0..255 and its alpha 0..255

  const bool is_bezier = mode >= FDC_SDF_BEZIER_STROKE_AA && mode <= FDC_SDF_BEZIER_STROKE_SQUARE_AA;
  const bool is_msdf = mode >= FDC_SDF_MSDF && mode <= FDC_SDF_MTSDF_ANNULAR;
  if (mode == FDC_SDF_ATLAS) {
    const float4 q7 = __ldg(Q + 0);  // atlas texel map
    const float tu = fmaf(s, q7.y, q7.x), tv = fmaf(t, q7.w, q7.z);
    float lambda = q3.w;
    if (flags & PF_GENERAL) {
      const float rx = len2(q7.y * dsdx, q7.w * dtdx), ry = len2(q7.y * dsdy, q7.w * dtdy);
      const float rho = fmaxf(rx, ry);
      lambda = rho > 0.0f ? log2f(rho) : -1000.0f;
    }
    float4 tex = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside) tex = atlas_sample(a.atlas, tu, tv, lambda);
    if (mask_write) {
      cov = tex.w * (1.0f / 255.0f);  // mask.frag:196-197: alpha = tex.a * color.a
    } else {
      sr = tex.x * col.x * (1.0f / 255.0f); sg = tex.y * col.y * (1.0f / 255.0f); sb = tex.z * col.z * (1.0f / 255.0f);
      salpha = tex.w * col.w * (1.0f / 255.0f);
      cov = 1.0f;
    }
  } else if (is_msdf && !mask_write) {
    const float4 q7 = __ldg(Q + 0);  // atlas texel map
    const float tu = fmaf(s, q7.y, q7.x), tv = fmaf(t, q7.w, q7.z);
    float4 tex = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside) tex = tex_mag(a.atlas, tu, tv);  // textureLod(.., 0): lambda = 0 selects the magnification filter
    const bool mtsdf = mode == FDC_SDF_MTSDF || mode == FDC_SDF_MTSDF_ANNULAR;
    const bool stroke = mode == FDC_SDF_MSDF_ANNULAR || mode == FDC_SDF_MTSDF_ANNULAR;
    const float sd = (mtsdf ? tex.w : fmaxf(fminf(tex.x, tex.y), fminf(fmaxf(tex.x, tex.y), tex.z))) * (1.0f / 255.0f);
    float spr = q3.w;
    if (flags & PF_GENERAL) {
      const float fwx = fabsf(q7.y * dsdx) + fabsf(q7.y * dsdy), fwy = fabsf(q7.w * dtdx) + fabsf(q7.w * dtdy);
      spr = fmaxf(0.5f * (q3.x / fwx + q3.x / fwy), 1.0f);
    }
    const float spd = spr * (sd - q3.y);
    cov = stroke ? sat(fmaxf(q1.y, 0.0f) * 0.5f - fabsf(spd) + 0.5f) : sat(spd + 0.5f);
  } else {
    float dist;
    const bool ell = flags & PF_ELLIPTICAL;
    const bool inset = mode == FDC_SDF_INSET_SHADOW && !mask_write;
    const float shx = inset ? q1.x : q1.z, shy = inset ? q1.y : q1.w;
    if (is_bezier) dist = sd_bezier(ppx, ppy, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w);
    else if (ell) dist = sd_elliptical_rounded_box(ppx, -ppy, shx, shy, q2.x, q2.y, q2.z, q2.w);
    else dist = sd_rounded_box(ppx, -ppy, shx, shy, q2.x, q2.y, q2.z, q2.w);

    if (mask_write) {
      // mask.frag:199-226
      if (is_bezier) dist = bezier_stroke_sd(dist, ppx, ppy, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, fmaxf(q3.x, 0.0f) * 0.5f, mode);
      if (mode == FDC_SDF_ANNULAR_AA) {
        const float hw = fmaxf(q3.x, 0.0f) * 0.5f;
        dist = fabsf(dist + hw) - hw;
      }
      cov = 1.0f - sat(fmaf(aa, dist, 0.5f));
    } else {
      switch (mode) {
        case FDC_SDF_BEZIER_STROKE_AA:
        case FDC_SDF_BEZIER_STROKE_BUTT_AA:
        case FDC_SDF_BEZIER_STROKE_SQUARE_AA: {
          const float sd = bezier_stroke_sd(dist, ppx, ppy, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, fmaxf(q3.x, 0.0f) * 0.5f, mode);
          cov = 1.0f - sat(fmaf(aa, sd, 0.5f));
          break;
        }
        case FDC_SDF_ANNULAR: {
          const float f = q3.x * 0.5f;
          cov = (fabsf(dist + f) - f) < 0.0f ? 1.0f : 0.0f;
          break;
        }
        case FDC_SDF_ANNULAR_AA: {
          const float f = q3.x * 0.5f;
          cov = 1.0f - sat(fmaf(aa, fabsf(dist + f) - f, 0.5f));
          break;
        }
        case FDC_SDF_DROP_SHADOW: {
          const float spread = fill_mode == 0 ? q3.y : 0.0f;
          const float sd = dist - spread;
          cov = sd > 0.0f ? fminf(exp2f(q3.w * sd * sd), 1.0f) : 1.0f;
          break;
        }
        case FDC_SDF_DROP_SHADOW_AA: {
          const float spread = fill_mode == 0 ? q3.y : 0.0f;
          const float sd = dist - spread;
          cov = sd >= 0.0f ? fminf(exp2f(q3.w * sd * sd), 1.0f) : 1.0f - sat(fmaf(aa, dist, 0.5f));
          break;
        }
        case FDC_SDF_INSET_SHADOW: {
          // atlas.frag:364-379: clip against the quad-sized box, shadow from the same box offset by params.zw
          const float clip_a = 1.0f - sat(fmaf(aa, dist, 0.5f));
          const float sx = ppx - q1.z, sy = -ppy - (-q1.w);
          const float sdist = ell ? sd_elliptical_rounded_box(sx, sy, q1.x, q1.y, q2.x, q2.y, q2.z, q2.w)
                                  : sd_rounded_box(sx, sy, q1.x, q1.y, q2.x, q2.y, q2.z, q2.w);
          const float spread = fill_mode == 0 ? q3.y : 0.0f;
          const float sd = sdist + spread;
          const float ia = sd < 0.0f ? fminf(exp2f(q3.w * sd * sd), 1.0f) : 1.0f;
          cov = clip_a * ia;
          break;
        }
        case FDC_SDF_BACKDROP_BLUR: {
          cov = 1.0f - sat(fmaf(aa, dist, 0.5f));
          if (inside && a.backdrop) {
            const float4 b = unpack255(__ldg(reinterpret_cast<const uint32_t*>(a.backdrop) + (size_t)iy * a.frame.W + ix));
            sr = b.x; sg = b.y; sb = b.z; salpha = b.w;
          }
          break;
        }
        default: cov = 1.0f - sat(fmaf(aa, dist, 0.5f)); break;
      }
    }
  }

  if (mask_write) {
    // alpha = cov * color.a * prevMask ; R8 target with blending on: r = a*a + dst*(1-a)   (SURVEY 8a' trap 1)
    float al = cov * col.w * (1.0f / 255.0f);
    if (depth > 1) al *= mask_get(px, depth - 1) * (1.0f / 255.0f);
    if (inside) {
      const float m = mask_get(px, depth);
      mask_set(px, depth, rint255(fmaf(al, 255.0f * al, m * (1.0f - al))));
    }
    continue;
  }
  float sa = salpha * (1.0f / 255.0f) * cov;
  if (depth > 0) sa *= mask_get(px, depth) * (1.0f / 255.0f);
  if (flags & PF_RECTMASK) {
    const RectMaskRec rm = a.rectmasks[((uint32_t)q6.w & 0xFFFFu) - 1u];
    sa *= rect_mask_alpha(rm, aa, (float)ix + 0.5f, (float)iy + 0.5f);
  }
  if (inside && sa > 0.0f) blend(px, sr, sg, sb, sa);
  }  // triangles
  return px;
}

}  // namespace

// One warp's walk over the ordered entry list of its tile for one 8x4-pixel block.
//
// kLean: the fine binner found nothing in this tile but unmasked PF_FAST primitives (tile_count bit 31 clear).  That
// instantiation contains no call at all -- the general path (shade_prim) and the out-of-line helpers are the only
// reason the other loop keeps a stack frame and spills its loop state around every visit (r02 profile: 4 local-memory
// instructions and 13 % of the stall samples per visit on a path cfg5 never takes) -- and no mask code.
//
// Surviving entries of a 32-entry step are compacted into a per-warp queue in shared memory with the primitive's
// ADDRESS already formed by the lane that owned the entry (one 64-bit multiply-add per lane, in parallel), so a visit
// starts with one broadcast LDS.128 instead of find-first-set + two shuffles + address arithmetic (31 -> ~12
// instructions of loop management per visit).
template <bool kLean>
__device__ __forceinline__ void walk_list(const ShadeArgs& a, const uint2* __restrict__ list, uint32_t n, uint32_t start, uint32_t ov_bit,
                                          uint32_t full_bit, uint4* __restrict__ queue, int lane, int ix, int iy, float fx, float fy,
                                          Pixel& px) {
  const uint32_t lt = (1u << lane) - 1u;
  for (uint32_t base = start; base < n; base += 32) {
    const uint32_t idx = base + lane;
    uint2 e = make_uint2(0u, 0u);
    if (idx < n) e = __ldg(&list[idx]);
    const bool hit = (e.y & ov_bit) != 0u;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
    if (!kLean && a.stats) {
      const uint32_t mf = __ballot_sync(0xFFFFFFFFu, hit && (e.y & full_bit));
      const uint32_t ms = __ballot_sync(0xFFFFFFFFu, hit && !(e.y & TE_FAST));
      if (lane == 0) {
        atomicAdd(&a.stats[3], 1ull);
        atomicAdd(&a.stats[0], (unsigned long long)__popc(m));
        atomicAdd(&a.stats[1], (unsigned long long)__popc(mf));
        atomicAdd(&a.stats[2], (unsigned long long)__popc(ms));
      }
    }
    if (m == 0u) continue;
    if (hit) {
      const unsigned long long addr = reinterpret_cast<unsigned long long>(a.prims + e.x);
      queue[__popc(m & lt)] = make_uint4((uint32_t)addr, (uint32_t)(addr >> 32), e.y, e.x);
    }
    __syncwarp();
#if FDC_QPTR
    const uint4* __restrict__ qp = queue;
    const uint4* const qe = queue + __popc(m);
    for (; qp != qe; qp++) {
      const uint4 q = *qp;  // same address in every lane: one broadcast
#else
    const int cnt = __popc(m);
    for (int k = 0; k < cnt; k++) {
      const uint4 q = queue[k];  // same address in every lane: one broadcast
#endif
      const float4* S = reinterpret_cast<const float4*>(((unsigned long long)q.y << 32) | (unsigned long long)q.x);
      const uint32_t info = q.z;
      const bool full = (info & full_bit) != 0u;
      if (kLean) {
        shade_fast<false, true>(S, a.exts + q.w, a.atlas, a.rectmasks, info, full, fx, fy, px);
      } else if (info & TE_FAST) {
        if (info & ((15u << TE_DEPTH_SHIFT) | TE_RECTMASK)) shade_fast<true, false>(S, a.exts + q.w, a.atlas, a.rectmasks, info, full, fx, fy, px);
        else shade_fast<false, false>(S, a.exts + q.w, a.atlas, a.rectmasks, info, full, fx, fy, px);
      } else {
        px = shade_prim(&a, reinterpret_cast<const Prim*>(S), ix, iy, px);
      }
    }
    __syncwarp();  // the queue is rewritten by the next step
  }
}

// One CTA owns kTilesPerCta consecutive tiles = 8*kTilesPerCta blocks of 8x4 pixels.  Its 8 warps pull blocks from a
// shared-memory counter instead of being pinned to one block of one tile: a warp that finishes a cheap block moves
// on immediately (no intra-CTA tail: the slowest block of a tile used to hold 7 idle warps' registers), while the
// warps of a CTA still work on neighbouring blocks of the same tiles at the same time, so the primitive records
// they load stay shared in L1.

// Two kernels share this body.  The LEAN kernel shades tiles whose lists hold only unmasked fast primitives; it contains
// no call, no mask code and no general path, so it is compiled for a smaller register budget (40 instead of 64
// registers, no stack frame) and keeps 48 instead of 32 warps resident per SM.  The FULL kernel contains both loops.
// Which one shades the frame is decided on the device from the fine binner's count of full-path tiles, identically by
// both launches:
//   no tile needs the full loop (cfg3, cfg5): the lean kernel shades everything, the full launch returns at once;
//   otherwise (masks, elliptical corners, rotated quads ... -- cfg2, cfg4): the full kernel shades every tile, picking
//          the loop per tile, and the lean launch returns at once.  (Splitting a mixed frame between the two launches
//          was measured slower: they run back to back on one stream, so the tail of the heavy full-path tiles is no
//          longer hidden behind lean work -- cfg4 0.387 -> 0.430 ms.)
template <int kTilesPerCta, bool kLeanKernel>
__global__ void __launch_bounds__(256, kLeanKernel ? FDC_LEAN_MIN_BLOCKS : FDC_SHADE_MIN_BLOCKS) shade_kernel(const __grid_constant__ ShadeArgs a) {
  __shared__ uint32_t s_next, s_dirty;
  __shared__ uint4 s_queue[8][32];  // per warp: surviving entries of the current 32-entry step (address, info, index)
  // finished pixels of the CTA's tiles, written out at the end as 16-byte row chunks (see copy-out below)
  __shared__ __align__(16) uint32_t s_out[kTilesPerCta][kTileW * kTileH];
  if (a.counters[kCntStickyOverflow] != 0) return;  // a bin list overflowed in this or an earlier segment: host regrows and replays the frame
  // debug counters live in the full loop only
  const bool all_lean = FDC_LEAN_LOOP && FDC_SPLIT_KERNELS && a.stats == nullptr && a.counters[kCntFullTiles] == 0u;
  if (kLeanKernel != all_lean) return;
  if (threadIdx.x == 0) { s_next = 0; s_dirty = 0; }
#if FDC_DEEP_MASK
  // levels 9..15 start at 0 like the register levels; within a block every level is cleared by its first mask
  // primitive before anything reads it, so once per CTA is enough
  if (!kLeanKernel) s_deep_mask[0][threadIdx.x] = s_deep_mask[1][threadIdx.x] = 0;
#endif
  __syncthreads();
  const FrameView& f = a.frame;
  const int lane = threadIdx.x & 31;
  const int n_tiles = f.tiles_x * (f.ty1 - f.ty0);
  const int tile0 = blockIdx.x * kTilesPerCta;
  const uint32_t n_blocks = (uint32_t)min(kTilesPerCta, n_tiles - tile0) * 8u;
  uint32_t* fb32 = reinterpret_cast<uint32_t*>(a.fb);
  uint4* queue = s_queue[threadIdx.x >> 5];

  for (;;) {
    uint32_t blk = 0;
    if (lane == 0) blk = atomicAdd(&s_next, 1u);
    blk = __shfl_sync(0xFFFFFFFFu, blk, 0);
    if (blk >= n_blocks) break;
    const int tile = tile0 + (int)(blk >> 3), sub = (int)(blk & 7u);
    // (a float reciprocal instead of this integer division was measured slower: register allocation, profiles/r02_shade_experiments.md)
    const int tx = tile % f.tiles_x, ty = f.ty0 + tile / f.tiles_x;
    const int wx0 = tx * kTileW + (sub & 1) * 8, wy0 = ty * kTileH + (sub >> 1) * 4;
    const int ix = wx0 + (lane & 7), iy = wy0 + (lane >> 3);
    const bool valid = ix < f.W && iy < f.H;
    const float fx = (float)ix, fy = (float)iy;
    const uint32_t tc = a.tile_count[ty * f.tiles_x + tx];
    const uint32_t n = tc & kTileCountMask;
    const bool lean_tile = FDC_LEAN_LOOP && (tc & kTileNeedsFullPath) == 0u && a.stats == nullptr;

    Pixel px;
    {
      uint32_t c = a.clear_rgba8;
      if (a.load_dst && valid) c = fb32[(size_t)iy * f.W + ix];
      // A later segment of the frame (after a backdrop blur) that paints nothing in this tile leaves its pixels alone.
      // (Both loads above are in flight together; peers still need the band's final pixels, so not when the gather is
      // fused into this kernel.)
      if (n == 0u && a.load_dst && a.n_peers == 0 && a.multicast == nullptr) continue;
      px.r = __uint_as_float(kBiasBits | (c & 255u));
      px.g = __uint_as_float(kBiasBits | ((c >> 8) & 255u));
      px.b = __uint_as_float(kBiasBits | ((c >> 16) & 255u));
      px.a = __uint_as_float(kBiasBits | (c >> 24));
      px.mlo = px.mhi = 0;
    }

    const uint2* __restrict__ list = reinterpret_cast<const uint2*>(a.tile_list + a.tile_start[ty * f.tiles_x + tx]);
    // Everything a warp needs to cull is in the 8-byte tile entries (fine_bin_kernel computed it once per tile):
    // bit `sub` of info = "bbox overlaps my block", bit `8+sub` = "my block lies inside the inner rect".
    const uint32_t ov_bit = 1u << (TE_OV_SHIFT + sub), full_bit = 1u << (TE_FULL_SHIFT + sub);

    // Occlusion: the last opaque fill whose inner rect covers this whole block makes everything before it
    // irrelevant for these pixels.  Scan the entries backwards, 32 at a time.
    uint32_t start = 0;
    if (wx0 < f.W && wy0 < f.H) {
      for (int base = (int)n - 1; base >= 0; base -= 32) {
        const int idx = base - lane;
        bool occ = false;
        if (idx >= 0) {
          const uint32_t info = __ldg(&list[idx]).y;
          occ = (info & full_bit) && (info & TE_OCCLUDER);
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, occ);
        if (a.stats && lane == 0) atomicAdd(&a.stats[4], 1ull);
        if (m) { start = (uint32_t)(base - (__ffs(m) - 1)); break; }
      }
    } else {
      start = n;  // block entirely outside the frame
    }

    if (kLeanKernel || lean_tile) walk_list<true>(a, list, n, start, ov_bit, full_bit, queue, lane, ix, iy, fx, fy, px);
    else walk_list<false>(a, list, n, start, ov_bit, full_bit, queue, lane, ix, iy, fx, fy, px);

    {
      const uint32_t out = (__float_as_uint(px.r) & 255u) | ((__float_as_uint(px.g) & 255u) << 8) |
                           ((__float_as_uint(px.b) & 255u) << 16) | ((__float_as_uint(px.a) & 255u) << 24);
      const int tl = (int)(blk >> 3);
      s_out[tl][((sub >> 1) * 4 + (lane >> 3)) * kTileW + (sub & 1) * 8 + (lane & 7)] = out;
      if (lane == 0) atomicOr(&s_dirty, 1u << tl);
    }
  }

  // Copy-out: the CTA's finished tiles leave shared memory as 16-byte chunks, consecutive threads on consecutive
  // addresses of a pixel row across the CTA's tiles (128-bit coalesced stores; up to 512 contiguous bytes per row
  // instead of one 32-byte segment per warp row).  The same chunks are what reaches the other GPUs of a tile-band
  // partition -- through the NVSwitch multicast mapping of the framebuffer (one multimem.st lands in every GPU's
  // copy, the band all-gather costs no extra pass and no extra kernel) or as plain stores into each peer's framebuffer.
  __syncthreads();
  const uint32_t dirty = s_dirty;
  if (dirty == 0u) return;
  const bool vec_ok = (f.W & 3) == 0;
  uint8_t* const mc = a.multicast;
  for (int c = (int)threadIdx.x; c < kTilesPerCta * 64; c += 256) {
    const int q = c & 3, tl = (c >> 2) % kTilesPerCta, r = (c >> 2) / kTilesPerCta;
    if (!((dirty >> tl) & 1u)) continue;
    const int tile = tile0 + tl;
    const int trow = tile / f.tiles_x;
    const int x = (tile - trow * f.tiles_x) * kTileW + q * 4, y = (f.ty0 + trow) * kTileH + r;
    if (y >= f.H || x >= f.W) continue;
    const uint4 v = *reinterpret_cast<const uint4*>(&s_out[tl][r * kTileW + q * 4]);
    const size_t off = ((size_t)y * f.W + x) * 4;
    if (vec_ok) {
      if (mc) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + off), "f"(__uint_as_float(v.x)),
                     "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                     : "memory");
      } else {
        *reinterpret_cast<uint4*>(a.fb + off) = v;
        for (int k = 0; k < a.n_peers; k++) {
          uint8_t* peer = a.peers[k];
          if (peer && peer != a.fb) *reinterpret_cast<uint4*>(peer + off) = v;
        }
      }
    } else {
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      for (int i = 0; i < 4 && x + i < f.W; i++) {
        if (mc) {
          asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(mc + off + 4 * i), "r"(w[i]) : "memory");
        } else {
          *reinterpret_cast<uint32_t*>(a.fb + off + 4 * i) = w[i];
          for (int k = 0; k < a.n_peers; k++) {
            uint8_t* peer = a.peers[k];
            if (peer && peer != a.fb) *reinterpret_cast<uint32_t*>(peer + off + 4 * i) = w[i];
          }
        }
      }
    }
  }
  // Remote stores are ordered for the peers by the release that follows on this stream: the frame's flag barrier
  // (signal_flags_kernel fences at system scope before it stores the flag).  Only a host that brings its own barrier
  // (legacy peer-store path) needs the fence here -- it costs every CTA a round trip before it can retire.
  if (a.fence_at_exit) __threadfence_system();
}

void launch_shade(const ShadeArgs& a, cudaStream_t stream) {
  const int n_tiles = a.frame.tiles_x * (a.frame.ty1 - a.frame.ty0);
  if (n_tiles <= 0) return;
  // Tiles per CTA: 8 keeps neighbouring blocks' records shared in L1, but a band of an 8-GPU partition or a small frame
  // must still give every SM several waves of CTAs (148 SMs x 4 resident CTAs): aim for >= 6 waves.
  const int slots = 148 * FDC_SHADE_MIN_BLOCKS * FDC_SHADE_WAVES;
  if (n_tiles >= 8 * slots) {
    shade_kernel<8, true><<<(n_tiles + 7) / 8, 256, 0, stream>>>(a);
    shade_kernel<8, false><<<(n_tiles + 7) / 8, 256, 0, stream>>>(a);
  } else if (n_tiles >= 4 * slots) {
    shade_kernel<4, true><<<(n_tiles + 3) / 4, 256, 0, stream>>>(a);
    shade_kernel<4, false><<<(n_tiles + 3) / 4, 256, 0, stream>>>(a);
  } else if (n_tiles >= 2 * slots) {
    shade_kernel<2, true><<<(n_tiles + 1) / 2, 256, 0, stream>>>(a);
    shade_kernel<2, false><<<(n_tiles + 1) / 2, 256, 0, stream>>>(a);
  } else {
    shade_kernel<1, true><<<n_tiles, 256, 0, stream>>>(a);
    shade_kernel<1, false><<<n_tiles, 256, 0, stream>>>(a);
  }
}

}  // namespace fdc
