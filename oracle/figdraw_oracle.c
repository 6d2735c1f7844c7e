/*
 * figdraw_oracle.c -- CPU restatement of figdraw's OpenGL render path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * It is the CHECKER for the CUDA backend, never a fallback: nothing under figdraw_b200/ links or calls it.
 *
 * What it restates (paths relative to the reference checkout, v0.35.1):
 *   host half   src/figdraw/opengl/glcontext.nim  transform stack :1991-2017, quad emission :1449-1559 /
 *               :1022-1095 / :1169-1302 / :908-982, radii packing :745-817, mode encode :1002-1008,
 *               mask protocol :1873-1949, rect mask :831-903, backdrop blur :1743-1841, atlas packer :541-586
 *               src/figdraw/figbackend.nim gradientColors/sampleColor :129-183
 *               src/figdraw/opengl/textures.nim mip chain :106-119
 *   device half src/figdraw/opengl/glsl/atlas.frag (all), atlas_rect_mask.frag:222-237, mask.frag:186-234,
 *               blur.frag:1-32; fixed-function blend src/figdraw/utils/glutils.nim:150-154
 *   and the OpenGL 3.3 rules the reference relies on: pixel-centre sampling, two triangles (3,0,1),(2,3,1)
 *   (glcontext.nim:418-429), top-left fill rule, affine varyings, bilinear / trilinear filtering,
 *   float -> UNORM8 round-to-nearest after every blended draw.
 *
 * It is deliberately a different algorithm from the product: immediate mode, one full quad rasterised per
 * call, real full-frame R8 mask textures and full-frame blur passes -- i.e. what GL does -- whereas the
 * product bins primitives into tiles and evaluates masks analytically.
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against the reference's six usable golden PNGs
 * (the PNG files of tests/expected/, copied to tests/golden/).  Arithmetic is float32, compiled with
 * -ffp-contract=off so every GLSL operation rounds once, like the shader as written.
 * Unpinned by any golden: MSDF/MTSDF, elliptical corners, Bezier, rect-mask edge values,
 * backdrop blur, atlas minification; for those this file follows the GLSL text.  The mip chain filter
 * (pixie `minifyBy2`, not vendored) is restated as a premultiplied 2x2 box: parity unpinned.
 *
 * Threading: band parallel.  Every OpenMP thread replays the whole call list on its own rows; blur passes
 * are separated by barriers.  Results are independent of the thread count.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { uint32_t op; uint32_t u[9]; float f[22]; } call_t; /* include/figdraw_cuda.h fdc_call */

enum {
  OP_NOP = 0, OP_SAVE = 1, OP_RESTORE = 2, OP_TRANSLATE = 3, OP_ROTATE = 4, OP_SCALE = 5, OP_APPLY = 6,
  OP_SET_AA = 7, OP_BEGIN_MASK = 8, OP_END_MASK = 9, OP_POP_MASK = 10, OP_BEGIN_RECT_MASK = 11,
  OP_POP_RECT_MASK = 12, OP_BACKDROP_BLUR = 13, OP_SET_SUBPIXEL = 14,
  OP_ROUNDED_RECT = 32, OP_IMAGE = 33, OP_MSDF = 34, OP_BEZIER = 35, OP_FILLED_QUAD = 36, OP_RECT = 37
};
enum { M_ATLAS = 0, M_CLIP_AA = 3, M_DROP = 7, M_DROP_AA = 8, M_INSET = 9, M_ANNULAR = 11, M_ANNULAR_AA = 12,
       M_MSDF = 13, M_MTSDF = 14, M_MSDF_ANN = 15, M_MTSDF_ANN = 16, M_BACKDROP = 17, M_BEZ = 18, M_BEZ_BUTT = 19,
       M_BEZ_SQUARE = 20 };
enum { FILL_COLORS4 = 0, FILL_COLOR = 1, FILL_LIN2 = 2, FILL_LIN3 = 3 };
enum { CAP_AUTO = 0, CAP_ROUND = 1, CAP_BUTT = 2, CAP_SQUARE = 3 };

typedef struct { float x, y; } v2;
typedef struct { float x, y, z, w; } v4;

#define MAX_MASKS 64
#define MAX_LEVELS 16
#define ATLAS_MARGIN 4 /* glcontext.nim:257 */
#define N_MODES 24

/* ------------------------------------------------------------------------------------------------ atlas */
typedef struct { uint64_t key; float x, y, w, h; int used; } entry_t; /* normalised rect, glcontext.nim:583 */

typedef struct oracle {
  int atlas_size, initial_atlas_size, n_levels;
  uint8_t* levels[MAX_LEVELS]; /* RGBA8 straight alpha, level l is (size>>l)^2 */
  uint16_t* heights;
  entry_t* entries; int n_entries, cap_entries;
  int rebuilds;
  int pixelate; /* newContext(pixelate = true): magnification filter GL_NEAREST (glcontext.nim:165-168) */
} oracle;

static void atlas_alloc(oracle* o, int size) {
  o->atlas_size = size;
  o->n_levels = 0;
  for (int s = size; s >= 1 && o->n_levels < MAX_LEVELS; s >>= 1) {
    o->levels[o->n_levels++] = (uint8_t*)calloc((size_t)s * s, 4); /* glGenerateMipmap of an empty texture */
  }
  o->heights = (uint16_t*)calloc((size_t)size, sizeof(uint16_t));
}
static void atlas_free(oracle* o) {
  for (int l = 0; l < o->n_levels; l++) free(o->levels[l]);
  free(o->heights);
  o->n_levels = 0;
}

oracle* orc_create(int atlas_size) {
  oracle* o = (oracle*)calloc(1, sizeof(oracle));
  o->initial_atlas_size = atlas_size;
  atlas_alloc(o, atlas_size);
  return o;
}
void orc_destroy(oracle* o) {
  if (!o) return;
  atlas_free(o);
  free(o->entries);
  free(o);
}
int orc_atlas_size(oracle* o) { return o->atlas_size; }
void orc_set_pixelate(oracle* o, int on) { o->pixelate = on != 0; }
int orc_rebuilds(oracle* o) { return o->rebuilds; }

static entry_t* find_entry(oracle* o, uint64_t key) {
  for (int i = o->n_entries - 1; i >= 0; i--)
    if (o->entries[i].used && o->entries[i].key == key) return &o->entries[i];
  return NULL;
}
int orc_get_image_rect(oracle* o, uint64_t key, float out[4]) {
  entry_t* e = find_entry(o, key);
  if (!e) return 0;
  out[0] = e->x; out[1] = e->y; out[2] = e->w; out[3] = e->h;
  return 1;
}

/* findEmptyRect, glcontext.nim:541-579.  Returns 0 when the atlas had to grow (everything dropped). */
static int find_empty_rect(oracle* o, int width, int height, int* rx, int* ry) {
  for (;;) {
    int imgW = width + ATLAS_MARGIN * 2, imgH = height + ATLAS_MARGIN * 2;
    int lowest = o->atlas_size, at = 0;
    for (int i = 0; i < o->atlas_size; i++) {
      int v = o->heights[i];
      if (v < lowest) {
        int fit = 1;
        for (int j = 0; j <= imgW; j++) {
          if (i + j >= o->atlas_size) { fit = 0; break; }
          if ((int)o->heights[i + j] > v) { fit = 0; break; }
        }
        if (fit) { lowest = v; at = i; }
      }
    }
    if (lowest + imgH > o->atlas_size) {
      /* grow -> resetImageAtlas(2*size), glcontext.nim:536-539, :634-641 */
      int next = o->atlas_size * 2;
      atlas_free(o);
      atlas_alloc(o, next);
      o->n_entries = 0;
      o->rebuilds++;
      continue;
    }
    for (int j = at; j < at + imgW; j++) o->heights[j] = (uint16_t)(lowest + imgH + ATLAS_MARGIN * 2);
    *rx = at + ATLAS_MARGIN;
    *ry = lowest + ATLAS_MARGIN;
    return 1;
  }
}

static void upload_level(oracle* o, int level, int x, int y, int w, int h, const uint8_t* rgba) {
  if (level >= o->n_levels) return;
  int s = o->atlas_size >> level;
  for (int j = 0; j < h; j++) {
    int yy = y + j;
    if (yy < 0 || yy >= s) continue;
    for (int i = 0; i < w; i++) {
      int xx = x + i;
      if (xx < 0 || xx >= s) continue;
      memcpy(o->levels[level] + ((size_t)yy * s + xx) * 4, rgba + ((size_t)j * w + i) * 4, 4);
    }
  }
}

/* updateSubImage, textures.nim:106-119: level chain while w>1 && h>1, offsets halve. minifyBy2 restated as a
 * 2x2 box on premultiplied colour (pixie stores premultiplied RGBX), converted back to straight alpha. */
static void upload_chain(oracle* o, int x, int y, int w, int h, const uint8_t* rgba) {
  uint8_t* cur = (uint8_t*)malloc((size_t)w * h * 4);
  memcpy(cur, rgba, (size_t)w * h * 4);
  int level = 0;
  while (w > 1 && h > 1) {
    upload_level(o, level, x, y, w, h, cur);
    int nw = w / 2, nh = h / 2;
    uint8_t* nxt = (uint8_t*)malloc((size_t)nw * nh * 4);
    for (int j = 0; j < nh; j++)
      for (int i = 0; i < nw; i++) {
        uint32_t acc[4] = {0, 0, 0, 0};
        for (int dj = 0; dj < 2; dj++)
          for (int di = 0; di < 2; di++) {
            const uint8_t* p = cur + ((size_t)(2 * j + dj) * w + (2 * i + di)) * 4;
            uint32_t a = p[3];
            acc[0] += (p[0] * a + 127) / 255;
            acc[1] += (p[1] * a + 127) / 255;
            acc[2] += (p[2] * a + 127) / 255;
            acc[3] += a;
          }
        uint8_t* q = nxt + ((size_t)j * nw + i) * 4;
        uint32_t a = (acc[3] + 2) >> 2;
        for (int c = 0; c < 3; c++) {
          uint32_t pm = (acc[c] + 2) >> 2;
          uint32_t s = a ? (pm * 255 + a / 2) / a : 0;
          q[c] = (uint8_t)(s > 255 ? 255 : s);
        }
        q[3] = (uint8_t)a;
      }
    free(cur);
    cur = nxt;
    w = nw; h = nh;
    x /= 2; y /= 2;
    level++;
  }
  free(cur);
}

/* putImage, glcontext.nim:581-586.  Returns 1 if the atlas was rebuilt by this call. */
int orc_put_image(oracle* o, uint64_t key, int w, int h, const uint8_t* rgba, float out_rect[4]) {
  int before = o->rebuilds, rx, ry;
  find_empty_rect(o, w, h, &rx, &ry);
  if (o->n_entries == o->cap_entries) {
    o->cap_entries = o->cap_entries ? o->cap_entries * 2 : 64;
    o->entries = (entry_t*)realloc(o->entries, sizeof(entry_t) * o->cap_entries);
  }
  entry_t* e = find_entry(o, key);
  if (!e) e = &o->entries[o->n_entries++];
  float as = (float)o->atlas_size;
  e->key = key; e->used = 1;
  e->x = (float)rx / as; e->y = (float)ry / as; e->w = (float)w / as; e->h = (float)h / as;
  if (out_rect) { out_rect[0] = e->x; out_rect[1] = e->y; out_rect[2] = e->w; out_rect[3] = e->h; }
  upload_chain(o, rx, ry, w, h, rgba);
  return o->rebuilds != before;
}

/* ------------------------------------------------------------------------------------------------ GLSL helpers */
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float length2(float x, float y) { return sqrtf(x * x + y * y); }
static inline float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
static inline float median3(float a, float b, float c) { return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c)); }

/* atlas.frag:51-69 */
static float sdRoundedBox(v2 p, v2 b, v4 r) {
  float rr;
  if (p.x > 0.0f) rr = (p.y > 0.0f) ? r.x : r.y;
  else rr = (p.y > 0.0f) ? r.z : r.w;
  float qx = fabsf(p.x) - b.x + rr, qy = fabsf(p.y) - b.y + rr;
  return fminf(fmaxf(qx, qy), 0.0f) + length2(fmaxf(qx, 0.0f), fmaxf(qy, 0.0f)) - rr;
}
/* atlas.frag:71-79 */
static float sdEllipse(v2 p, v2 radii) {
  float sx = fmaxf(radii.x, 0.000001f), sy = fmaxf(radii.y, 0.000001f);
  float k0 = length2(p.x / sx, p.y / sy);
  if (k0 <= 0.000001f) return -fminf(sx, sy);
  float k1 = length2(p.x / (sx * sx), p.y / (sy * sy));
  return k0 * (k0 - 1.0f) / fmaxf(k1, 0.000001f);
}
static float selectCornerRadius(v4 r, v2 p) {
  if (p.x > 0.0f) return (p.y > 0.0f) ? r.x : r.y;
  return (p.y > 0.0f) ? r.z : r.w;
}
/* atlas.frag:96-115 */
static float sdEllipticalRoundedBox(v2 p, v2 b, v4 packed) {
  float sel = selectCornerRadius(packed, p);
  if (sel < 0.0f) {
    float r = -sel - 1.0f;
    v4 rr = {r, r, r, r};
    return sdRoundedBox(p, b, rr);
  }
  float pv = floorf(sel + 0.5f);
  float rx = (pv - 4096.0f * floorf(pv / 4096.0f)) * b.x / 4095.0f; /* mod(pv, 4096) */
  float ry = floorf(pv / 4096.0f) * b.y / 4095.0f;
  if (rx <= 0.0f || ry <= 0.0f) {
    float qx = fabsf(p.x) - b.x, qy = fabsf(p.y) - b.y;
    return fminf(fmaxf(qx, qy), 0.0f) + length2(fmaxf(qx, 0.0f), fmaxf(qy, 0.0f));
  }
  if (rx == ry) {
    v4 rr = {rx, rx, rx, rx};
    return sdRoundedBox(p, b, rr);
  }
  float qx = fabsf(p.x) - b.x + rx, qy = fabsf(p.y) - b.y + ry;
  if (qx > 0.0f && qy > 0.0f) {
    v2 q = {qx, qy}, rad = {rx, ry};
    return sdEllipse(q, rad);
  }
  return fmaxf(qx - rx, qy - ry);
}
/* atlas.frag:121-160 */
static float sdBezier(v2 pos, v2 A, v2 B, v2 C) {
  v2 a = {B.x - A.x, B.y - A.y};
  v2 b = {A.x - 2.0f * B.x + C.x, A.y - 2.0f * B.y + C.y};
  float bb = b.x * b.x + b.y * b.y;
  if (bb <= 0.000001f) {
    v2 ba = {C.x - A.x, C.y - A.y};
    float h = clampf(((pos.x - A.x) * ba.x + (pos.y - A.y) * ba.y) / fmaxf(ba.x * ba.x + ba.y * ba.y, 0.000001f),
                     0.0f, 1.0f);
    return length2(pos.x - (A.x + ba.x * h), pos.y - (A.y + ba.y * h));
  }
  v2 c = {a.x * 2.0f, a.y * 2.0f};
  v2 d = {A.x - pos.x, A.y - pos.y};
  float kk = 1.0f / bb;
  float kx = kk * (a.x * b.x + a.y * b.y);
  float ky = kk * (2.0f * (a.x * a.x + a.y * a.y) + (d.x * b.x + d.y * b.y)) / 3.0f;
  float kz = kk * (d.x * a.x + d.y * a.y);
  float p = ky - kx * kx;
  float p3 = p * p * p;
  float q = kx * (2.0f * kx * kx - 3.0f * ky) + kz;
  float h = q * q + 4.0f * p3;
  float res;
  if (h >= 0.0f) {
    h = sqrtf(h);
    float x0 = (h - q) / 2.0f, x1 = (-h - q) / 2.0f;
    float r0 = signf(x0) * powf(fabsf(x0), 1.0f / 3.0f), r1 = signf(x1) * powf(fabsf(x1), 1.0f / 3.0f);
    float t = clampf(r0 + r1 - kx, 0.0f, 1.0f);
    float ex = d.x + (c.x + b.x * t) * t, ey = d.y + (c.y + b.y * t) * t;
    res = ex * ex + ey * ey;
  } else {
    float z = sqrtf(-p);
    float v = acosf(clampf(q / (p * z * 2.0f), -1.0f, 1.0f)) / 3.0f;
    float m = cosf(v), n = sinf(v) * 1.732050808f;
    float t1 = clampf((m + m) * z - kx, 0.0f, 1.0f), t2 = clampf((-n - m) * z - kx, 0.0f, 1.0f);
    float e1x = d.x + (c.x + b.x * t1) * t1, e1y = d.y + (c.y + b.y * t1) * t1;
    float e2x = d.x + (c.x + b.x * t2) * t2, e2y = d.y + (c.y + b.y * t2) * t2;
    res = fminf(e1x * e1x + e1y * e1y, e2x * e2x + e2y * e2y);
  }
  return sqrtf(res);
}
static v2 safeNormalize(v2 v, v2 fb) {
  float len = length2(v.x, v.y);
  if (len <= 0.000001f) return fb;
  v2 r = {v.x / len, v.y / len};
  return r;
}
/* atlas.frag:178-209 */
static float bezierStrokeSd(float dist, v2 pos, v2 A, v2 B, v2 C, float halfW, int mode) {
  if (mode == M_BEZ) return dist - halfW;
  v2 chord = {C.x - A.x, C.y - A.y}, e10 = {1.0f, 0.0f};
  v2 fb = safeNormalize(chord, e10);
  v2 ba = {B.x - A.x, B.y - A.y}, cb = {C.x - B.x, C.y - B.y};
  v2 startT = safeNormalize(ba, fb), endT = safeNormalize(cb, fb);
  float startProj = (pos.x - A.x) * startT.x + (pos.y - A.y) * startT.y;
  float endProj = (pos.x - C.x) * endT.x + (pos.y - C.y) * endT.y;
  float trim = (mode == M_BEZ_SQUARE) ? halfW : 0.0f;
  float tube = dist;
  if (mode == M_BEZ_SQUARE) {
    if (startProj < 0.0f) tube = fminf(tube, fabsf((pos.x - A.x) * startT.y - (pos.y - A.y) * startT.x));
    if (endProj > 0.0f) tube = fminf(tube, fabsf((pos.x - C.x) * endT.y - (pos.y - C.y) * endT.x));
  }
  float cap = fmaxf(-startProj - trim, endProj - trim);
  return fmaxf(tube - halfW, cap);
}
/* atlas.frag:211-216 */
static float shadowProfile(float sd, float blurRadius) {
  float sigma = fmaxf(0.5f * blurRadius, 0.5f);
  float z = sd / sigma;
  return expf(-0.5f * z * z);
}
/* atlas.frag:218-250 */
static v4 evalFillColor(v4 color, v4 mid, v4 stop, int fillMode, float midPos, v2 uv) {
  if (fillMode == 0) return color;
  float t;
  switch (fillMode) {
    case 1: t = uv.x; break;
    case 2: t = uv.y; break;
    case 3: t = 0.5f * (uv.x + uv.y); break;
    case 4: t = 0.5f * (uv.x + (1.0f - uv.y)); break;
    default: t = 0.0f; break;
  }
  t = clampf(t, 0.0f, 1.0f);
  float m = clampf(midPos, 0.01f, 0.99f);
  v4 r;
  if (t <= m) {
    float k = t / m;
    r.x = mixf(color.x, mid.x, k); r.y = mixf(color.y, mid.y, k); r.z = mixf(color.z, mid.z, k); r.w = mixf(color.w, mid.w, k);
  } else {
    float k = (t - m) / (1.0f - m);
    r.x = mixf(mid.x, stop.x, k); r.y = mixf(mid.y, stop.y, k); r.z = mixf(mid.z, stop.z, k); r.w = mixf(mid.w, stop.w, k);
  }
  return r;
}

/* ------------------------------------------------------------------------------------------------ textures */
static inline int wrap_repeat(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }
static inline int clamp_i(int i, int lo, int hi) { return i < lo ? lo : (i > hi ? hi : i); }

/* GL_LINEAR with GL_REPEAT on one RGBA8 level; uv normalised. */
static v4 tex_bilinear_rgba(const uint8_t* img, int size, float u, float v) {
  float x = u * (float)size - 0.5f, y = v * (float)size - 0.5f;
  float fx = floorf(x), fy = floorf(y);
  float ax = x - fx, ay = y - fy;
  int i0 = wrap_repeat((int)fx, size), i1 = wrap_repeat((int)fx + 1, size);
  int j0 = wrap_repeat((int)fy, size), j1 = wrap_repeat((int)fy + 1, size);
  const uint8_t* t00 = img + ((size_t)j0 * size + i0) * 4;
  const uint8_t* t10 = img + ((size_t)j0 * size + i1) * 4;
  const uint8_t* t01 = img + ((size_t)j1 * size + i0) * 4;
  const uint8_t* t11 = img + ((size_t)j1 * size + i1) * 4;
  float out[4];
  for (int c = 0; c < 4; c++) {
    float a = (float)t00[c] / 255.0f, b = (float)t10[c] / 255.0f, cc = (float)t01[c] / 255.0f, d = (float)t11[c] / 255.0f;
    float top = a * (1.0f - ax) + b * ax, bot = cc * (1.0f - ax) + d * ax;
    out[c] = top * (1.0f - ay) + bot * ay;
  }
  v4 r = {out[0], out[1], out[2], out[3]};
  return r;
}

/* GL_NEAREST with GL_REPEAT on one level: the texel whose cell contains the coordinate. */
static v4 tex_nearest_rgba(const uint8_t* img, int size, float u, float v) {
  int i = wrap_repeat((int)floorf(u * (float)size), size), j = wrap_repeat((int)floorf(v * (float)size), size);
  const uint8_t* t = img + ((size_t)j * size + i) * 4;
  v4 r = {(float)t[0] / 255.0f, (float)t[1] / 255.0f, (float)t[2] / 255.0f, (float)t[3] / 255.0f};
  return r;
}
/* Level 0 through the MAGNIFICATION filter (lambda <= 0, and textureLod(.., 0)): LINEAR, or NEAREST in a `pixelate` context. */
static v4 tex_mag_rgba(const oracle* o, float u, float v) {
  return o->pixelate ? tex_nearest_rgba(o->levels[0], o->atlas_size, u, v) : tex_bilinear_rgba(o->levels[0], o->atlas_size, u, v);
}

/* texture(atlasTex, uv) with implicit LOD: min LINEAR_MIPMAP_LINEAR, mag LINEAR / NEAREST (glcontext.nim:157-169).
 * duv* are the screen-space derivatives of the normalised coordinate. */
static v4 atlas_sample_lod(const oracle* o, float u, float v, float dudx, float dvdx, float dudy, float dvdy) {
  float s = (float)o->atlas_size;
  float rx = length2(dudx * s, dvdx * s), ry = length2(dudy * s, dvdy * s);
  float rho = fmaxf(rx, ry);
  float lambda = (rho > 0.0f) ? log2f(rho) : -1000.0f;
  if (lambda <= 0.0f) return tex_mag_rgba(o, u, v);
  int maxl = o->n_levels - 1;
  if (lambda >= (float)maxl) return tex_bilinear_rgba(o->levels[maxl], o->atlas_size >> maxl, u, v);
  int d1 = (int)floorf(lambda);
  float fr = lambda - (float)d1;
  v4 a = tex_bilinear_rgba(o->levels[d1], o->atlas_size >> d1, u, v);
  v4 b = tex_bilinear_rgba(o->levels[d1 + 1], o->atlas_size >> (d1 + 1), u, v);
  v4 r = {a.x * (1.0f - fr) + b.x * fr, a.y * (1.0f - fr) + b.y * fr, a.z * (1.0f - fr) + b.z * fr,
          a.w * (1.0f - fr) + b.w * fr};
  return r;
}

/* ------------------------------------------------------------------------------------------------ render state */
typedef struct { int fast; v4 params, radii, matX, matY; } rect_mask_t;

typedef struct quad {
  v2 pos[4];   /* BL, BR, TR, TL after ceil (glcontext.nim:1498-1503) */
  v2 uv[4];
  v4 color[4]; /* u8 / 255 */
  v4 mid, stop, params, radii;
  int mode_packed; /* sdfMode + 128*elliptical + 256*fillMode (glcontext.nim:1002-1008) */
  v2 factors;
  float subpixel;
  int has_rm;
  rect_mask_t rm;
  int call_index;
  int is_rounded_rect;
} oquad_t;

typedef struct shared {
  const oracle* o;
  int W, H;
  uint8_t* fb;               /* RGBA8, top-left origin */
  uint8_t* masks[MAX_MASKS]; /* R8, level 0 unused ("white") */
  uint8_t* backdrop;         /* RGBA8 */
  uint8_t* temp;
  int64_t frag_counts[N_MODES][8]; /* reduced at the end */
  int count_only; /* orc_count_fragments: rasterise coverage, skip shading */
  /* reference binning (orc_collect_quads): quads are recorded instead of rasterised */
  int collect;
  int32_t* rec;      /* 8 ints per record: segment, call_index, x0, y0, x1, y1, is_mask, level */
  int64_t rec_cap, rec_n;
  int band_y0, band_y1;
  int segment;
} shared_t;

typedef struct tctx {
  shared_t* sh;
  int y0, y1;      /* rows this thread owns */
  float mat[16];
  float stack[64][16];
  int n_stack;
  float aa;
  int subpixel_enabled;
  float subpixel_shift;
  int mask_write, mask_begun;
  rect_mask_t rm_stack[MAX_MASKS];
  int n_rm;
  int64_t frag_counts[2 * N_MODES]; /* [mode] solid fill, [N_MODES + mode] gradient fill (3-stop or vertex colours) */
  int error;
  /* reference binning: clip rect of every texture-mask level (x0,y0,x1,y1), number of mask quads drawn into it,
   * and the quads themselves so they can be re-emitted after a backdrop blur */
  int clip[MAX_MASKS][4];
  int level_quads[MAX_MASKS];
  int level_ok[MAX_MASKS];
  struct { int call_index; int box[4]; } level_rec[MAX_MASKS][8];
  int64_t level_first_rec[MAX_MASKS];
} tctx;

/* vmath: m[col*4+row]; a*b with each entry summed left to right. */
static void mat_identity(float* m) { memset(m, 0, 64); m[0] = m[5] = m[10] = m[15] = 1.0f; }
static void mat_mul(float* out, const float* a, const float* b) {
  float r[16];
  for (int c = 0; c < 4; c++)
    for (int row = 0; row < 4; row++)
      r[c * 4 + row] = a[0 * 4 + row] * b[c * 4 + 0] + a[1 * 4 + row] * b[c * 4 + 1] + a[2 * 4 + row] * b[c * 4 + 2] +
                       a[3 * 4 + row] * b[c * 4 + 3];
  memcpy(out, r, 64);
}
/* `ctx.mat * vec2` (glcontext.nim:905-906): (m * vec3(x, y, 0)).xy */
static v2 mat_apply(const float* m, float x, float y) {
  v2 r;
  r.x = m[0] * x + m[4] * y + m[8] * 0.0f + m[12];
  r.y = m[1] * x + m[5] * y + m[9] * 0.0f + m[13];
  return r;
}
/* general 4x4 inverse (cofactor expansion), for makeRectMask glcontext.nim:837 */
static int mat_inverse(const float* m, float* inv) {
  float t[16];
  t[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  t[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  t[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  t[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  t[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  t[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  t[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  t[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  t[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  t[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  t[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  t[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  t[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  t[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  t[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  t[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  float det = m[0] * t[0] + m[1] * t[4] + m[2] * t[8] + m[3] * t[12];
  if (det == 0.0f) return 0;
  float id = 1.0f / det;
  for (int i = 0; i < 16; i++) inv[i] = t[i] * id;
  return 1;
}

static v4 unpack_color(uint32_t c) {
  v4 r = {(float)(c & 255) / 255.0f, (float)((c >> 8) & 255) / 255.0f, (float)((c >> 16) & 255) / 255.0f,
          (float)((c >> 24) & 255) / 255.0f};
  return r;
}

/* ------------------------------------------------------------------------------------------------ fragment */
static inline uint8_t quant8(float x) { /* float -> UNORM8, round to nearest */
  x = clampf(x, 0.0f, 1.0f);
  return (uint8_t)floorf(x * 255.0f + 0.5f);
}

typedef struct frag_in {
  float px, py; /* `pos` varying == pixel centre in top-left coordinates */
  v2 uv;
  v4 color;
  v2 duvdx, duvdy;
} frag_in;

/* atlas_rect_mask.frag:222-237 */
static float rectMaskAlpha(const oquad_t* q, float aa, float px, float py) {
  const rect_mask_t* rm = &q->rm;
  if (rm->params.z < 0.0f || rm->params.w < 0.0f) return 1.0f;
  float lx = (rm->matX.x * px + rm->matX.y * py) + rm->matX.z;
  float ly = (rm->matY.x * px + rm->matY.y * py) + rm->matY.z;
  v2 qq = {lx - rm->params.x, -(ly - rm->params.y)};
  v2 he = {rm->params.z, rm->params.w};
  float dist = rm->matY.w > 0.5f ? sdEllipticalRoundedBox(qq, he, rm->radii) : sdRoundedBox(qq, he, rm->radii);
  return 1.0f - clampf(aa * dist + 0.5f, 0.0f, 1.0f);
}

/* atlas.frag:252-405 (+ atlas_rect_mask.frag:425).  Returns straight-alpha fragColor. */
static v4 main_frag(const tctx* t, const oquad_t* q, const frag_in* in, int mask_read) {
  const shared_t* sh = t->sh;
  int packed = q->mode_packed;
  int fillMode = packed / 256;
  int mode = packed - fillMode * 256;
  int elliptical = mode >= 128;
  if (elliptical) mode -= 128;
  v2 qh = {q->params.x, q->params.y};
  int inset = (mode == M_INSET);
  v2 sh_he = inset ? qh : (v2){q->params.z, q->params.w};
  v2 p = {(in->uv.x - 0.5f) * 2.0f * qh.x, (in->uv.y - 0.5f) * 2.0f * qh.y};
  int bez = (mode == M_BEZ || mode == M_BEZ_BUTT || mode == M_BEZ_SQUARE);
  v2 pf = {p.x, -p.y};
  float dist;
  v2 A = {q->params.z, q->params.w}, B = {q->radii.x, q->radii.y}, C = {q->radii.z, q->radii.w};
  if (bez) dist = sdBezier(p, A, B, C);
  else if (elliptical) dist = sdEllipticalRoundedBox(pf, sh_he, q->radii);
  else dist = sdRoundedBox(pf, sh_he, q->radii);

  float sdfFactor = q->factors.x;
  float sdfSpread = (fillMode == 0) ? q->factors.y : 0.0f;
  v4 fillColor = evalFillColor(in->color, q->mid, q->stop, fillMode, q->factors.y, in->uv);
  float aa = t->aa;
  float alpha = 0.0f;
  v4 frag;
  if (mode == M_ATLAS) {
    float au = in->uv.x;
    if (t->subpixel_enabled) au -= q->subpixel * (1.0f / fmaxf((float)sh->o->atlas_size, 1.0f));
    v4 tex = atlas_sample_lod(sh->o, au, in->uv.y, in->duvdx.x, in->duvdx.y, in->duvdy.x, in->duvdy.y);
    frag.x = tex.x * in->color.x; frag.y = tex.y * in->color.y; frag.z = tex.z * in->color.z; frag.w = tex.w * in->color.w;
  } else if (mode == M_MSDF || mode == M_MTSDF || mode == M_MSDF_ANN || mode == M_MTSDF_ANN) {
    float pxRange = q->factors.x, thr = q->factors.y;
    v4 tex = tex_mag_rgba(sh->o, in->uv.x, in->uv.y); /* textureLod(.., 0): lambda = 0 selects the magnification filter */
    int isMtsdf = (mode == M_MTSDF || mode == M_MTSDF_ANN), isStroke = (mode == M_MSDF_ANN || mode == M_MTSDF_ANN);
    float sd = isMtsdf ? tex.w : median3(tex.x, tex.y, tex.z);
    /* msdfScreenPxRange, atlas.frag:45-49; fwidth = |dFdx| + |dFdy| */
    float ts = (float)sh->o->atlas_size;
    float ux = pxRange / ts, uy = pxRange / ts;
    float fwx = fabsf(in->duvdx.x) + fabsf(in->duvdy.x), fwy = fabsf(in->duvdx.y) + fabsf(in->duvdy.y);
    float spr = fmaxf(0.5f * (ux * (1.0f / fwx) + uy * (1.0f / fwy)), 1.0f);
    float spd = spr * (sd - thr);
    if (isStroke) {
      float halfW = fmaxf(q->params.y, 0.0f) * 0.5f;
      alpha = clampf(halfW - fabsf(spd) + 0.5f, 0.0f, 1.0f);
    } else {
      alpha = clampf(spd + 0.5f, 0.0f, 1.0f);
    }
    frag.x = fillColor.x; frag.y = fillColor.y; frag.z = fillColor.z; frag.w = fillColor.w * alpha;
  } else {
    int is_backdrop = 0;
    switch (mode) {
      case M_BEZ: case M_BEZ_BUTT: case M_BEZ_SQUARE: {
        float sd = bezierStrokeSd(dist, p, A, B, C, fmaxf(sdfFactor, 0.0f) * 0.5f, mode);
        alpha = 1.0f - clampf(aa * sd + 0.5f, 0.0f, 1.0f);
        break;
      }
      case M_ANNULAR: {
        float f = sdfFactor * 0.5f;
        float sd = fabsf(dist + f) - f;
        alpha = (sd < 0.0f) ? 1.0f : 0.0f;
        break;
      }
      case M_ANNULAR_AA: {
        float f = sdfFactor * 0.5f;
        float sd = fabsf(dist + f) - f;
        alpha = 1.0f - clampf(aa * sd + 0.5f, 0.0f, 1.0f);
        break;
      }
      case M_DROP: {
        float sd = dist - sdfSpread;
        float a = shadowProfile(sd, sdfFactor);
        alpha = (sd > 0.0f) ? fminf(a, 1.0f) : 1.0f;
        break;
      }
      case M_DROP_AA: {
        float insideAlpha = 1.0f - clampf(aa * dist + 0.5f, 0.0f, 1.0f);
        float sd = dist - sdfSpread;
        float a = shadowProfile(sd, sdfFactor);
        alpha = (sd >= 0.0f) ? fminf(a, 1.0f) : insideAlpha;
        break;
      }
      case M_INSET: {
        v2 qClip = pf;
        v2 qShadow = {qClip.x - q->params.z, qClip.y - (-q->params.w)};
        float clipDist = elliptical ? sdEllipticalRoundedBox(qClip, qh, q->radii) : sdRoundedBox(qClip, qh, q->radii);
        float clipAlpha = 1.0f - clampf(aa * clipDist + 0.5f, 0.0f, 1.0f);
        float shadowDist = elliptical ? sdEllipticalRoundedBox(qShadow, qh, q->radii) : sdRoundedBox(qShadow, qh, q->radii);
        float sd = shadowDist + sdfSpread;
        float a = shadowProfile(sd, sdfFactor);
        float insetAlpha = (sd < 0.0f) ? fminf(a, 1.0f) : 1.0f;
        alpha = clipAlpha * insetAlpha;
        break;
      }
      case M_BACKDROP: {
        alpha = 1.0f - clampf(aa * dist + 0.5f, 0.0f, 1.0f);
        /* texture(backdropTex, (x/W, 1-y/H)) at an exact texel centre == fetch of this pixel */
        int ix = clamp_i((int)floorf(in->px), 0, sh->W - 1), iy = clamp_i((int)floorf(in->py), 0, sh->H - 1);
        const uint8_t* b = sh->backdrop + ((size_t)iy * sh->W + ix) * 4;
        frag.x = (float)b[0] / 255.0f; frag.y = (float)b[1] / 255.0f; frag.z = (float)b[2] / 255.0f;
        frag.w = ((float)b[3] / 255.0f) * alpha;
        is_backdrop = 1;
        break;
      }
      default: {
        alpha = 1.0f - clampf(aa * dist + 0.5f, 0.0f, 1.0f);
        break;
      }
    }
    if (!is_backdrop) { frag.x = fillColor.x; frag.y = fillColor.y; frag.z = fillColor.z; frag.w = fillColor.w * alpha; }
  }
  if (mask_read != 0) {
    int ix = clamp_i((int)floorf(in->px), 0, sh->W - 1), iy = clamp_i((int)floorf(in->py), 0, sh->H - 1);
    frag.w *= (float)sh->masks[mask_read][(size_t)iy * sh->W + ix] / 255.0f;
  }
  if (q->has_rm) frag.w *= rectMaskAlpha(q, aa, in->px, in->py);
  return frag;
}

/* mask.frag:186-234.  Returns alpha. */
static float mask_frag(const tctx* t, const oquad_t* q, const frag_in* in, int mask_read) {
  const shared_t* sh = t->sh;
  int packed = q->mode_packed;
  int fillMode = packed / 256;
  int mode = packed - fillMode * 256;
  int elliptical = mode >= 128;
  if (elliptical) mode -= 128;
  float alpha;
  if (mode == M_ATLAS) {
    v4 tex = atlas_sample_lod(sh->o, in->uv.x, in->uv.y, in->duvdx.x, in->duvdx.y, in->duvdy.x, in->duvdy.y);
    alpha = tex.w * in->color.w;
  } else {
    v2 qh = {q->params.x, q->params.y}, she = {q->params.z, q->params.w};
    v2 p = {(in->uv.x - 0.5f) * 2.0f * qh.x, (in->uv.y - 0.5f) * 2.0f * qh.y};
    v2 pf = {p.x, -p.y};
    float dist;
    if (mode == M_BEZ || mode == M_BEZ_BUTT || mode == M_BEZ_SQUARE) {
      v2 A = {q->params.z, q->params.w}, B = {q->radii.x, q->radii.y}, C = {q->radii.z, q->radii.w};
      float bd = sdBezier(p, A, B, C);
      dist = bezierStrokeSd(bd, p, A, B, C, fmaxf(q->factors.x, 0.0f) * 0.5f, mode);
    } else if (elliptical) dist = sdEllipticalRoundedBox(pf, she, q->radii);
    else dist = sdRoundedBox(pf, she, q->radii);
    if (mode == M_ANNULAR_AA) {
      float hw = fmaxf(q->factors.x, 0.0f) * 0.5f;
      dist = fabsf(dist + hw) - hw;
    }
    alpha = (1.0f - clampf(t->aa * dist + 0.5f, 0.0f, 1.0f)) * in->color.w;
  }
  if (mask_read != 0) {
    int ix = clamp_i((int)floorf(in->px), 0, sh->W - 1), iy = clamp_i((int)floorf(in->py), 0, sh->H - 1);
    alpha *= (float)sh->masks[mask_read][(size_t)iy * sh->W + ix] / 255.0f;
  }
  return alpha;
}

/* ------------------------------------------------------------------------------------------------ rasteriser */
/* One triangle (A,B,C are indices into q->pos).  Vertices are integers after ceil, so edge functions at pixel
 * centres are exact in int64 when everything is doubled.  Top-left rule on ties. */
static void raster_tri(tctx* t, const oquad_t* q, int ia, int ib, int ic, int mask_read) {
  shared_t* sh = t->sh;
  const v2 *A = &q->pos[ia], *B = &q->pos[ib], *C = &q->pos[ic];
  /* doubled integer coordinates */
  int64_t ax = (int64_t)llrintf(A->x * 2.0f), ay = (int64_t)llrintf(A->y * 2.0f);
  int64_t bx = (int64_t)llrintf(B->x * 2.0f), by = (int64_t)llrintf(B->y * 2.0f);
  int64_t cx = (int64_t)llrintf(C->x * 2.0f), cy = (int64_t)llrintf(C->y * 2.0f);
  int64_t area = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
  if (area == 0) return;
  int64_t sgn = area > 0 ? 1 : -1;
  float minx = fminf(A->x, fminf(B->x, C->x)), maxx = fmaxf(A->x, fmaxf(B->x, C->x));
  float miny = fminf(A->y, fminf(B->y, C->y)), maxy = fmaxf(A->y, fmaxf(B->y, C->y));
  int x0 = clamp_i((int)floorf(minx), 0, sh->W), x1 = clamp_i((int)ceilf(maxx), 0, sh->W);
  int y0 = clamp_i((int)floorf(miny), t->y0, t->y1), y1 = clamp_i((int)ceilf(maxy), t->y0, t->y1);
  if (x0 >= x1 || y0 >= y1) return;
  /* edge e(p) = sgn * ((Q-P) x (p-P)); inside when all three >= 0 (with tie rule).  a = d/dx, b = d/dy. */
  const int64_t ex[3] = {ax, bx, cx}, ey[3] = {ay, by, cy};
  int64_t ea[3], eb[3];
  for (int k = 0; k < 3; k++) {
    int n = (k + 1) % 3;
    int64_t dx = ex[n] - ex[k], dy = ey[n] - ey[k];
    ea[k] = -sgn * dy;
    eb[k] = sgn * dx;
  }
  /* affine varyings: attr(p) = aA + (p - A) . grad; gradients in float32 */
  float farea = (B->x - A->x) * (C->y - A->y) - (B->y - A->y) * (C->x - A->x);
  float inv_area = 1.0f / farea;
  float bcy = (C->y - A->y) * inv_area, bby = (B->y - A->y) * inv_area;
  float bcx = (C->x - A->x) * inv_area, bbx = (B->x - A->x) * inv_area;
#define GRADX(fa, fb, fc) (((fb) - (fa)) * bcy - ((fc) - (fa)) * bby)
#define GRADY(fa, fb, fc) (((fc) - (fa)) * bbx - ((fb) - (fa)) * bcx)
  v2 duvdx = {GRADX(q->uv[ia].x, q->uv[ib].x, q->uv[ic].x), GRADX(q->uv[ia].y, q->uv[ib].y, q->uv[ic].y)};
  v2 duvdy = {GRADY(q->uv[ia].x, q->uv[ib].x, q->uv[ic].x), GRADY(q->uv[ia].y, q->uv[ib].y, q->uv[ic].y)};
  float cdx[4], cdy[4];
  const float* ca = &q->color[ia].x; const float* cb = &q->color[ib].x; const float* cc = &q->color[ic].x;
  for (int k = 0; k < 4; k++) { cdx[k] = GRADX(ca[k], cb[k], cc[k]); cdy[k] = GRADY(ca[k], cb[k], cc[k]); }
  int packed = q->mode_packed;
  int mode = (packed % 256) % 128;
  int to_mask = t->mask_begun;
  int64_t nfrag = 0;
  for (int y = y0; y < y1; y++) {
    int64_t py2 = 2 * (int64_t)y + 1;
    for (int x = x0; x < x1; x++) {
      int64_t px2 = 2 * (int64_t)x + 1;
      int inside = 1;
      for (int k = 0; k < 3; k++) {
        int64_t e = ea[k] * (px2 - ex[k]) + eb[k] * (py2 - ey[k]);
        if (e < 0) { inside = 0; break; }
        if (e == 0 && !(ea[k] > 0 || (ea[k] == 0 && eb[k] > 0))) { inside = 0; break; }
      }
      if (!inside) continue;
      nfrag++;
      if (sh->count_only) continue;
      frag_in in;
      in.px = (float)x + 0.5f; in.py = (float)y + 0.5f;
      float rx = in.px - A->x, ry = in.py - A->y;
      in.uv.x = q->uv[ia].x + (rx * duvdx.x + ry * duvdy.x);
      in.uv.y = q->uv[ia].y + (rx * duvdx.y + ry * duvdy.y);
      in.color.x = ca[0] + (rx * cdx[0] + ry * cdy[0]);
      in.color.y = ca[1] + (rx * cdx[1] + ry * cdy[1]);
      in.color.z = ca[2] + (rx * cdx[2] + ry * cdy[2]);
      in.color.w = ca[3] + (rx * cdx[3] + ry * cdy[3]);
      in.duvdx = duvdx; in.duvdy = duvdy;
      if (to_mask) {
        /* blend stays enabled on the R8 target: r = a*a + dst*(1-a)  (glutils.nim:150-154, mask.frag:233) */
        float a = mask_frag(t, q, &in, mask_read);
        uint8_t* m = &sh->masks[t->mask_write][(size_t)y * sh->W + x];
        float d = (float)*m / 255.0f;
        *m = quant8(a * a + d * (1.0f - a));
      } else {
        v4 s = main_frag(t, q, &in, mask_read);
        uint8_t* d8 = sh->fb + ((size_t)y * sh->W + x) * 4;
        float dr = (float)d8[0] / 255.0f, dg = (float)d8[1] / 255.0f, db = (float)d8[2] / 255.0f, da = (float)d8[3] / 255.0f;
        float ia_ = 1.0f - s.w;
        d8[0] = quant8(s.x * s.w + dr * ia_);
        d8[1] = quant8(s.y * s.w + dg * ia_);
        d8[2] = quant8(s.z * s.w + db * ia_);
        d8[3] = quant8(s.w + da * ia_); /* glBlendFuncSeparate(.., GL_ONE, GL_ONE_MINUS_SRC_ALPHA) */
      }
    }
  }
  {
    int grad = (packed / 256) != 0 || memcmp(&q->color[0], &q->color[1], sizeof(v4)) != 0 ||
               memcmp(&q->color[0], &q->color[2], sizeof(v4)) != 0 || memcmp(&q->color[0], &q->color[3], sizeof(v4)) != 0;
    if (mode < N_MODES) t->frag_counts[mode + (grad ? N_MODES : 0)] += nfrag;
  }
#undef GRADX
#undef GRADY
}


/* ------------------------------------------------------------------------------------------------ reference binning
 * The binning RULE of the product (DESIGN.md "Binning rule"), restated the plainest way: a quad's bin box is the
 * bbox of its four ceil'd corners, intersected with the frame, the rank's row band and the clip box of the
 * texture-mask level it lives under (content) or of the parent level (mask quads).  A level's clip box is the bin
 * box of its single mask quad; a level with several mask quads does not clip; a level whose mask quad was dropped
 * clips everything.  Segments end at each backdrop blur; open mask levels are re-emitted at the start of the next. */
static int sat_i(float v) { v = fminf(fmaxf(v, -30000.0f), 30000.0f); return (int)v; }
static void emit_record_seg(tctx* t, int segment, int call_index, const int box[4], int is_mask, int level) {
  shared_t* sh = t->sh;
  if (sh->rec_n < sh->rec_cap) {
    int32_t* r = sh->rec + sh->rec_n * 8;
    r[0] = segment; r[1] = call_index; r[2] = box[0]; r[3] = box[1]; r[4] = box[2]; r[5] = box[3]; r[6] = is_mask; r[7] = level;
  }
  sh->rec_n++;
}
static void emit_record(tctx* t, int call_index, const int box[4], int is_mask, int level) {
  emit_record_seg(t, t->sh->segment, call_index, box, is_mask, level);
}
static int box_empty(const int* b) { return b[0] >= b[2] || b[1] >= b[3]; }
/* Binning rule.  Content is binned into bbox(ceil'd corners) /\ frame /\ band /\ clip box of the mask
 * level it is drawn under.  A level's clip box is the bin box of its single rounded-rect mask quad; a level with several
 * mask quads (or a first quad that is not a rounded rect) does not clip, and because GL cleared the WHOLE mask texture
 * at beginMask (glcontext.nim:1901-1902) its first quad is binned over the parent's whole clip box instead of its own
 * bbox (it carries the clear).  That is only known when the second quad arrives: the first quad's record is then
 * widened in place; a first quad with an empty bin box keeps a placeholder record (segment -1) for that purpose. */
static void collect_quad(tctx* t, const oquad_t* q) {
  shared_t* sh = t->sh;
  int box[4];
  box[0] = sat_i(fminf(fminf(q->pos[0].x, q->pos[1].x), fminf(q->pos[2].x, q->pos[3].x)));
  box[2] = sat_i(fmaxf(fmaxf(q->pos[0].x, q->pos[1].x), fmaxf(q->pos[2].x, q->pos[3].x)));
  box[1] = sat_i(fminf(fminf(q->pos[0].y, q->pos[1].y), fminf(q->pos[2].y, q->pos[3].y)));
  box[3] = sat_i(fmaxf(fmaxf(q->pos[0].y, q->pos[1].y), fmaxf(q->pos[2].y, q->pos[3].y)));
  const int L = t->mask_write;
  const int* clip = t->mask_begun ? t->clip[L - 1] : t->clip[L];
  int wide[4] = {clip[0], clip[1], clip[2], clip[3]};
  if (wide[0] < 0) wide[0] = 0;
  if (wide[2] > sh->W) wide[2] = sh->W;
  if (wide[1] < sh->band_y0) wide[1] = sh->band_y0;
  if (wide[3] > sh->band_y1) wide[3] = sh->band_y1;
  if (box[0] < wide[0]) box[0] = wide[0];
  if (box[1] < wide[1]) box[1] = wide[1];
  if (box[2] > wide[2]) box[2] = wide[2];
  if (box[3] > wide[3]) box[3] = wide[3];
  int empty = box_empty(box);
  if (t->mask_begun) {
    int k = t->level_quads[L]++;
    int is_rect = q->is_rounded_rect;
    if (k == 0) {
      t->level_first_rec[L] = sh->rec_n;
      if (is_rect) {
        t->level_ok[L] = 1;
        if (empty) { t->clip[L][0] = t->clip[L][1] = t->clip[L][2] = t->clip[L][3] = 0; }
        else memcpy(t->clip[L], box, sizeof(box));
      } else {
        t->level_ok[L] = 0;
        memcpy(t->clip[L], t->clip[L - 1], sizeof(box));
        memcpy(box, wide, sizeof(box));
        empty = box_empty(box);
      }
      t->level_rec[L][0].call_index = q->call_index;
      memcpy(t->level_rec[L][0].box, box, sizeof(box));
      emit_record_seg(t, empty ? -1 : sh->segment, q->call_index, box, 1, L);
      return;
    }
    if (k == 1 && t->level_ok[L]) {
      /* the level now holds several quads: the first one carries the clear of the whole parent clip box */
      const int64_t at = t->level_first_rec[L];
      if (at < sh->rec_cap) {
        int32_t* r = sh->rec + at * 8;
        r[0] = box_empty(wide) ? -1 : sh->segment;
        r[2] = wide[0]; r[3] = wide[1]; r[4] = wide[2]; r[5] = wide[3];
      }
      memcpy(t->level_rec[L][0].box, wide, sizeof(wide));
    }
    t->level_ok[L] = 0;
    memcpy(t->clip[L], t->clip[L - 1], sizeof(box));
    if (k < 8) { t->level_rec[L][k].call_index = q->call_index; memcpy(t->level_rec[L][k].box, box, sizeof(box)); }
  }
  if (!empty) emit_record(t, q->call_index, box, t->mask_begun, L);
}
static void collect_begin_segment(tctx* t) {
  t->sh->segment++;
  for (int L = 1; L <= t->mask_write; L++)
    for (int k = 0; k < t->level_quads[L] && k < 8; k++) {
      const int* b = t->level_rec[L][k].box;
      if (!(b[0] >= b[2] || b[1] >= b[3])) emit_record(t, t->level_rec[L][k].call_index, b, 1, L);
    }
}

static void draw_quad(tctx* t, oquad_t* q) {
  int mask_read;
  if (t->mask_begun) {
    mask_read = t->mask_write - 1; /* flush(maskTextureWrite - 1), glcontext.nim:720, :1920 */
  } else {
    mask_read = t->mask_write;
    /* setRectMaskVert4, glcontext.nim:864-899: topmost fast rect mask */
    q->has_rm = 0;
    for (int i = t->n_rm - 1; i >= 0; i--)
      if (t->rm_stack[i].fast) { q->has_rm = 1; q->rm = t->rm_stack[i]; break; }
  }
  q->subpixel = t->subpixel_enabled ? fmaxf(0.0f, fminf(t->subpixel_shift, 0.999f)) : 0.0f;
  if (t->sh->collect) { collect_quad(t, q); return; }
  if (t->sh->count_only && q->pos[0].x == q->pos[3].x && q->pos[1].x == q->pos[2].x && q->pos[0].y == q->pos[1].y &&
      q->pos[2].y == q->pos[3].y) {
    /* axis-aligned: the two triangles cover exactly [X0,X1) x [Y0,Y1) */
    float xa = fminf(q->pos[3].x, q->pos[1].x), xb = fmaxf(q->pos[3].x, q->pos[1].x);
    float ya = fminf(q->pos[3].y, q->pos[1].y), yb = fmaxf(q->pos[3].y, q->pos[1].y);
    int x0 = clamp_i((int)fminf(fmaxf(xa, -1e6f), 1e6f), 0, t->sh->W), x1 = clamp_i((int)fminf(fmaxf(xb, -1e6f), 1e6f), 0, t->sh->W);
    int y0 = clamp_i((int)fminf(fmaxf(ya, -1e6f), 1e6f), t->y0, t->y1), y1 = clamp_i((int)fminf(fmaxf(yb, -1e6f), 1e6f), t->y0, t->y1);
    int mode = (q->mode_packed % 256) % 128;
    int grad = (q->mode_packed / 256) != 0 || memcmp(&q->color[0], &q->color[1], sizeof(v4)) != 0 ||
               memcmp(&q->color[0], &q->color[2], sizeof(v4)) != 0 || memcmp(&q->color[0], &q->color[3], sizeof(v4)) != 0;
    if (x1 > x0 && y1 > y0 && mode < N_MODES) t->frag_counts[mode + (grad ? N_MODES : 0)] += (int64_t)(x1 - x0) * (y1 - y0);
    return;
  }
  raster_tri(t, q, 3, 0, 1, mask_read); /* indices glcontext.nim:418-429 */
  raster_tri(t, q, 2, 3, 1, mask_read);
}

/* ------------------------------------------------------------------------------------------------ host half */
static float nim_round(float x) { return roundf(x); } /* half away from zero */

/* clampRadius, glcontext.nim:745-749 */
static float clampRadius(float r, float maxr) {
  if (r <= 0.0f) return 0.0f;
  return nim_round(fmaxf(1.0f, fminf(r, maxr)));
}
/* roundedRadiiVec, glcontext.nim:751-817.  rx/ry in DirectionCorners order TL,TR,BL,BR. */
static v4 roundedRadiiVec(const float* rx, const float* ry, v2 he, int* elliptical) {
  enum { TL = 0, TR = 1, BL = 2, BR = 3 };
  int circ = 1;
  for (int i = 0; i < 4; i++) if (rx[i] != ry[i]) circ = 0;
  if (circ) {
    float mr = fminf(he.x, he.y);
    v4 r = {clampRadius(rx[TR], mr), clampRadius(rx[BR], mr), clampRadius(rx[TL], mr), clampRadius(rx[BL], mr)};
    *elliptical = 0;
    return r;
  }
  float cmr = fminf(he.x, he.y);
  float out[4];
  const int order[4] = {TR, BR, TL, BL};
  for (int k = 0; k < 4; k++) {
    int c = order[k];
    float cx = clampRadius(rx[c], he.x), cy = clampRadius(ry[c], he.y);
    if (rx[c] == ry[c]) out[k] = -(clampRadius(rx[c], cmr) + 1.0f);
    else if (cx == cy) out[k] = -(cx + 1.0f);
    else {
      float qx = nim_round(clampf(cx / fmaxf(he.x, 0.000001f), 0.0f, 1.0f) * 4095.0f);
      float qy = nim_round(clampf(cy / fmaxf(he.y, 0.000001f), 0.0f, 1.0f) * 4095.0f);
      out[k] = qx + qy * 4096.0f;
    }
  }
  *elliptical = 1;
  v4 r = {out[0], out[1], out[2], out[3]};
  return r;
}

/* lerpColor / sampleColor / gradientColors, figbackend.nim:129-183 */
static uint32_t lerpColor(uint32_t a, uint32_t b, float tt) {
  float ct = clampf(tt, 0.0f, 1.0f), inv = 1.0f - ct;
  uint32_t r = 0;
  for (int k = 0; k < 4; k++) {
    float av = (float)((a >> (8 * k)) & 255), bv = (float)((b >> (8 * k)) & 255);
    r |= ((uint32_t)(uint8_t)nim_round(av * inv + bv * ct)) << (8 * k);
  }
  return r;
}
static uint32_t sampleColor(int kind, const uint32_t* c, float midPos, float tt) {
  if (kind == FILL_COLOR) return c[0];
  if (kind == FILL_LIN2) return lerpColor(c[0], c[1], tt);
  float ct = clampf(tt, 0.0f, 1.0f);
  if (ct <= midPos) return lerpColor(c[0], c[1], ct / midPos);
  return lerpColor(c[1], c[2], (ct - midPos) / (1.0f - midPos));
}
static void gradientColors(int kind, int axis, const uint32_t* c, float midPos, uint32_t out[4]) {
  static const float ts[4][4] = {{0.0f, 1.0f, 1.0f, 0.0f}, {1.0f, 1.0f, 0.0f, 0.0f}, {0.5f, 1.0f, 0.5f, 0.0f}, {0.0f, 0.5f, 1.0f, 0.5f}};
  if (kind == FILL_COLORS4) { memcpy(out, c, 16); return; }
  if (kind == FILL_COLOR) axis = 0;
  for (int k = 0; k < 4; k++) out[k] = sampleColor(kind, c, midPos, ts[axis & 3][k]);
}

static void quad_positions(const tctx* t, oquad_t* q, float atx, float aty, float tox, float toy) {
  v2 p0 = mat_apply(t->mat, atx, toy), p1 = mat_apply(t->mat, tox, toy), p2 = mat_apply(t->mat, tox, aty),
     p3 = mat_apply(t->mat, atx, aty);
  q->pos[0].x = ceilf(p0.x); q->pos[0].y = ceilf(p0.y);
  q->pos[1].x = ceilf(p1.x); q->pos[1].y = ceilf(p1.y);
  q->pos[2].x = ceilf(p2.x); q->pos[2].y = ceilf(p2.y);
  q->pos[3].x = ceilf(p3.x); q->pos[3].y = ceilf(p3.y);
}
static void quad_uvs(oquad_t* q, float uax, float uay, float utx, float uty) {
  q->uv[0].x = uax; q->uv[0].y = uty;
  q->uv[1].x = utx; q->uv[1].y = uty;
  q->uv[2].x = utx; q->uv[2].y = uay;
  q->uv[3].x = uax; q->uv[3].y = uay;
}

/* drawRoundedRectSdf (all overloads) + drawRoundedRectSdfOpenGl, glcontext.nim:1449-1617 */
static void op_rounded_rect(tctx* t, const float* rect, const float* rx, const float* ry, int mode, float factor,
                            float spread, float ssx, float ssy, int fkind, int axis, const uint32_t* fc, float midPos,
                            int call_index) {
  if (rect[2] <= 0.0f || rect[3] <= 0.0f) return;
  oquad_t q;
  memset(&q, 0, sizeof(q));
  q.call_index = call_index;
  q.is_rounded_rect = 1;
  int fillMode = 0;
  uint32_t cols[4], mid = 0, stop = 0;
  float fMid = 0.5f;
  if (fkind == FILL_LIN3 && (mode == M_CLIP_AA || mode == M_ANNULAR || mode == M_ANNULAR_AA)) {
    fillMode = 1 + (axis & 3);
    cols[0] = cols[1] = cols[2] = cols[3] = fc[0];
    mid = fc[1]; stop = fc[2]; fMid = midPos;
  } else {
    gradientColors(fkind, axis, fc, midPos, cols);
  }
  v2 qh = {rect[2] * 0.5f, rect[3] * 0.5f};
  int inset = (mode == M_INSET);
  v2 rs = (ssx > 0.0f && ssy > 0.0f) ? (v2){ssx, ssy} : (v2){rect[2], rect[3]};
  v2 she = inset ? qh : (v2){rs.x * 0.5f, rs.y * 0.5f};
  if (inset) q.params = (v4){qh.x, qh.y, ssx, ssy};
  else q.params = (v4){qh.x, qh.y, she.x, she.y};
  int ell = 0;
  q.radii = roundedRadiiVec(rx, ry, she, &ell);
  quad_positions(t, &q, rect[0], rect[1], rect[0] + rect[2], rect[1] + rect[3]);
  quad_uvs(&q, 0.0f, 0.0f, 1.0f, 1.0f);
  for (int k = 0; k < 4; k++) q.color[k] = unpack_color(cols[k]);
  q.mid = unpack_color(mid); q.stop = unpack_color(stop);
  if (fillMode == 0) q.factors = (v2){factor, spread};
  else q.factors = (v2){factor, clampf(fMid, 0.01f, 0.99f)};
  q.mode_packed = mode + (ell ? 128 : 0) + fillMode * 256;
  draw_quad(t, &q);
}

/* drawUvRect, glcontext.nim:1169-1302 */
static void op_uv_rect(tctx* t, float atx, float aty, float tox, float toy, float uax, float uay, float utx, float uty,
                       const uint32_t* cols, int call_index) {
  oquad_t q;
  memset(&q, 0, sizeof(q));
  q.call_index = call_index;
  quad_positions(t, &q, atx, aty, tox, toy);
  quad_uvs(&q, uax, uay, utx, uty);
  for (int k = 0; k < 4; k++) q.color[k] = unpack_color(cols[k]);
  q.mode_packed = M_ATLAS;
  draw_quad(t, &q);
}

static const uint64_t RECT_KEY = 0x7265637472656374ull; /* stands for hash("rect"), glcontext.nim:966 */

static int ensure_rect_image(tctx* t, float rect[4]) {
  /* the 4x4 white image is uploaded before the frame by orc_render (atlas is shared between threads) */
  return orc_get_image_rect((oracle*)t->sh->o, RECT_KEY, rect);
}

static void make_rect_mask(tctx* t, const float* rect, const float* rx, const float* ry) {
  rect_mask_t* rm = &t->rm_stack[t->n_rm++];
  v2 he = {rect[2] * 0.5f, rect[3] * 0.5f};
  float inv[16];
  if (!mat_inverse(t->mat, inv)) mat_identity(inv);
  int ell = 0;
  rm->fast = 1;
  rm->params = (v4){rect[0] + he.x, rect[1] + he.y, he.x, he.y};
  rm->radii = roundedRadiiVec(rx, ry, he, &ell);
  rm->matX = (v4){inv[0], inv[4], inv[12], 1.0f};
  rm->matY = (v4){inv[1], inv[5], inv[13], ell ? 1.0f : 0.0f};
}

static void begin_mask(tctx* t, const float* rect, const float* rx, const float* ry, int call_index) {
  shared_t* sh = t->sh;
  if (t->mask_begun) { t->error = 3; return; }
  t->mask_begun = 1;
  t->mask_write++;
  if (t->mask_write >= MAX_MASKS) { t->error = 4; t->mask_write = MAX_MASKS - 1; return; }
  t->level_quads[t->mask_write] = 0;
  t->level_ok[t->mask_write] = 0;
  /* until a mask quad is drawn the level is all zero: it clips everything */
  t->clip[t->mask_write][0] = t->clip[t->mask_write][1] = t->clip[t->mask_write][2] = t->clip[t->mask_write][3] = 0;
  /* glClear(0) of the whole mask texture, glcontext.nim:1901-1902 (own rows) */
  if (!sh->collect) memset(sh->masks[t->mask_write] + (size_t)t->y0 * sh->W, 0, (size_t)(t->y1 - t->y0) * sh->W);
  uint32_t red[4] = {0xFF0000FFu, 0, 0, 0};
  op_rounded_rect(t, rect, rx, ry, M_CLIP_AA, 4.0f, 0.0f, 0.0f, 0.0f, FILL_COLOR, 0, red, 0.5f, call_index);
}

/* blur.frag + runBackdropSeparableBlur glcontext.nim:1743-1786.  src/dst RGBA8, rows [y0,y1). */
static void blur_pass(const shared_t* sh, const uint8_t* src, uint8_t* dst, int y0, int y1, float radius_in, int vertical) {
  int W = sh->W, H = sh->H;
  float radius = clampf(radius_in, 0.0f, 64.0f);
  float sigma = fmaxf(0.5f * radius, 0.5f);
  float stepPx = fmaxf(radius / 8.0f, 1.0f);
  float tsx = vertical ? 0.0f : 1.0f / (float)W, tsy = vertical ? 1.0f / (float)H : 0.0f;
  for (int y = y0; y < y1; y++)
    for (int x = 0; x < W; x++) {
      uint8_t* d = dst + ((size_t)y * W + x) * 4;
      if (radius <= 0.5f) { memcpy(d, src + ((size_t)y * W + x) * 4, 4); continue; }
      float u = ((float)x + 0.5f) / (float)W, v = ((float)y + 0.5f) / (float)H;
      float acc[4] = {0, 0, 0, 0}, wsum = 0.0f;
      for (int i = -8; i <= 8; i++) {
        float xx = (float)i * stepPx;
        float w = expf(-0.5f * (xx * xx) / (sigma * sigma));
        /* texture(srcTex, uv + texelStep * xx): LINEAR, CLAMP_TO_EDGE */
        float tu = (u + tsx * xx) * (float)W - 0.5f, tv = (v + tsy * xx) * (float)H - 0.5f;
        float fx = floorf(tu), fy = floorf(tv);
        float axw = tu - fx, ayw = tv - fy;
        int i0 = clamp_i((int)fx, 0, W - 1), i1 = clamp_i((int)fx + 1, 0, W - 1);
        int j0 = clamp_i((int)fy, 0, H - 1), j1 = clamp_i((int)fy + 1, 0, H - 1);
        const uint8_t* t00 = src + ((size_t)j0 * W + i0) * 4; const uint8_t* t10 = src + ((size_t)j0 * W + i1) * 4;
        const uint8_t* t01 = src + ((size_t)j1 * W + i0) * 4; const uint8_t* t11 = src + ((size_t)j1 * W + i1) * 4;
        for (int c = 0; c < 4; c++) {
          float a = (float)t00[c] / 255.0f, b = (float)t10[c] / 255.0f, cc = (float)t01[c] / 255.0f, dd = (float)t11[c] / 255.0f;
          float top = a * (1.0f - axw) + b * axw, bot = cc * (1.0f - axw) + dd * axw;
          acc[c] += (top * (1.0f - ayw) + bot * ayw) * w;
        }
        wsum += w;
      }
      float inv = fmaxf(wsum, 1e-5f);
      for (int c = 0; c < 4; c++) d[c] = quant8(acc[c] / inv);
    }
}

static void barrier(void) {
#ifdef _OPENMP
#pragma omp barrier
#endif
}

static void exec_call(tctx* t, const call_t* c, int idx) {
  shared_t* sh = t->sh;
  const float* f = c->f;
  const uint32_t* u = c->u;
  switch (c->op) {
    case OP_SAVE: if (t->n_stack < 64) memcpy(t->stack[t->n_stack++], t->mat, 64); else t->error = 4; break;
    case OP_RESTORE: if (t->n_stack > 0) memcpy(t->mat, t->stack[--t->n_stack], 64); else t->error = 3; break;
    case OP_TRANSLATE: { float m[16]; mat_identity(m); m[12] = f[0]; m[13] = f[1]; mat_mul(t->mat, t->mat, m); break; }
    case OP_ROTATE: { /* vmath rotateZ: +angle turns +x toward -y on screen (pinned by render_line_rect.png) */
      float m[16]; mat_identity(m);
      float cs = cosf(f[0]), sn = sinf(f[0]);
      m[0] = cs; m[1] = -sn; m[4] = sn; m[5] = cs;
      mat_mul(t->mat, t->mat, m); break; }
    case OP_SCALE: { float m[16]; mat_identity(m); m[0] = f[0]; m[5] = f[1]; mat_mul(t->mat, t->mat, m); break; }
    case OP_APPLY: mat_mul(t->mat, t->mat, f); break;
    case OP_SET_AA: t->aa = f[0]; break;
    case OP_SET_SUBPIXEL: t->subpixel_enabled = (int)u[0]; t->subpixel_shift = f[0]; break;
    case OP_BEGIN_MASK: begin_mask(t, f, f + 4, f + 8, idx); break;
    case OP_END_MASK: if (!t->mask_begun) t->error = 3; t->mask_begun = 0; break;
    case OP_POP_MASK: if (t->mask_write <= 0) t->error = 3; else t->mask_write--; break;
    case OP_BEGIN_RECT_MASK:
      if (t->mask_begun) { t->error = 3; break; }
      if (t->n_rm == 0 && f[2] > 0.0f && f[3] > 0.0f) make_rect_mask(t, f, f + 4, f + 8);
      else { begin_mask(t, f, f + 4, f + 8, idx); t->mask_begun = 0; t->rm_stack[t->n_rm++].fast = 0; }
      break;
    case OP_POP_RECT_MASK:
      if (t->n_rm <= 0) { t->error = 3; break; }
      if (!t->rm_stack[--t->n_rm].fast) { if (t->mask_write > 0) t->mask_write--; else t->error = 3; }
      break;
    case OP_BACKDROP_BLUR: {
      if (f[12] <= 0.0f || f[2] <= 0.0f || f[3] <= 0.0f) break;
      if (sh->collect) {
        collect_begin_segment(t);
        uint32_t whitec[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        op_rounded_rect(t, f, f + 4, f + 8, M_BACKDROP, f[12], 0.0f, 0.0f, 0.0f, FILL_COLORS4, 0, whitec, 0.5f, idx);
        break;
      }
      /* glCopyTexSubImage2D(full frame) then H and V passes over the whole frame */
      barrier();
      memcpy(sh->backdrop + (size_t)t->y0 * sh->W * 4, sh->fb + (size_t)t->y0 * sh->W * 4, (size_t)(t->y1 - t->y0) * sh->W * 4);
      if (f[12] > 0.5f) {
        blur_pass(sh, sh->backdrop, sh->temp, t->y0, t->y1, f[12], 0);
        barrier();
        blur_pass(sh, sh->temp, sh->backdrop, t->y0, t->y1, f[12], 1);
      }
      barrier();
      uint32_t white[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
      op_rounded_rect(t, f, f + 4, f + 8, M_BACKDROP, f[12], 0.0f, 0.0f, 0.0f, FILL_COLORS4, 0, white, 0.5f, idx);
      break; }
    case OP_ROUNDED_RECT:
      op_rounded_rect(t, f, f + 4, f + 8, (int)u[0], f[12], f[13], f[14], f[15], (int)u[1], (int)u[2], u + 3, f[16], idx);
      break;
    case OP_IMAGE: { /* drawImage glcontext.nim:1350-1367 */
      float r[4];
      uint64_t key = (uint64_t)u[0] | ((uint64_t)u[1] << 32);
      if (!orc_get_image_rect((oracle*)sh->o, key, r)) break;
      float as = (float)sh->o->atlas_size;
      float sw = f[2], shh = f[3];
      if (!(sw > 0.0f && shh > 0.0f)) { sw = r[2] * as; shh = r[3] * as; }
      float uay = r[1], uty = r[1] + r[3];
      if (u[7]) { uay = r[1] + r[3]; uty = r[1]; }
      op_uv_rect(t, f[0], f[1], f[0] + sw, f[1] + shh, r[0], uay, r[0] + r[2], uty, u + 3, idx);
      break; }
    case OP_MSDF: { /* drawMsdfImage / drawMtsdfImage glcontext.nim:1097-1155, drawUvRectAtlasSdf :1022-1095 */
      float r[4];
      uint64_t key = (uint64_t)u[0] | ((uint64_t)u[1] << 32);
      if (!orc_get_image_rect((oracle*)sh->o, key, r)) break;
      float strokeW = fmaxf(0.0f, f[6]);
      int mtsdf = (int)u[2];
      int mode = strokeW > 0.0f ? (mtsdf ? M_MTSDF_ANN : M_MSDF_ANN) : (mtsdf ? M_MTSDF : M_MSDF);
      oquad_t q;
      memset(&q, 0, sizeof(q));
      q.call_index = idx;
      quad_positions(t, &q, f[0], f[1], f[0] + f[2], f[1] + f[3]);
      float uay = r[1], uty = r[1] + r[3];
      if (u[7]) { uay = r[1] + r[3]; uty = r[1]; }
      quad_uvs(&q, r[0], uay, r[0] + r[2], uty);
      for (int k = 0; k < 4; k++) q.color[k] = unpack_color(u[3]);
      q.params = (v4){(float)sh->o->atlas_size, strokeW, 0.0f, 0.0f};
      q.factors = (v2){f[4], f[5]};
      q.mode_packed = mode;
      draw_quad(t, &q);
      break; }
    case OP_BEZIER: { /* drawQuadraticBezierSdf glcontext.nim:1619-1741 */
      if (f[2] <= 0.0f || f[3] <= 0.0f || f[10] <= 0.0f) break;
      oquad_t q;
      memset(&q, 0, sizeof(q));
      q.call_index = idx;
      int fkind = (int)u[1], axis = (int)u[2], fillMode = 0;
      uint32_t cols[4], mid = 0, stop = 0;
      if (fkind == FILL_LIN3) { fillMode = 1 + (axis & 3); cols[0] = cols[1] = cols[2] = cols[3] = u[3]; mid = u[4]; stop = u[5]; }
      else gradientColors(fkind, axis, u + 3, f[16], cols);
      q.params = (v4){f[2] * 0.5f, f[3] * 0.5f, f[4], f[5]};
      q.radii = (v4){f[6], f[7], f[8], f[9]};
      quad_positions(t, &q, f[0], f[1], f[0] + f[2], f[1] + f[3]);
      quad_uvs(&q, 0.0f, 0.0f, 1.0f, 1.0f);
      for (int k = 0; k < 4; k++) q.color[k] = unpack_color(cols[k]);
      q.mid = unpack_color(mid); q.stop = unpack_color(stop);
      q.factors = fillMode == 0 ? (v2){f[10], 0.0f} : (v2){f[10], clampf(f[16], 0.01f, 0.99f)};
      int cap = (int)u[0];
      int mode = cap == CAP_BUTT ? M_BEZ_BUTT : (cap == CAP_SQUARE ? M_BEZ_SQUARE : M_BEZ);
      q.mode_packed = mode + fillMode * 256;
      draw_quad(t, &q);
      break; }
    case OP_FILLED_QUAD: { /* drawFilledQuad glcontext.nim:963-982 + drawQuad :908-961 */
      float r[4];
      if (!ensure_rect_image(t, r)) break;
      oquad_t q;
      memset(&q, 0, sizeof(q));
      q.call_index = idx;
      for (int k = 0; k < 4; k++) {
        v2 p = mat_apply(t->mat, f[2 * k], f[2 * k + 1]);
        q.pos[k].x = ceilf(p.x); q.pos[k].y = ceilf(p.y);
        q.uv[k].x = r[0] + r[2] / 2.0f; q.uv[k].y = r[1] + r[3] / 2.0f;
        q.color[k] = unpack_color(u[3 + k]);
      }
      q.mode_packed = M_ATLAS;
      draw_quad(t, &q);
      break; }
    case OP_RECT: { /* drawRect glcontext.nim:1402-1418 */
      float r[4];
      if (!ensure_rect_image(t, r)) break;
      uint32_t cols[4] = {u[3], u[3], u[3], u[3]};
      float cu = r[0] + r[2] / 2.0f, cv = r[1] + r[3] / 2.0f;
      op_uv_rect(t, f[0], f[1], f[0] + f[2], f[1] + f[3], cu, cv, cu, cv, cols, idx);
      break; }
    default: break;
  }
}

/* Renders one frame.  fb_inout: W*H*4 RGBA8 top-left origin; when clear != 0 it is first filled with clear_rgba
 * (glClear, glcontext.nim:2086-2091), otherwise its content is kept (GL keeps the back buffer).
 * frag_counts: optional int64[2*N_MODES] = fragments shaded per SdfMode, solid fills then gradient fills
 * (for algorithmic flop counts).
 * Returns 0 or an fdc_status-like code. */
int orc_render_rows(oracle* o, int W, int H, int clear, const float* clear_rgba, const call_t* calls, int64_t n_calls,
                    uint8_t* fb_inout, int64_t* frag_counts, int n_threads, int row0, int row1);
int orc_render(oracle* o, int W, int H, int clear, const float* clear_rgba, const call_t* calls, int64_t n_calls,
               uint8_t* fb_inout, int64_t* frag_counts, int n_threads) {
  return orc_render_rows(o, W, H, clear, clear_rgba, calls, n_calls, fb_inout, frag_counts, n_threads, 0, H);
}
/* As orc_render but only rows [row0,row1) are produced (a bounded sample of a large frame for CPU timing).
 * Backdrop blurs read rows outside the sample, so samples are only meaningful for scenes without blur. */
int orc_render_rows(oracle* o, int W, int H, int clear, const float* clear_rgba, const call_t* calls, int64_t n_calls,
                    uint8_t* fb_inout, int64_t* frag_counts, int n_threads, int row0, int row1) {
  if (W <= 0 || H <= 0) return 1;
  if (row0 < 0) row0 = 0;
  if (row1 > H) row1 = H;
  if (row1 <= row0) return 1;
  int needs_rect = 0, max_depth = 1, depth = 0, needs_blur = 0;
  for (int64_t i = 0; i < n_calls; i++) {
    uint32_t op = calls[i].op;
    if (op == OP_FILLED_QUAD || op == OP_RECT) needs_rect = 1;
    if (op == OP_BACKDROP_BLUR) needs_blur = 1;
    if (op == OP_BEGIN_MASK || op == OP_BEGIN_RECT_MASK) { depth++; if (depth + 1 > max_depth) max_depth = depth + 1; }
    if (op == OP_POP_MASK || op == OP_POP_RECT_MASK) depth--;
  }
  if (max_depth >= MAX_MASKS) return 4;
  float dummy[4];
  if (needs_rect && !orc_get_image_rect(o, RECT_KEY, dummy)) {
    uint8_t white[64];
    memset(white, 255, 64);
    orc_put_image(o, RECT_KEY, 4, 4, white, NULL);
  }
  shared_t sh;
  memset(&sh, 0, sizeof(sh));
  sh.o = o; sh.W = W; sh.H = H; sh.fb = fb_inout;
  for (int l = 1; l <= max_depth; l++) sh.masks[l] = (uint8_t*)calloc((size_t)W * H, 1);
  if (needs_blur) { sh.backdrop = (uint8_t*)calloc((size_t)W * H, 4); sh.temp = (uint8_t*)calloc((size_t)W * H, 4); }
  if (clear) {
    uint8_t c8[4] = {quant8(clear_rgba[0]), quant8(clear_rgba[1]), quant8(clear_rgba[2]), quant8(clear_rgba[3])};
    for (size_t i = 0; i < (size_t)W * H; i++) memcpy(fb_inout + i * 4, c8, 4);
  }
  if (n_threads < 1) n_threads = 1;
  if (n_threads > row1 - row0) n_threads = row1 - row0;
  int err = 0;
  int64_t totals[2 * N_MODES];
  memset(totals, 0, sizeof(totals));
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    int tid = 0, nt = 1;
#endif
    tctx* t = (tctx*)calloc(1, sizeof(tctx));
    t->sh = &sh;
    /* bands of whole 16-row groups keep the split independent of nothing but nt */
    int rows = (row1 - row0 + nt - 1) / nt;
    t->y0 = row0 + tid * rows; t->y1 = t->y0 + rows;
    if (t->y0 > row1) t->y0 = row1;
    if (t->y1 > row1) t->y1 = row1;
    mat_identity(t->mat);
    t->aa = 1.2f; /* DefaultSdfAaFactor, figbackend.nim:34 */
    for (int64_t i = 0; i < n_calls; i++) exec_call(t, &calls[i], (int)i);
    if (t->mask_write != 0 || t->n_rm != 0 || t->mask_begun) t->error = t->error ? t->error : 3; /* endFrame asserts */
#ifdef _OPENMP
#pragma omp critical
#endif
    {
      if (t->error && !err) err = t->error;
      for (int m = 0; m < 2 * N_MODES; m++) totals[m] += t->frag_counts[m];
    }
    free(t);
  }
  if (frag_counts) memcpy(frag_counts, totals, sizeof(totals));
  for (int l = 1; l <= max_depth; l++) free(sh.masks[l]);
  free(sh.backdrop);
  free(sh.temp);
  return err;
}

/* Fragment counts per SdfMode ([mode] solid, [N_MODES+mode] gradient) without shading anything. */
int orc_count_fragments(oracle* o, int W, int H, const call_t* calls, int64_t n_calls, int64_t* frag_counts, int n_threads) {
  if (W <= 0 || H <= 0) return 1;
  int needs_rect = 0, max_depth = 1, depth = 0;
  for (int64_t i = 0; i < n_calls; i++) {
    uint32_t op = calls[i].op;
    if (op == OP_FILLED_QUAD || op == OP_RECT) needs_rect = 1;
    if (op == OP_BEGIN_MASK || op == OP_BEGIN_RECT_MASK) { depth++; if (depth + 1 > max_depth) max_depth = depth + 1; }
    if (op == OP_POP_MASK || op == OP_POP_RECT_MASK) depth--;
  }
  if (max_depth >= MAX_MASKS) return 4;
  float dummy[4];
  if (needs_rect && !orc_get_image_rect(o, RECT_KEY, dummy)) {
    uint8_t white[64];
    memset(white, 255, 64);
    orc_put_image(o, RECT_KEY, 4, 4, white, NULL);
  }
  shared_t sh;
  memset(&sh, 0, sizeof(sh));
  sh.o = o; sh.W = W; sh.H = H; sh.count_only = 1;
  uint8_t* scratch = (uint8_t*)calloc((size_t)W * H, 1); /* one shared dummy mask level: only cleared, never read */
  for (int l = 1; l <= max_depth; l++) sh.masks[l] = scratch;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > H) n_threads = H;
  int64_t totals[2 * N_MODES];
  memset(totals, 0, sizeof(totals));
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    int tid = 0, nt = 1;
#endif
    tctx* t = (tctx*)calloc(1, sizeof(tctx));
    t->sh = &sh;
    int rows = (H + nt - 1) / nt;
    t->y0 = tid * rows; t->y1 = t->y0 + rows;
    if (t->y0 > H) t->y0 = H;
    if (t->y1 > H) t->y1 = H;
    mat_identity(t->mat);
    t->aa = 1.2f;
    for (int64_t i = 0; i < n_calls; i++)
      if (calls[i].op != OP_BACKDROP_BLUR) exec_call(t, &calls[i], (int)i);
      else if (calls[i].f[12] > 0.0f && calls[i].f[2] > 0.0f && calls[i].f[3] > 0.0f) {
        uint32_t whitec[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        op_rounded_rect(t, calls[i].f, calls[i].f + 4, calls[i].f + 8, M_BACKDROP, calls[i].f[12], 0.0f, 0.0f, 0.0f, FILL_COLORS4, 0,
                        whitec, 0.5f, (int)i);
      }
#ifdef _OPENMP
#pragma omp critical
#endif
    for (int m = 0; m < 2 * N_MODES; m++) totals[m] += t->frag_counts[m];
    free(t);
  }
  memcpy(frag_counts, totals, sizeof(totals));
  free(scratch);
  return 0;
}

int orc_n_modes(void) { return N_MODES; }
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Reference bin boxes.  Writes up to `cap` records of 8 int32 (segment, call_index, x0, y0, x1, y1, is_mask, level)
 * in emission order and returns how many the frame has.  Single threaded. */
int64_t orc_collect_quads(oracle* o, int W, int H, int band_y0, int band_y1, const call_t* calls, int64_t n_calls,
                          int32_t* out, int64_t cap) {
  int needs_rect = 0;
  for (int64_t i = 0; i < n_calls; i++)
    if (calls[i].op == OP_FILLED_QUAD || calls[i].op == OP_RECT) needs_rect = 1;
  float dummy[4];
  if (needs_rect && !orc_get_image_rect(o, RECT_KEY, dummy)) {
    uint8_t white[64];
    memset(white, 255, 64);
    orc_put_image(o, RECT_KEY, 4, 4, white, NULL);
  }
  shared_t sh;
  memset(&sh, 0, sizeof(sh));
  sh.o = o; sh.W = W; sh.H = H;
  sh.collect = 1; sh.rec = out; sh.rec_cap = cap; sh.band_y0 = band_y0; sh.band_y1 = band_y1;
  tctx* t = (tctx*)calloc(1, sizeof(tctx));
  t->sh = &sh;
  t->y0 = 0; t->y1 = H;
  mat_identity(t->mat);
  t->aa = 1.2f;
  t->clip[0][0] = -40000; t->clip[0][1] = -40000; t->clip[0][2] = 40000; t->clip[0][3] = 40000;
  for (int64_t i = 0; i < n_calls; i++) exec_call(t, &calls[i], (int)i);
  free(t);
  return sh.rec_n;
}
