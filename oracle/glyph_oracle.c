/* glyph_oracle.c -- CPU restatement of the glyph coverage rasteriser (TEST INFRASTRUCTURE ONLY: loaded by tests/ and
 * __graft_entry__.smoke(), never by the product).
 *
 * What it stands in for: the reference rasterises one glyph at a time on the CPU with pixie (`image.fillText`,
 * src/figdraw/common/textrasters/pixie_raster.nim:45-95) and optionally applies FreeType's 5-tap LCD filter
 * (`applyLcdFilter`, :12-43) before `putImage`.  pixie is a third-party Nim package that is NOT vendored under
 * /root/reference (figdraw.nimble:8 `pixie >= 5.0.1`, no lock file), so its anti-aliasing cannot be reproduced bit for
 * bit here: PARITY UNPINNED.  This file restates the published algorithm the CUDA path implements instead -- exact
 * area coverage by signed-area accumulation (font-rs `Raster::draw_line`, Raph Levien 2016; the same scheme as
 * stb_truetype v2) -- sequentially, in float32 with -ffp-contract=off, and tests/test_glyph_raster.py checks IT
 * against brute-force supersampling.  The LCD filter is the reference's integer arithmetic, line for line in meaning.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x0, y0, x1, y1, cx, cy; uint32_t kind, pad; } seg_t; /* = fdc_outline_seg */

static void draw_line(float* acc, int stride, int w, int h, float x0, float y0, float x1, float y1) {
  if (y0 == y1) return;
  float dir = 1.0f;
  if (y0 > y1) {
    dir = -1.0f;
    float t = x0; x0 = x1; x1 = t;
    t = y0; y0 = y1; y1 = t;
  }
  float dxdy = (x1 - x0) / (y1 - y0);
  float x = x0;
  if (y0 < 0.0f) { x -= y0 * dxdy; y0 = 0.0f; }
  int ya = (int)fmaxf(floorf(y0), 0.0f), yb = (int)ceilf(y1);
  if (yb > h) yb = h;
  for (int y = ya; y < yb; y++) {
    float dy = fminf((float)(y + 1), y1) - fmaxf((float)y, y0);
    float xnext = x + dxdy * dy;
    float d = dy * dir;
    float xa = fminf(x, xnext), xb = fmaxf(x, xnext);
    xa = fminf(fmaxf(xa, 0.0f), (float)w);
    xb = fminf(fmaxf(xb, 0.0f), (float)w);
    float* row = acc + (size_t)y * stride;
    float x0f = floorf(xa);
    int x0i = (int)x0f;
    float x1c = ceilf(xb);
    int x1i = (int)x1c;
    if (x1i <= x0i + 1) {
      float xmf = 0.5f * (xa + xb) - x0f;
      row[x0i] += d - d * xmf;
      row[x0i + 1] += d * xmf;
    } else {
      float s = 1.0f / (xb - xa);
      float x0fr = xa - x0f;
      float a0 = 0.5f * s * (1.0f - x0fr) * (1.0f - x0fr);
      float x1fr = xb - x1c + 1.0f;
      float am = 0.5f * s * x1fr * x1fr;
      row[x0i] += d * a0;
      if (x1i == x0i + 2) {
        row[x0i + 1] += d * (1.0f - a0 - am);
      } else {
        float a1 = s * (1.5f - x0fr);
        row[x0i + 1] += d * (a1 - a0);
        for (int xi = x0i + 2; xi < x1i - 1; xi++) row[xi] += d * s;
        float a2 = a1 + (float)(x1i - x0i - 3) * s;
        row[x1i - 1] += d * (1.0f - a2 - am);
      }
      row[x1i] += d * am;
    }
    x = xnext;
  }
}

/* One glyph: `n` outline segments -> w x h straight-alpha RGBA8 (white, alpha = coverage). */
int orc_rasterize_glyph(const seg_t* segs, int n, int w, int h, int lcd_filter, uint8_t* out_rgba) {
  if (w <= 0 || h <= 0) return 1;
  const int stride = w + 2;
  float* acc = (float*)calloc((size_t)stride * h, sizeof(float));
  uint8_t* cov = (uint8_t*)malloc((size_t)w * h);
  if (!acc || !cov) return 2;
  const float tol = 0.025f; /* max chord deviation of a flattened quadratic, pixels */
  for (int i = 0; i < n; i++) {
    const seg_t* s = &segs[i];
    if (s->kind == 0) {
      draw_line(acc, stride, w, h, s->x0, s->y0, s->x1, s->y1);
    } else {
      float ddx = s->x0 - 2.0f * s->cx + s->x1, ddy = s->y0 - 2.0f * s->cy + s->y1;
      float dd = sqrtf(ddx * ddx + ddy * ddy);
      int nn = (int)ceilf(sqrtf(dd / (4.0f * tol)));
      if (nn < 1) nn = 1;
      if (nn > 64) nn = 64;
      float px = s->x0, py = s->y0;
      for (int k = 1; k <= nn; k++) {
        float t = (float)k / (float)nn, mt = 1.0f - t;
        float qx = k == nn ? s->x1 : mt * mt * s->x0 + 2.0f * mt * t * s->cx + t * t * s->x1;
        float qy = k == nn ? s->y1 : mt * mt * s->y0 + 2.0f * mt * t * s->cy + t * t * s->y1;
        draw_line(acc, stride, w, h, px, py, qx, qy);
        px = qx; py = qy;
      }
    }
  }
  for (int y = 0; y < h; y++) {
    float run = 0.0f;
    for (int x = 0; x < w; x++) {
      run += acc[(size_t)y * stride + x];
      cov[(size_t)y * w + x] = (uint8_t)lrintf(fminf(fabsf(run), 1.0f) * 255.0f);
    }
  }
  static const int wts[5] = {8, 77, 86, 77, 8}; /* lcdFilterWeights, pixie_raster.nim:12 */
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      uint32_t a = cov[(size_t)y * w + x];
      if (lcd_filter) { /* applyLcdFilter, pixie_raster.nim:15-43 */
        int sum = 0;
        for (int k = 0; k < 5; k++) {
          int sx = x + k - 2;
          if (sx < 0) sx = 0;
          if (sx > w - 1) sx = w - 1;
          sum += (int)cov[(size_t)y * w + sx] * wts[k];
        }
        a = (uint32_t)((sum + 128) >> 8);
      }
      uint8_t* q = out_rgba + ((size_t)y * w + x) * 4;
      q[0] = q[1] = q[2] = a ? 255 : 0; /* premultiplied white -> straight alpha, textures.nim:90-92 */
      q[3] = (uint8_t)a;
    }
  free(acc);
  free(cov);
  return 0;
}
