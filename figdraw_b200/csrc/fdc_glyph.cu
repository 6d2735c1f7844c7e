// Glyph coverage rasterisation on the GPU.
//
// Replaces, for hosts that hand over outlines instead of bitmaps, the CPU rasterisation the reference does per glyph
// with pixie before `putImage` (common/textrasters/pixie_raster.nim:45-95: typeset one rune, `image.fillText`, optional
// `applyLcdFilter` :12-43, `loadGlyphImage`).  Input: the glyph's outline as line / quadratic-Bezier segments in the
// pixel space of its bitmap (x right, y down, closed contours, outer contours and holes wound in opposite senses, as
// TrueType stores them).  Output: straight-alpha white texels (255,255,255,coverage) written STRAIGHT INTO THE ATLAS
// slot the packer assigned -- no host bitmap, no upload -- followed by the usual mip chain.
//
// Coverage is exact area coverage by signed-area accumulation (the scheme of font-rs / stb_truetype v2): every line
// deposits, per pixel row it crosses, the signed area it cuts off to its right into an accumulation row; a running sum
// along x turns that into winding-weighted coverage; |sum| clamped to 1 is the pixel's alpha.  One CTA per glyph, the
// accumulation rows in shared memory (strips of rows when the bitmap is large), float atomics for the deposits.
//
// PARITY UNPINNED: pixie (not vendored, no lock file) anti-aliases with its own scheme (sub-scanline sampling), so
// these bitmaps are NOT expected to equal pixie's bit for bit; the oracle for this file is oracle/glyph_oracle.c, the
// same algorithm written sequentially, itself checked against brute-force supersampling (tests/test_glyph_raster.py).
// The LCD filter is the reference's integer arithmetic exactly.
#include <cuda_runtime.h>

#include "fdc_kernels.h"

namespace fdc {

namespace {

constexpr int kGlyphThreads = 128;
constexpr int kAccFloats = 12288;       // 48 KB of accumulation rows per CTA
constexpr float kFlatTolerance = 0.025f;  // max chord deviation of a flattened quadratic, pixels

// Deposit the line (x0,y0)-(x1,y1), y in strip-local rows [0, rows), into acc[row * stride + x].
__device__ void deposit_line(float* acc, int stride, int w, int rows, float x0, float y0, float x1, float y1) {
  if (y0 == y1) return;
  float dir = 1.0f;
  if (y0 > y1) {
    dir = -1.0f;
    float t = x0; x0 = x1; x1 = t;
    t = y0; y0 = y1; y1 = t;
  }
  const float dxdy = (x1 - x0) / (y1 - y0);
  float x = x0;
  if (y0 < 0.0f) { x -= y0 * dxdy; y0 = 0.0f; }
  const int ya = (int)fmaxf(floorf(y0), 0.0f), yb = min((int)ceilf(y1), rows);
  for (int y = ya; y < yb; y++) {
    const float dy = fminf((float)(y + 1), y1) - fmaxf((float)y, y0);
    const float xnext = x + dxdy * dy;
    const float d = dy * dir;
    float xa = fminf(x, xnext), xb = fmaxf(x, xnext);
    // everything left of the bitmap still shadows every pixel of the row; right of it shadows none
    xa = fminf(fmaxf(xa, 0.0f), (float)w);
    xb = fminf(fmaxf(xb, 0.0f), (float)w);
    float* row = acc + y * stride;
    const float x0f = floorf(xa);
    const int x0i = (int)x0f;
    const float x1c = ceilf(xb);
    const int x1i = (int)x1c;
    if (x1i <= x0i + 1) {
      const float xmf = 0.5f * (xa + xb) - x0f;  // mean x inside the column
      atomicAdd(&row[x0i], d - d * xmf);
      atomicAdd(&row[x0i + 1], d * xmf);
    } else {
      const float s = 1.0f / (xb - xa);
      const float x0fr = xa - x0f;
      const float a0 = 0.5f * s * (1.0f - x0fr) * (1.0f - x0fr);
      const float x1fr = xb - x1c + 1.0f;
      const float am = 0.5f * s * x1fr * x1fr;
      atomicAdd(&row[x0i], d * a0);
      if (x1i == x0i + 2) {
        atomicAdd(&row[x0i + 1], d * (1.0f - a0 - am));
      } else {
        const float a1 = s * (1.5f - x0fr);
        atomicAdd(&row[x0i + 1], d * (a1 - a0));
        for (int xi = x0i + 2; xi < x1i - 1; xi++) atomicAdd(&row[xi], d * s);
        const float a2 = a1 + (float)(x1i - x0i - 3) * s;
        atomicAdd(&row[x1i - 1], d * (1.0f - a2 - am));
      }
      atomicAdd(&row[x1i], d * am);
    }
    x = xnext;
  }
}

struct GlyphDev {
  uint32_t first_seg, n_segs;
  int32_t w, h;    // bitmap size
  int32_t ax, ay;  // atlas position of its top-left texel
};

__global__ void __launch_bounds__(kGlyphThreads) glyph_raster_kernel(const GlyphDev* __restrict__ glyphs,
                                                                     const fdc_outline_seg* __restrict__ segs, uint8_t* __restrict__ atlas,
                                                                     int atlas_size, int lcd_filter) {
  __shared__ float acc[kAccFloats];
  const GlyphDev g = glyphs[blockIdx.x];
  if (g.w <= 0 || g.h <= 0) return;
  const int stride = g.w + 2;  // a deposit may touch column w and w + 1
  const int strip = max(1, min(g.h, kAccFloats / stride));
  uint8_t* cov = reinterpret_cast<uint8_t*>(acc);  // the strip's alpha bytes overwrite the accumulation rows they came from
  for (int r0 = 0; r0 < g.h; r0 += strip) {
    const int rows = min(strip, g.h - r0);
    for (int i = threadIdx.x; i < rows * stride; i += kGlyphThreads) acc[i] = 0.0f;
    __syncthreads();
    for (uint32_t si = threadIdx.x; si < g.n_segs; si += kGlyphThreads) {
      const fdc_outline_seg s = segs[g.first_seg + si];
      const float oy = (float)r0;
      if (s.kind == 0u) {
        deposit_line(acc, stride, g.w, rows, s.x0, s.y0 - oy, s.x1, s.y1 - oy);
      } else {
        // flatten: n chords keep the deviation |P0 - 2 P1 + P2| / (4 n^2) under the tolerance
        const float ddx = s.x0 - 2.0f * s.cx + s.x1, ddy = s.y0 - 2.0f * s.cy + s.y1;
        const float dd = sqrtf(ddx * ddx + ddy * ddy);
        const int n = max(1, min(64, (int)ceilf(sqrtf(dd / (4.0f * kFlatTolerance)))));
        float px = s.x0, py = s.y0;
        for (int k = 1; k <= n; k++) {
          const float t = (float)k / (float)n, mt = 1.0f - t;
          const float qx = k == n ? s.x1 : mt * mt * s.x0 + 2.0f * mt * t * s.cx + t * t * s.x1;
          const float qy = k == n ? s.y1 : mt * mt * s.y0 + 2.0f * mt * t * s.cy + t * t * s.y1;
          deposit_line(acc, stride, g.w, rows, px, py - oy, qx, qy - oy);
          px = qx; py = qy;
        }
      }
    }
    __syncthreads();
    // running sum along x, one warp per row; coverage -> alpha byte (in place: byte x of row y at cov[y * stride * 4 + x])
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int y = warp; y < rows; y += kGlyphThreads / 32) {
      float carry = 0.0f;
      for (int xb = 0; xb < g.w; xb += 32) {
        const int x = xb + lane;
        float v = x < g.w ? acc[y * stride + x] : 0.0f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float t = __shfl_up_sync(0xFFFFFFFFu, v, o);
          if (lane >= o) v += t;
        }
        v += carry;
        carry = __shfl_sync(0xFFFFFFFFu, v, 31);
        __syncwarp();
        if (x < g.w) cov[(size_t)y * stride * 4 + x] = (uint8_t)__float2int_rn(fminf(fabsf(v), 1.0f) * 255.0f);
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rows * g.w; i += kGlyphThreads) {
      const int y = i / g.w, x = i - y * g.w;
      const uint8_t* row = cov + (size_t)y * stride * 4;
      uint32_t a = row[x];
      if (lcd_filter) {  // applyLcdFilter, pixie_raster.nim:12-43: FreeType's 5-tap weights, edge texels repeated
        const int wts[5] = {8, 77, 86, 77, 8};
        int sum = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) sum += (int)row[min(max(x + k - 2, 0), g.w - 1)] * wts[k];
        a = (uint32_t)((sum + 128) >> 8);
      }
      // white text, premultiplied (a,a,a,a) in pixie -> straight alpha on upload (textures.nim:90-92): rgb 255 where a > 0
      const uint32_t px = a ? (0x00FFFFFFu | (a << 24)) : 0u;
      reinterpret_cast<uint32_t*>(atlas)[(size_t)(g.ay + r0 + y) * atlas_size + g.ax + x] = px;
    }
    __syncthreads();
  }
}

}  // namespace

void launch_glyph_raster(const void* glyphs_dev, int n_glyphs, const fdc_outline_seg* segs_dev, uint8_t* atlas_level0, int atlas_size,
                         int lcd_filter, cudaStream_t stream) {
  if (n_glyphs <= 0) return;
  glyph_raster_kernel<<<n_glyphs, kGlyphThreads, 0, stream>>>(reinterpret_cast<const GlyphDev*>(glyphs_dev), segs_dev, atlas_level0,
                                                              atlas_size, lcd_filter);
}

}  // namespace fdc
