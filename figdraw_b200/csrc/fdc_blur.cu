// Separable Gaussian backdrop blur (blur.frag:1-32, runBackdropSeparableBlur glcontext.nim:1743-1786) and the
// atlas mip-chain builder (textures.nim:106-119).
//
// The reference copies the WHOLE frame and runs both passes over the WHOLE frame for every blur node; only the
// pixels under the composite quad are ever sampled afterwards (atlas.frag:381-388).  Here the vertical pass
// runs on the quad's bbox and the horizontal pass on that bbox grown by the vertical tap reach -- identical
// results, a fraction of the traffic.  Each pass stages its source rows/columns in shared memory with the tap
// halo, so every source texel is read from L2/HBM once per pass, and both passes store RGBA8 (the reference
// renders them into RGBA8 textures, so the intermediate is quantised too).
#include <cuda_runtime.h>
#include <math.h>

#include "fdc_kernels.h"

namespace fdc {

namespace {

struct BlurParams {
  float w[9];      // weights for |i| = 0..8
  float inv_sum;   // 1 / max(sum, 1e-5)
  float step;      // tap spacing in pixels
  int reach;       // ceil(8*step) + 1
  int copy_only;   // radius <= 0.5: texture() pass-through
};

__device__ __forceinline__ float4 unpack255(uint32_t c) {
  return make_float4((float)(c & 255u), (float)((c >> 8) & 255u), (float)((c >> 16) & 255u), (float)(c >> 24));
}
__device__ __forceinline__ uint32_t quant_pack(float4 v) {  // values already in 0..255
  const uint32_t r = (uint32_t)__float2int_rn(fminf(fmaxf(v.x, 0.0f), 255.0f));
  const uint32_t g = (uint32_t)__float2int_rn(fminf(fmaxf(v.y, 0.0f), 255.0f));
  const uint32_t b = (uint32_t)__float2int_rn(fminf(fmaxf(v.z, 0.0f), 255.0f));
  const uint32_t a = (uint32_t)__float2int_rn(fminf(fmaxf(v.w, 0.0f), 255.0f));
  return r | (g << 8) | (b << 16) | (a << 24);
}

constexpr int kBlurTile = 128;  // pixels along the pass axis per CTA
#ifndef FDC_BLUR_LINES
#define FDC_BLUR_LINES 1
#endif
constexpr int kBlurLines = FDC_BLUR_LINES;   // lines (rows for H, columns for V) per CTA
constexpr int kMaxReach = 66;   // ceil(8 * 64/8) + 1 + slack

// One pass.  kVertical=false: taps along x, reads `src` rows; kVertical=true: taps along y.
// Region [x0,x1) x [y0,y1) of dst is produced.  Source indices clamp to the frame (CLAMP_TO_EDGE).
struct RowSources {
  const uint32_t* rank[kMaxRanks];
  int n, band_px;
};
__device__ __forceinline__ const uint32_t* row_base(const RowSources& rs, const uint32_t* src, int y) {
  if (rs.n == 0) return src;
  const int r = min(y / rs.band_px, rs.n - 1);
  return rs.rank[r];
}

template <bool kVertical>
__global__ void __launch_bounds__(kBlurTile) blur_pass_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                             int W, int H, int x0, int y0, int x1, int y1, BlurParams bp,
                                                             RowSources rs) {
  __shared__ uint32_t line[kBlurLines][kBlurTile + 2 * kMaxReach];
  const int along0 = (kVertical ? y0 : x0) + blockIdx.x * kBlurTile;  // first pixel along the pass axis
  const int across0 = (kVertical ? x0 : y0) + blockIdx.y * kBlurLines;
  const int along_end = kVertical ? y1 : x1, across_end = kVertical ? x1 : y1;
  const int limit = kVertical ? H : W;
  const int span = kBlurTile + 2 * bp.reach;
  for (int l = 0; l < kBlurLines; l++) {
    const int across = across0 + l;
    if (across >= across_end) break;
    for (int k = threadIdx.x; k < span; k += kBlurTile) {
      int a = along0 - bp.reach + k;
      a = a < 0 ? 0 : (a >= limit ? limit - 1 : a);
      // H pass: row `across` may live in a neighbour's framebuffer (halo rows of a band partition): volatile load, the
      // line must come from the owner's L2, not from a stale local cache
      if (kVertical) line[l][k] = __ldg(src + (size_t)a * W + across);
      else if (rs.n == 0) line[l][k] = __ldg(src + (size_t)across * W + a);
      else line[l][k] = __ldcv(row_base(rs, src, across) + (size_t)across * W + a);
    }
  }
  __syncthreads();
  const int along = along0 + threadIdx.x;
  if (along >= along_end) return;
  for (int l = 0; l < kBlurLines; l++) {
    const int across = across0 + l;
    if (across >= across_end) break;
    uint32_t out;
    const int c = threadIdx.x + bp.reach;  // centre index in the staged line
    if (bp.copy_only) {
      out = line[l][c];
    } else {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = -8; i <= 8; i++) {
        const float off = (float)i * bp.step;
        const float fo = floorf(off);
        const float fr = off - fo;
        const int k0 = c + (int)fo;
        // Frame-edge clamping happened when staging; (k0, k0+1) are neighbours in the clamped line only when the
        // unclamped indices are both inside or both outside the frame, which holds because clamping is monotone.
        const float4 t0 = unpack255(line[l][k0]);
        const float w = bp.w[i < 0 ? -i : i];
        if (fr > 0.0f) {
          const float4 t1 = unpack255(line[l][k0 + 1]);
          acc.x = fmaf(fmaf(t1.x - t0.x, fr, t0.x), w, acc.x);
          acc.y = fmaf(fmaf(t1.y - t0.y, fr, t0.y), w, acc.y);
          acc.z = fmaf(fmaf(t1.z - t0.z, fr, t0.z), w, acc.z);
          acc.w = fmaf(fmaf(t1.w - t0.w, fr, t0.w), w, acc.w);
        } else {
          acc.x = fmaf(t0.x, w, acc.x); acc.y = fmaf(t0.y, w, acc.y); acc.z = fmaf(t0.z, w, acc.z); acc.w = fmaf(t0.w, w, acc.w);
        }
      }
      acc.x *= bp.inv_sum; acc.y *= bp.inv_sum; acc.z *= bp.inv_sum; acc.w *= bp.inv_sum;
      out = quant_pack(acc);
    }
    if (kVertical) dst[(size_t)along * W + across] = out;
    else dst[(size_t)across * W + along] = out;
  }
}

// Vertical pass as a 2-D tile: 32 columns x kVTile rows of output per CTA.  The source rows (tile + tap halo) are staged
// row by row, so every global load and store is 32 consecutive pixels (the line-per-thread layout of the horizontal
// pass would make a warp touch 32 different rows); a thread then walks down its column in shared memory (bank = column).
constexpr int kVTile = 64;
constexpr int kVRows = 8;  // thread rows per CTA
__global__ void __launch_bounds__(32 * kVRows) blur_v_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int W, int H,
                                                             int x0, int y0, int x1, int y1, BlurParams bp) {
  __shared__ uint32_t col[kVTile + 2 * kMaxReach][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = x0 + blockIdx.x * 32 + tx;
  const int ybase = y0 + blockIdx.y * kVTile;
  const int span = kVTile + 2 * bp.reach;
  const int xs = min(x, W - 1);
  for (int k = ty; k < span; k += kVRows) {
    int a = ybase - bp.reach + k;
    a = a < 0 ? 0 : (a >= H ? H - 1 : a);  // CLAMP_TO_EDGE
    col[k][tx] = __ldg(src + (size_t)a * W + xs);
  }
  __syncthreads();
  if (x >= x1) return;
  for (int r = ty; r < kVTile; r += kVRows) {
    const int y = ybase + r;
    if (y >= y1) break;
    const int c = r + bp.reach;
    uint32_t out;
    if (bp.copy_only) {
      out = col[c][tx];
    } else {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = -8; i <= 8; i++) {
        const float off = (float)i * bp.step;
        const float fo = floorf(off);
        const float fr = off - fo;
        const int k0 = c + (int)fo;
        const float4 t0 = unpack255(col[k0][tx]);
        const float w = bp.w[i < 0 ? -i : i];
        if (fr > 0.0f) {
          const float4 t1 = unpack255(col[k0 + 1][tx]);
          acc.x = fmaf(fmaf(t1.x - t0.x, fr, t0.x), w, acc.x);
          acc.y = fmaf(fmaf(t1.y - t0.y, fr, t0.y), w, acc.y);
          acc.z = fmaf(fmaf(t1.z - t0.z, fr, t0.z), w, acc.z);
          acc.w = fmaf(fmaf(t1.w - t0.w, fr, t0.w), w, acc.w);
        } else {
          acc.x = fmaf(t0.x, w, acc.x); acc.y = fmaf(t0.y, w, acc.y); acc.z = fmaf(t0.z, w, acc.z); acc.w = fmaf(t0.w, w, acc.w);
        }
      }
      acc.x *= bp.inv_sum; acc.y *= bp.inv_sum; acc.z *= bp.inv_sum; acc.w *= bp.inv_sum;
      out = quant_pack(acc);
    }
    dst[(size_t)y * W + x] = out;
  }
}

// 2x2 box on premultiplied colour, back to straight alpha (same integer arithmetic as the oracle's upload_chain).
__global__ void mip_down_kernel(const uint8_t* __restrict__ src, int src_size, uint8_t* __restrict__ dst, int dst_size,
                                int sx, int sy, int dw, int dh, int dx, int dy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= dw || j >= dh) return;
  uint32_t acc[4] = {0, 0, 0, 0};
  for (int dj = 0; dj < 2; dj++)
    for (int di = 0; di < 2; di++) {
      const uint8_t* p = src + ((size_t)(sy + 2 * j + dj) * src_size + (sx + 2 * i + di)) * 4;
      const uint32_t a = p[3];
      acc[0] += (p[0] * a + 127) / 255;
      acc[1] += (p[1] * a + 127) / 255;
      acc[2] += (p[2] * a + 127) / 255;
      acc[3] += a;
    }
  const int ox = dx + i, oy = dy + j;
  if (ox < 0 || oy < 0 || ox >= dst_size || oy >= dst_size) return;
  uint8_t* q = dst + ((size_t)oy * dst_size + ox) * 4;
  const uint32_t a = (acc[3] + 2) >> 2;
  for (int c = 0; c < 3; c++) {
    const uint32_t pm = (acc[c] + 2) >> 2;
    const uint32_t s = a ? (pm * 255 + a / 2) / a : 0;
    q[c] = (uint8_t)(s > 255 ? 255 : s);
  }
  q[3] = (uint8_t)a;
}

}  // namespace

void launch_backdrop_blur(const BlurArgs& a, cudaStream_t stream, int* n_launches) {
  if (a.x0 >= a.x1 || a.y0 >= a.y1) return;
  BlurParams bp;
  const float radius = fminf(fmaxf(a.radius, 0.0f), 64.0f);  // blur.frag:12
  bp.copy_only = radius <= 0.5f;
  const float sigma = fmaxf(0.5f * radius, 0.5f);
  bp.step = fmaxf(radius / 8.0f, 1.0f);
  float sum = 0.0f;
  for (int i = -8; i <= 8; i++) {  // same accumulation order as the shader loop
    const float x = (float)i * bp.step;
    const float w = expf(-0.5f * (x * x) / (sigma * sigma));
    bp.w[i < 0 ? -i : i] = w;
    sum += w;
  }
  bp.inv_sum = 1.0f / fmaxf(sum, 1e-5f);
  bp.reach = bp.copy_only ? 0 : (int)ceilf(8.0f * bp.step) + 1;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(a.src);
  RowSources rs;
  rs.n = a.n_src;
  rs.band_px = a.band_px > 0 ? a.band_px : 1;
  for (int r = 0; r < kMaxRanks; r++) rs.rank[r] = r < a.n_src ? reinterpret_cast<const uint32_t*>(a.src_rank[r]) : nullptr;
  uint32_t* temp = reinterpret_cast<uint32_t*>(a.temp);
  uint32_t* dst = reinterpret_cast<uint32_t*>(a.dst);
  // H pass over the rows the V pass will read
  const int hy0 = max(a.y0 - bp.reach, 0), hy1 = min(a.y1 + bp.reach, a.H);
  {
    dim3 grid((a.x1 - a.x0 + kBlurTile - 1) / kBlurTile, (hy1 - hy0 + kBlurLines - 1) / kBlurLines);
    blur_pass_kernel<false><<<grid, kBlurTile, 0, stream>>>(src, temp, a.W, a.H, a.x0, hy0, a.x1, hy1, bp, rs);
  }
  {
    dim3 grid((a.x1 - a.x0 + 31) / 32, (a.y1 - a.y0 + kVTile - 1) / kVTile);
    blur_v_kernel<<<grid, 32 * kVRows, 0, stream>>>(temp, dst, a.W, a.H, a.x0, a.y0, a.x1, a.y1, bp);
  }
  if (n_launches) *n_launches += 2;
}

namespace {
struct FlagPtrs {
  uint32_t* p[kMaxRanks];
};
__global__ void signal_flags_kernel(FlagPtrs f, int n, int my_rank, uint32_t value) {
  const int r = threadIdx.x;
  if (r >= n) return;
  __threadfence_system();  // everything this stream wrote before (the segment's pixels) is visible system-wide first
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[r] + my_rank), "r"(value) : "memory");
}
// Spin until every rank's slot reached `value`.  A peer that never arrives (crashed process, a frame submitted on
// one rank only) must not hang the GPU: after kBarrierTimeoutNs the kernel records the failure in flags[kFlagError]
// and lets the stream continue; the host reports it when the frame is resolved.
constexpr unsigned long long kBarrierTimeoutNs = 2000000000ull;
__global__ void wait_flags_kernel(uint32_t* flags, int n, uint32_t value) {
  const int r = threadIdx.x;
  if (r >= n) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t v;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
    if ((int32_t)(v - value) >= 0) break;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > kBarrierTimeoutNs) {
      atomicOr(flags + kFlagError, 1u << r);
      break;
    }
    __nanosleep(200);
  }
}
}  // namespace

void launch_signal_flags(uint32_t* const* flag_arrays, int n, int my_rank, uint32_t value, cudaStream_t stream) {
  FlagPtrs f;
  for (int r = 0; r < kMaxRanks; r++) f.p[r] = r < n ? flag_arrays[r] : nullptr;
  signal_flags_kernel<<<1, 32, 0, stream>>>(f, n, my_rank, value);
}
void launch_wait_flags(uint32_t* my_flags, int n, uint32_t value, cudaStream_t stream) {
  wait_flags_kernel<<<1, 32, 0, stream>>>(my_flags, n, value);
}

void launch_mip_down(const uint8_t* src, int src_size, uint8_t* dst, int dst_size, int sx, int sy, int sw, int sh, int dx,
                     int dy, cudaStream_t stream) {
  const int dw = sw / 2, dh = sh / 2;
  if (dw <= 0 || dh <= 0) return;
  dim3 block(16, 16), grid((dw + 15) / 16, (dh + 15) / 16);
  mip_down_kernel<<<grid, block, 0, stream>>>(src, src_size, dst, dst_size, sx, sy, dw, dh, dx, dy);
}

}  // namespace fdc
