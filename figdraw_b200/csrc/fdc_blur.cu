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

#include <algorithm>

#include "fdc_kernels.h"

namespace fdc {

namespace {

struct BlurParams {
  float w[9];      // weights for |i| = 0..8
  float inv_sum;   // 1 / max(sum, 1e-5)
  float step;      // tap spacing in pixels
  int reach;       // ceil(8*step) + 1
  int copy_only;   // radius <= 0.5: texture() pass-through
  int tap_off[17]; // floor(i * step), i = -8..8  (the same float arithmetic as blur.frag, done once on the host)
  float tap_fr[17];// i * step - floor(i * step)
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
constexpr int kMaxReach = 66;   // ceil(8 * 64/8) + 1 + slack

struct RowSources {
  const uint32_t* rank[kMaxRanks];
  int n;
  int band_end[kMaxRanks];  // first row NOT owned by rank r
};
__device__ __forceinline__ const uint32_t* row_base(const RowSources& rs, const uint32_t* src, int y) {
  if (rs.n == 0) return src;
  int r = 0;
  while (r < rs.n - 1 && y >= rs.band_end[r]) r++;
  return rs.rank[r];
}

// 17 taps over a line of unpacked texels (`stride` float4s apart); c = index of the centre texel.  Frame-edge clamping
// happened when staging; (k0, k0+1) are neighbours in the clamped line only when the unclamped indices are both inside
// or both outside the frame, which holds because clamping is monotone.
__device__ __forceinline__ uint32_t blur_taps(const float4* __restrict__ line, int stride, int c, const BlurParams& bp) {
  if (bp.copy_only) {
    const float4 t = line[c * stride];
    return quant_pack(t);
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 17; i++) {
    const int k0 = c + bp.tap_off[i];
    const float fr = bp.tap_fr[i];
    const float w = bp.w[i < 8 ? 8 - i : i - 8];
    const float4 t0 = line[k0 * stride];
    if (fr > 0.0f) {
      const float4 t1 = line[(k0 + 1) * stride];
      acc.x = fmaf(fmaf(t1.x - t0.x, fr, t0.x), w, acc.x);
      acc.y = fmaf(fmaf(t1.y - t0.y, fr, t0.y), w, acc.y);
      acc.z = fmaf(fmaf(t1.z - t0.z, fr, t0.z), w, acc.z);
      acc.w = fmaf(fmaf(t1.w - t0.w, fr, t0.w), w, acc.w);
    } else {
      acc.x = fmaf(t0.x, w, acc.x); acc.y = fmaf(t0.y, w, acc.y); acc.z = fmaf(t0.z, w, acc.z); acc.w = fmaf(t0.w, w, acc.w);
    }
  }
  acc.x *= bp.inv_sum; acc.y *= bp.inv_sum; acc.z *= bp.inv_sum; acc.w *= bp.inv_sum;
  return quant_pack(acc);
}

// Horizontal pass: one row segment of kBlurTile pixels per CTA, staged UNPACKED (float4 per texel: each staged texel
// is converted once instead of once per tap that touches it) with its tap halo.  Region [x0,x1) x [y0,y1) of dst is
// produced.  Source indices clamp to the frame (CLAMP_TO_EDGE).  Row `y` may live in a neighbour's framebuffer (halo
// rows of a band partition): volatile loads, the line must come from the owner's L2, not from a stale local cache.
__global__ void __launch_bounds__(kBlurTile) blur_h_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int W, int H,
                                                           int x0, int y0, int x1, int y1, BlurParams bp, RowSources rs) {
  __shared__ float4 line[kBlurTile + 2 * kMaxReach];
  const int xa = x0 + blockIdx.x * kBlurTile;
  const int y = y0 + blockIdx.y;
  if (y >= y1) return;
  const int span = kBlurTile + 2 * bp.reach;
  const uint32_t* row = row_base(rs, src, y) + (size_t)y * W;
  for (int k = threadIdx.x; k < span; k += kBlurTile) {
    int a = xa - bp.reach + k;
    a = a < 0 ? 0 : (a >= W ? W - 1 : a);
    line[k] = unpack255(rs.n == 0 ? __ldg(row + a) : __ldcv(row + a));
  }
  __syncthreads();
  const int x = xa + threadIdx.x;
  if (x >= x1) return;
  dst[(size_t)y * W + x] = blur_taps(line, 1, threadIdx.x + bp.reach, bp);
}

// Vertical pass as a 2-D tile: 32 columns x kVTile rows of output per CTA.  The source rows (tile + tap halo) are staged
// row by row, so every global load and store is 32 consecutive pixels (a line-per-thread layout would make a warp
// touch 32 different rows); a thread then walks down its column in shared memory.  Dynamic shared memory:
// (kVTile + 2*reach) x 32 float4.
constexpr int kVTile = 32;
constexpr int kVRows = 8;  // thread rows per CTA
__global__ void __launch_bounds__(32 * kVRows) blur_v_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int W, int H,
                                                             int x0, int y0, int x1, int y1, BlurParams bp) {
  extern __shared__ float4 col[];  // [span][32]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = x0 + blockIdx.x * 32 + tx;
  const int ybase = y0 + blockIdx.y * kVTile;
  const int span = kVTile + 2 * bp.reach;
  const int xs = min(x, W - 1);
  for (int k = ty; k < span; k += kVRows) {
    int a = ybase - bp.reach + k;
    a = a < 0 ? 0 : (a >= H ? H - 1 : a);  // CLAMP_TO_EDGE
    col[k * 32 + tx] = unpack255(__ldg(src + (size_t)a * W + xs));
  }
  __syncthreads();
  if (x >= x1) return;
  for (int r = ty; r < kVTile; r += kVRows) {
    const int y = ybase + r;
    if (y >= y1) break;
    dst[(size_t)y * W + x] = blur_taps(col + tx, 32, r + bp.reach, bp);
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
  for (int i = -8; i <= 8; i++) {
    const float off = (float)i * bp.step;
    const float fo = floorf(off);
    bp.tap_off[i + 8] = (int)fo;
    bp.tap_fr[i + 8] = off - fo;
  }
  const uint32_t* src = reinterpret_cast<const uint32_t*>(a.src);
  RowSources rs;
  rs.n = a.n_src;
  for (int r = 0; r < kMaxRanks; r++) {
    rs.rank[r] = r < a.n_src ? reinterpret_cast<const uint32_t*>(a.src_rank[r]) : nullptr;
    rs.band_end[r] = r < a.n_src ? a.band_end_px[r] : 0;
  }
  uint32_t* temp = reinterpret_cast<uint32_t*>(a.temp);
  uint32_t* dst = reinterpret_cast<uint32_t*>(a.dst);
  // H pass over the rows the V pass will read
  const int hy0 = max(a.y0 - bp.reach, 0), hy1 = min(a.y1 + bp.reach, a.H);
  {
    dim3 grid((a.x1 - a.x0 + kBlurTile - 1) / kBlurTile, hy1 - hy0);
    blur_h_kernel<<<grid, kBlurTile, 0, stream>>>(src, temp, a.W, a.H, a.x0, hy0, a.x1, hy1, bp, rs);
  }
  {
    // (kVTile + 2*kMaxReach) x 32 float4 = 82 KB at the largest radius: above the 48 KB default, per device
    const size_t max_smem = (size_t)(kVTile + 2 * kMaxReach) * 32 * sizeof(float4);
    cudaFuncSetAttribute(blur_v_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
    const size_t smem = (size_t)(kVTile + 2 * bp.reach) * 32 * sizeof(float4);
    dim3 grid((a.x1 - a.x0 + 31) / 32, (a.y1 - a.y0 + kVTile - 1) / kVTile);
    blur_v_kernel<<<grid, 32 * kVRows, smem, stream>>>(temp, dst, a.W, a.H, a.x0, a.y0, a.x1, a.y1, bp);
  }
  if (n_launches) *n_launches += 2;
}

namespace {
struct FlagPtrs {
  uint32_t* p[kMaxRanks];
};
// The barrier VALUES live on the device (two counters in the rank's own flag page), not in kernel arguments: every
// rank runs the same sequence of signals and waits, so "the k-th wait waits for the k-th signal of every rank" needs no
// host bookkeeping -- and a frame captured into a CUDA graph can be replayed with the barriers inside it.
__global__ void signal_flags_kernel(FlagPtrs f, int n, int my_rank) {
  __shared__ uint32_t value;
  if (threadIdx.x == 0) value = ++f.p[my_rank][kFlagSignalSeq];
  __syncthreads();
  const int r = threadIdx.x;
  if (r >= n) return;
  __threadfence_system();  // everything this stream wrote before (the segment's pixels) is visible system-wide first
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[r] + my_rank), "r"(value) : "memory");
}
// Spin until every rank's slot reached this rank's next wait number.  A peer that never arrives (crashed process, a
// frame submitted on one rank only) must not hang the GPU: after kBarrierTimeoutNs the kernel records the failure in
// flags[kFlagError] and lets the stream continue; the host reports it when the frame is resolved.
constexpr unsigned long long kBarrierTimeoutNs = 2000000000ull;
__global__ void wait_flags_kernel(uint32_t* flags, int n) {
  __shared__ uint32_t value;
  if (threadIdx.x == 0) value = ++flags[kFlagWaitSeq];
  __syncthreads();
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

namespace {
__global__ void __launch_bounds__(256) push_to_peers_kernel(const uint4* __restrict__ src, size_t byte_off, size_t n16, uint8_t* multicast,
                                                            uint8_t* const* __restrict__ peers, int n_peers, int my_rank) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = src[i];
    if (multicast) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(multicast + byte_off + i * 16),
                   "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                   : "memory");
    } else {
      for (int r = 0; r < n_peers; r++)
        if (r != my_rank && peers[r]) *reinterpret_cast<uint4*>(peers[r] + byte_off + i * 16) = v;
    }
  }
}
}  // namespace

void launch_push_to_peers(const uint8_t* src, size_t byte_off, size_t bytes, uint8_t* multicast, uint8_t* const* peers, int n_peers,
                          int my_rank, cudaStream_t stream) {
  const size_t n16 = bytes / 16;
  if (n16 == 0) return;
  const unsigned grid = (unsigned)std::min<size_t>((n16 + 255) / 256, 148 * 4);
  push_to_peers_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(src), byte_off, n16, multicast, peers, n_peers, my_rank);
}

void launch_signal_flags(uint32_t* const* flag_arrays, int n, int my_rank, cudaStream_t stream) {
  FlagPtrs f;
  for (int r = 0; r < kMaxRanks; r++) f.p[r] = r < n ? flag_arrays[r] : nullptr;
  signal_flags_kernel<<<1, 32, 0, stream>>>(f, n, my_rank);
}
void launch_wait_flags(uint32_t* my_flags, int n, cudaStream_t stream) {
  wait_flags_kernel<<<1, 32, 0, stream>>>(my_flags, n);
}

void launch_mip_down(const uint8_t* src, int src_size, uint8_t* dst, int dst_size, int sx, int sy, int sw, int sh, int dx,
                     int dy, cudaStream_t stream) {
  const int dw = sw / 2, dh = sh / 2;
  if (dw <= 0 || dh <= 0) return;
  dim3 block(16, 16), grid((dw + 15) / 16, (dh + 15) / 16);
  mip_down_kernel<<<grid, block, 0, stream>>>(src, src_size, dst, dst_size, sx, sy, dw, dh, dx, dy);
}

}  // namespace fdc
