// Host side of the CUDA backend and its C ABI (include/figdraw_cuda.h).
//
// This is the C++ counterpart of `OpenGlContext` (src/figdraw/opengl/glcontext.nim): it keeps exactly the state
// the GL context keeps on the host -- transform stack (:1991-2017), AA factor (:1157-1167), mask / rect-mask
// stacks (:1873-1949), the atlas skyline packer (:541-586) -- but instead of filling vertex arrays it appends one
// 128-byte record per draw to a pinned staging buffer.  `fdc_end_frame` uploads the records and launches:
//     prim_setup -> coarse count -> coarse scan -> coarse scatter -> fine bin -> shade   [-> blur H -> blur V] ...
// once per segment (a segment ends at each backdrop blur, which must read everything painted before it).
// All quad arithmetic (ceil, radii packing, gradient colours, mode encoding) happens on the device.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "fdc_flatten.h"
#include "fdc_kernels.h"

using namespace fdc;

namespace {

thread_local std::string g_create_error;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    // slack: a scene that grows by a few records per frame must not reallocate (cudaFree is a device-wide sync)
    size_t want = std::max(n + n / 8 + 64, cap + cap / 2);
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  // grow keeping the first `keep` elements (device-to-device copy on `st`)
  cudaError_t reserve_keep(size_t n, size_t keep, cudaStream_t st) {
    if (n <= cap) return cudaSuccess;
    size_t want = std::max(n + n / 8 + 64, cap * 2);
    T* np = nullptr;
    cudaError_t e = cudaMalloc(&np, want * sizeof(T));
    if (e != cudaSuccess) return e;
    keep = std::min(keep, cap);
    if (p && keep) {
      e = cudaMemcpyAsync(np, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (p) cudaFree(p);
    p = np;
    cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <typename T>
struct PinnedBuf {
  T* p = nullptr;
  size_t cap = 0, n = 0;
  bool reserve(size_t want) {
    if (want <= cap) return true;
    size_t nc = std::max(want, std::max<size_t>(cap * 2, 1024));
    T* np = nullptr;
    if (cudaMallocHost(&np, nc * sizeof(T)) != cudaSuccess) return false;
    if (p) {
      memcpy(np, p, n * sizeof(T));
      cudaFreeHost(p);
    }
    p = np;
    cap = nc;
    return true;
  }
  bool push(const T& v) {
    if (n == cap && !reserve(n + 1)) return false;
    p[n++] = v;
    return true;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = n = 0;
  }
};

struct Mat4 {
  float m[16];  // vmath: m[col*4 + row]
};
Mat4 mat_identity() {
  Mat4 r;
  memset(r.m, 0, sizeof(r.m));
  r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
  return r;
}
// a * b, each entry summed left to right (vmath `*`); compiled with -ffp-contract=off
Mat4 mat_mul(const Mat4& a, const Mat4& b) {
  Mat4 r;
  for (int c = 0; c < 4; c++)
    for (int row = 0; row < 4; row++)
      r.m[c * 4 + row] = a.m[0 * 4 + row] * b.m[c * 4 + 0] + a.m[1 * 4 + row] * b.m[c * 4 + 1] +
                         a.m[2 * 4 + row] * b.m[c * 4 + 2] + a.m[3 * 4 + row] * b.m[c * 4 + 3];
  return r;
}
// 4x4 inverse by cofactors (`ctx.mat.inverse()`, glcontext.nim:837)
bool mat_inverse(const Mat4& a, Mat4& out) {
  const float* m = a.m;
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
  if (det == 0.0f) return false;
  float id = 1.0f / det;
  for (int i = 0; i < 16; i++) out.m[i] = t[i] * id;
  return true;
}

float clamp_radius_h(float r, float maxr) {
  if (r <= 0.0f) return 0.0f;
  return roundf(fmaxf(1.0f, fminf(r, maxr)));
}
// roundedRadiiVec (glcontext.nim:751-817), host copy used only for fast rect masks
bool rounded_radii_vec_h(const float* rx, const float* ry, float hx, float hy, float out[4]) {
  const int order[4] = {1, 3, 0, 2};
  bool circ = true;
  for (int i = 0; i < 4; i++) circ = circ && rx[i] == ry[i];
  float mr = fminf(hx, hy);
  if (circ) {
    for (int k = 0; k < 4; k++) out[k] = clamp_radius_h(rx[order[k]], mr);
    return false;
  }
  for (int k = 0; k < 4; k++) {
    int c = order[k];
    float cx = clamp_radius_h(rx[c], hx), cy = clamp_radius_h(ry[c], hy);
    if (rx[c] == ry[c]) out[k] = -(clamp_radius_h(rx[c], mr) + 1.0f);
    else if (cx == cy) out[k] = -(cx + 1.0f);
    else {
      float qx = roundf(fminf(fmaxf(cx / fmaxf(hx, 0.000001f), 0.0f), 1.0f) * 4095.0f);
      float qy = roundf(fminf(fmaxf(cy / fmaxf(hy, 0.000001f), 0.0f), 1.0f) * 4095.0f);
      out[k] = qx + qy * 4096.0f;
    }
  }
  return true;
}

struct Segment {
  uint32_t first = 0, count = 0;  // draws
  bool has_blur = false;          // a backdrop blur follows this segment
  float blur_radius = 0.0f;
  int rx0 = 0, ry0 = 0, rx1 = 0, ry1 = 0;  // blur region (superset of the composite quad bbox, clipped to the frame)
};

struct MaskLevel {
  std::vector<fdc_call> draws;   // mask primitives of this level (re-emitted at segment boundaries)
  std::vector<RunState> states;  // their run states (first_draw rewritten on emission)
  int32_t mask_draw = -1;        // draw index of the single clipping primitive, or -1 when not usable as a clip
  int32_t clip = -1;             // clip_draw for content at this level
  int32_t first_run = -1;        // index in ctx->runs of the first mask draw's run (patched to PF_MASK_WIDE when a second draw arrives)
};

struct RectMaskEntry {
  bool fast;
  uint32_t index;  // fast: index+1 into rectmask table
};

}  // namespace

struct fdc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string error;
  float pixel_scale = 1.0f;
  bool pixelate = false;  // `newContext(pixelate = true)`: the atlas magnifies with GL_NEAREST
  int rank = 0, n_ranks = 1;

  // ---- atlas
  int atlas_size = 0, initial_atlas_size = 0, n_levels = 0;
  uint8_t* levels[kMaxAtlasLevels] = {};
  std::vector<uint16_t> heights;
  struct Rect4 { float x, y, w, h; };
  std::unordered_map<uint64_t, Rect4> entries;
  // What the reference keeps beside `entries` (figbackend.nim:61-75 AtlasEntryMeta, :185-187 owner sets), for hosts that
  // let the library do the bookkeeping: texel rect + insertion order (re-packing on regrow), entry kind and ids
  // (clearFontGlyphs / clearTypefaceGlyphs), owner tokens (eviction when the last owner lets go).
  struct EntryInfo { int px = 0, py = 0, w = 0, h = 0; uint64_t order = 0; int kind = 0; uint64_t a = 0, b = 0; };
  std::unordered_map<uint64_t, EntryInfo> entry_info;
  std::unordered_map<uint64_t, std::vector<uint64_t>> image_owners, font_owners;
  uint64_t put_counter = 0, atlas_generation = 1, atlas_rebuilds = 0;
  bool atlas_replay = false;  // on regrow re-pack the live entries natively instead of dropping them
  // fdc_render_frame: two page-locked record buffers used alternately -- frame k+1 is flattened into one while the
  // asynchronous upload of frame k may still be reading the other (fdc_begin_frame(k+1) then waits for frame k)
  PinnedBuf<fdc_call> flat[2];
  int flat_idx = 0;
  struct { uint8_t* out = nullptr; int x = 0, y = 0, w = 0, h = 0; } pending_read;  // fdc_read_pixels_async in flight
  DevBuf<AtlasEntry> d_table;
  uint32_t table_cap = 0;
  bool table_dirty = true;

  // ---- backend state
  Mat4 mat = mat_identity();
  std::vector<Mat4> mats;
  float aa = 1.2f;  // DefaultSdfAaFactor figbackend.nim:34
  bool subpixel_enabled = false;
  float subpixel_shift = 0.0f;
  bool frame_begun = false, mask_begun = false;
  int mask_write = 0;
  MaskLevel mask_levels[kMaxMaskDepth + 2];
  std::vector<RectMaskEntry> rm_stack;

  // ---- frame recording
  int W = 0, H = 0;
  bool clear = true;
  uint32_t clear_rgba8 = 0xFFFFFFFFu;
  PinnedBuf<fdc_call> draws;  // staged draw records (individually issued draws and short runs)
  struct Upload { uint32_t dst, count; size_t staged_off; };
  std::vector<Upload> uploads;  // staged ranges still to be copied to d_draws at endFrame
  uint32_t n_draws = 0;         // draws of the frame so far (staged + directly uploaded)
  PinnedBuf<RunState> runs;
  PinnedBuf<Xform> xforms;
  PinnedBuf<RectMaskRec> rectmasks;
  std::vector<Segment> segments;
  bool state_dirty = true, xform_dirty = true, begin_pending = false;
  uint32_t call_ordinal = 0;       // backend calls seen this frame
  uint32_t last_draw_ordinal = 0;  // ordinal of the previous draw
  bool have_frame = false;         // a recorded frame is resident on the device (replay / debug)
  uint32_t n_replays = 0;          // frames re-run by resolve_frame after a bin-list regrow

  // ---- device frame data
  DevBuf<fdc_call> d_draws;
  DevBuf<RunState> d_runs;
  DevBuf<Xform> d_xforms;
  DevBuf<RectMaskRec> d_rectmasks;
  DevBuf<Prim> d_prims;
  DevBuf<PrimBin> d_prim_bins;
  DevBuf<QuadGeom> d_geoms;
  DevBuf<PrimExt> d_exts;
  DevBuf<uint32_t> d_prim_call;
  DevBuf<uint32_t> d_seg_table, d_coarse_list, d_tile_start, d_tile_count, d_counters;
  DevBuf<uint32_t> d_row_cost;          // tile entries per tile row of the last frame (fdc_get_tile_row_costs)
  std::vector<int> band_bounds;         // fdc_set_band_tile_rows: n_ranks + 1 tile-row boundaries; empty = equal bands
  DevBuf<TileEntry> d_tile_list;
  DevBuf<uint8_t> d_fb, d_backdrop, d_temp;
  DevBuf<uint8_t> d_snapshot;      // pre-frame pixels of a multi-segment frame without clearMain (restored before an overflow replay)
  bool snapshot_valid = false;
  bool frame_resolved = false;     // resolve_frame has already checked the frame in flight
  uint32_t dbg_coarse_limit = 0, dbg_tile_limit = 0;  // fdc_debug_limit_lists: pretend the bin lists are this small (0: real size)
  uint8_t* ext_fb = nullptr;
  // fdc_export_framebuffer: the framebuffer as a CUDA VMM allocation whose POSIX file descriptor a presenter imports
  struct { unsigned long long handle = 0, va = 0; size_t size = 0; int fd = -1; } exported;
  uint8_t* mc_fb = nullptr;        // NVSwitch multicast mapping of the (shared) framebuffer, or nullptr
  bool frame_barrier = false;      // end every frame with a cross-rank flag barrier (shared framebuffer: the gather is fused)
  DevBuf<uint8_t*> d_peers;
  DevBuf<fdc_rect64> d_rects64;    // compact draw records of this frame (fdc_submit_rects64)
  uint32_t n_rects64 = 0;
  // Record exchange area behind the flags of a shared framebuffer (tile-band partitions): every rank uploads only ITS
  // 1/n_ranks slice of a long run of compact records over its own PCIe link and pushes the slice into all copies over
  // NVLink (multicast or peer stores) -- the host->device traffic of the whole job is one copy of the stream, not n.
  size_t rec_off = 0, rec_bytes = 0;
  bool rec_shared_frame = false;   // this frame's compact records live in the exchange area
  struct Exchange { uint32_t first, count; };  // record range this rank uploaded and has to push to the peers
  std::vector<Exchange> exchanges;
  std::vector<uint8_t*> h_peers;   // host copy of the peer framebuffer pointers (own entry = own framebuffer)
  size_t flag_off = 0;             // byte offset of the cross-rank flag array inside a reserved framebuffer (0: none)
  uint32_t frame_barrier_base = 0; // barrier_seq at the start of the frame in flight
  uint32_t barrier_seq = 0;        // cross-rank barriers issued so far (the values themselves are counted on the device)
  // fdc_replay_frame: the resident frame's launches captured once into a CUDA graph, then one cudaGraphLaunch per replay
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_enabled = true, capturing = false;
  DevBuf<unsigned long long> d_stats;
  bool want_stats = false;
  int n_peers = 0;
  size_t fb_bytes = 0;

  // ---- stats
  fdc_frame_stats stats = {};
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  int gather_mode = FDC_GATHER_STORES, gather_sub_bands = 4;
  static constexpr int kCopyStreams = 4, kMaxSubBands = 16;
  cudaStream_t copy_streams[kCopyStreams] = {};
  cudaEvent_t ev_sub[kMaxSubBands] = {}, ev_copy[kCopyStreams] = {};
  std::vector<cudaEvent_t> ev_pool;
  struct Span { int a, b, kind; };  // event indices, kind 0 bin 1 shade 2 blur
  std::vector<Span> spans;
  size_t ev_used = 0;

  FrameView frame = {};

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    error = buf;
    return code;
  }
  int cuda_fail(cudaError_t e, const char* what) { return fail(FDC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e)); }
  uint8_t* fb() { return ext_fb ? ext_fb : d_fb.p; }
};

#define CK(expr)                                              \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ctx->cuda_fail(_e, #expr);  \
  } while (0)

namespace {

int atlas_alloc(fdc_ctx* ctx, int size, bool free_old = true) {
  if (free_old) {
    for (int l = 0; l < ctx->n_levels; l++) {
      cudaFree(ctx->levels[l]);
      ctx->levels[l] = nullptr;
    }
  }
  ctx->n_levels = 0;
  ctx->atlas_size = size;
  for (int s = size; s >= 1 && ctx->n_levels < kMaxAtlasLevels; s >>= 1) {
    uint8_t* p = nullptr;
    CK(cudaMalloc(&p, (size_t)s * s * 4));
    CK(cudaMemsetAsync(p, 0, (size_t)s * s * 4, ctx->stream));  // glGenerateMipmap of an empty texture
    ctx->levels[ctx->n_levels++] = p;
  }
  ctx->heights.assign((size_t)size, 0);
  ctx->entries.clear();
  ctx->table_dirty = true;
  return FDC_OK;
}

int find_empty_rect(fdc_ctx* ctx, int width, int height, int* rx, int* ry, bool* grew);
int build_mips(fdc_ctx* ctx, int x, int y, int w, int h);

// grow (glcontext.nim:536-539) = resetImageAtlas at twice the size.  The reference drops every entry and relies on the
// host to replay its images (noteAtlasRebuilt -> replayImageMessages, figbackend.nim:202-207); with atlas_replay the
// library does it: the live entries are packed into the new atlas in their original insertion order and their texels
// copied device-to-device out of the old one (removed entries are not carried over, which is what reclaims their space).
int atlas_grow(fdc_ctx* ctx) {
  ctx->atlas_generation++;
  ctx->atlas_rebuilds++;
  if (!ctx->atlas_replay) {
    ctx->entry_info.clear();
    return atlas_alloc(ctx, ctx->atlas_size * 2);
  }
  uint8_t* old_levels[kMaxAtlasLevels];
  const int old_n = ctx->n_levels, old_size = ctx->atlas_size;
  for (int l = 0; l < old_n; l++) old_levels[l] = ctx->levels[l];
  std::vector<std::pair<uint64_t, uint64_t>> live;  // (insertion order, key)
  for (auto& kv : ctx->entries) {
    auto it = ctx->entry_info.find(kv.first);
    if (it != ctx->entry_info.end()) live.push_back({it->second.order, kv.first});
  }
  std::sort(live.begin(), live.end());
  int size = old_size;
  for (;;) {  // a size at which everything fits again
    size *= 2;
    if (size > 16384) return ctx->fail(FDC_ERR_CAPACITY, "atlas cannot grow beyond 16384");
    int rc = atlas_alloc(ctx, size, false);
    if (rc) return rc;
    bool ok = true;
    for (auto& ok_key : live) {
      fdc_ctx::EntryInfo& e = ctx->entry_info[ok_key.second];
      int rx = 0, ry = 0;
      // find_empty_rect must not recurse into another grow here: probe the height map directly
      const int imgW = e.w + kAtlasMargin * 2, imgH = e.h + kAtlasMargin * 2;
      int lowest = ctx->atlas_size, at = 0;
      for (int i = 0; i < ctx->atlas_size; i++) {
        const int v = ctx->heights[i];
        if (v < lowest) {
          bool fit = true;
          for (int j = 0; j <= imgW; j++) {
            if (i + j >= ctx->atlas_size || (int)ctx->heights[i + j] > v) { fit = false; break; }
          }
          if (fit) { lowest = v; at = i; }
        }
      }
      if (lowest + imgH > ctx->atlas_size) { ok = false; break; }
      for (int j = at; j < at + imgW; j++) ctx->heights[j] = (uint16_t)(lowest + imgH + kAtlasMargin * 2);
      rx = at + kAtlasMargin;
      ry = lowest + kAtlasMargin;
      if (e.w > 1 && e.h > 1) {
        CK(cudaMemcpy2DAsync(ctx->levels[0] + ((size_t)ry * size + rx) * 4, (size_t)size * 4,
                             old_levels[0] + ((size_t)e.py * old_size + e.px) * 4, (size_t)old_size * 4, (size_t)e.w * 4, (size_t)e.h,
                             cudaMemcpyDeviceToDevice, ctx->stream));
        rc = build_mips(ctx, rx, ry, e.w, e.h);
        if (rc) return rc;
      }
      e.px = rx; e.py = ry;
      const float as = (float)size;
      ctx->entries[ok_key.second] = {(float)rx / as, (float)ry / as, (float)e.w / as, (float)e.h / as};
    }
    if (ok) break;
    for (int l = 0; l < ctx->n_levels; l++) cudaFree(ctx->levels[l]);  // still too small: try the next size
  }
  CK(cudaStreamSynchronize(ctx->stream));
  for (int l = 0; l < old_n; l++) cudaFree(old_levels[l]);
  // entry_info of keys that are no longer live goes too
  for (auto it = ctx->entry_info.begin(); it != ctx->entry_info.end();)
    it = ctx->entries.count(it->first) ? std::next(it) : ctx->entry_info.erase(it);
  ctx->table_dirty = true;
  return FDC_OK;
}

// findEmptyRect (glcontext.nim:541-579).  *grew is set when the atlas doubled (all entries dropped).
int find_empty_rect(fdc_ctx* ctx, int width, int height, int* rx, int* ry, bool* grew) {
  for (;;) {
    const int imgW = width + kAtlasMargin * 2, imgH = height + kAtlasMargin * 2;
    int lowest = ctx->atlas_size, at = 0;
    for (int i = 0; i < ctx->atlas_size; i++) {
      const int v = ctx->heights[i];
      if (v < lowest) {
        bool fit = true;
        for (int j = 0; j <= imgW; j++) {
          if (i + j >= ctx->atlas_size) { fit = false; break; }
          if ((int)ctx->heights[i + j] > v) { fit = false; break; }
        }
        if (fit) { lowest = v; at = i; }
      }
    }
    if (lowest + imgH > ctx->atlas_size) {
      if (ctx->atlas_size >= 16384) return ctx->fail(FDC_ERR_CAPACITY, "atlas cannot grow beyond 16384 for a %dx%d image", width, height);
      int rc = atlas_grow(ctx);  // grow -> resetImageAtlas, glcontext.nim:536-539
      if (rc) return rc;
      *grew = true;
      continue;
    }
    for (int j = at; j < at + imgW; j++) ctx->heights[j] = (uint16_t)(lowest + imgH + kAtlasMargin * 2);
    *rx = at + kAtlasMargin;
    *ry = lowest + kAtlasMargin;
    return FDC_OK;
  }
}

// The mip chain of a level-0 region: GPU box filter per level while w > 1 && h > 1 (updateSubImage, textures.nim:106-119).
int build_mips(fdc_ctx* ctx, int x, int y, int w, int h) {
  int level = 0;
  while (true) {
    int nw = w / 2, nh = h / 2;
    if (!(nw > 1 && nh > 1) || level + 1 >= ctx->n_levels) break;
    launch_mip_down(ctx->levels[level], ctx->atlas_size >> level, ctx->levels[level + 1], ctx->atlas_size >> (level + 1), x, y, w, h,
                    x / 2, y / 2, ctx->stream);
    x /= 2; y /= 2; w = nw; h = nh;
    level++;
  }
  CK(cudaGetLastError());
  return FDC_OK;
}

int upload_chain(fdc_ctx* ctx, int x, int y, int w, int h, const uint8_t* rgba) {
  // updateSubImage (textures.nim:106-119): level 0 copy, then the mip chain
  if (!(w > 1 && h > 1)) return FDC_OK;  // the reference's loop uploads nothing for 1-pixel-wide images
  CK(cudaMemcpy2DAsync(ctx->levels[0] + ((size_t)y * ctx->atlas_size + x) * 4, (size_t)ctx->atlas_size * 4, rgba, (size_t)w * 4,
                       (size_t)w * 4, (size_t)h, cudaMemcpyHostToDevice, ctx->stream));
  return build_mips(ctx, x, y, w, h);
}

int sync_table(fdc_ctx* ctx) {
  if (!ctx->table_dirty) return FDC_OK;
  uint32_t cap = 64;
  while (cap < ctx->entries.size() * 2 + 2) cap *= 2;
  std::vector<AtlasEntry> tab(cap);
  memset(tab.data(), 0, cap * sizeof(AtlasEntry));
  for (auto& kv : ctx->entries) {
    uint32_t h = atlas_hash(kv.first) & (cap - 1);
    while (tab[h].used) h = (h + 1) & (cap - 1);
    tab[h].key = kv.first;
    tab[h].used = 1;
    tab[h].x = kv.second.x; tab[h].y = kv.second.y; tab[h].w = kv.second.w; tab[h].h = kv.second.h;
  }
  CK(ctx->d_table.reserve(cap));
  CK(cudaMemcpyAsync(ctx->d_table.p, tab.data(), cap * sizeof(AtlasEntry), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // `tab` is pageable and about to go out of scope
  ctx->table_cap = cap;
  ctx->table_dirty = false;
  return FDC_OK;
}

AtlasView atlas_view(fdc_ctx* ctx) {
  AtlasView v;
  memset(&v, 0, sizeof(v));
  for (int l = 0; l < ctx->n_levels; l++) v.level[l] = ctx->levels[l];
  v.table = ctx->d_table.p;
  v.table_mask = ctx->table_cap ? ctx->table_cap - 1 : 0;
  v.size = ctx->atlas_size;
  v.n_levels = ctx->n_levels;
  v.pixelate = ctx->pixelate ? 1 : 0;
  return v;
}

int put_image_impl(fdc_ctx* ctx, uint64_t key, int w, int h, const uint8_t* rgba, float out_rect[4], int* out_rebuilt) {
  if (w <= 0 || h <= 0 || !rgba) return ctx->fail(FDC_ERR_INVALID, "putImage: bad image %dx%d", w, h);
  int rx = 0, ry = 0;
  bool grew = false;
  int rc = find_empty_rect(ctx, w, h, &rx, &ry, &grew);
  if (rc) return rc;
  const float as = (float)ctx->atlas_size;
  fdc_ctx::Rect4 r = {(float)rx / as, (float)ry / as, (float)w / as, (float)h / as};
  ctx->entries[key] = r;
  {
    fdc_ctx::EntryInfo& e = ctx->entry_info[key];  // kind and ids survive a re-put (replaceImageInAtlas keeps the meta)
    e.px = rx; e.py = ry; e.w = w; e.h = h;
    e.order = ++ctx->put_counter;
  }
  ctx->table_dirty = true;
  if (out_rect) { out_rect[0] = r.x; out_rect[1] = r.y; out_rect[2] = r.w; out_rect[3] = r.h; }
  if (out_rebuilt) *out_rebuilt = grew ? 1 : 0;
  return upload_chain(ctx, rx, ry, w, h, rgba);
}

int ensure_rect_image(fdc_ctx* ctx) {
  if (ctx->entries.count(kRectKey)) return FDC_OK;
  uint8_t white[64];
  memset(white, 255, sizeof(white));
  return put_image_impl(ctx, kRectKey, 4, 4, white, nullptr, nullptr);
}

// ---------------------------------------------------------------------------------------------- recording
uint32_t current_xform(fdc_ctx* ctx) {
  if (ctx->xform_dirty || ctx->xforms.n == 0) {
    const float* m = ctx->mat.m;
    Xform x = {m[0], m[4], m[12], m[1], m[5], m[13]};
    ctx->xforms.push(x);
    ctx->xform_dirty = false;
    ctx->state_dirty = true;
  }
  return (uint32_t)ctx->xforms.n - 1;
}

RunState current_state(fdc_ctx* ctx) {
  RunState rs;
  memset(&rs, 0, sizeof(rs));
  rs.xform = current_xform(ctx);
  rs.aa = ctx->aa;
  rs.subpixel_shift = ctx->subpixel_enabled ? fmaxf(0.0f, fminf(ctx->subpixel_shift, 0.999f)) : -1.0f;
  const int L = ctx->mask_write;
  if (ctx->mask_begun) {
    rs.flags = PF_MASK_WRITE | ((uint32_t)L << PF_DEPTH_SHIFT);
    rs.clip_draw = L >= 1 ? ctx->mask_levels[L - 1].clip : -1;
    rs.rectmask = 0;  // setRectMaskVert4 returns early while a mask is being drawn (glcontext.nim:865-866)
  } else {
    rs.flags = (uint32_t)L << PF_DEPTH_SHIFT;
    rs.clip_draw = ctx->mask_levels[L].clip;
    rs.rectmask = 0;
    for (int i = (int)ctx->rm_stack.size() - 1; i >= 0; i--)
      if (ctx->rm_stack[i].fast) { rs.rectmask = ctx->rm_stack[i].index; break; }
  }
  return rs;
}

// Stages one draw record; it gets global draw index ctx->n_draws.
bool stage_draw(fdc_ctx* ctx, const fdc_call& d) {
  const size_t off = ctx->draws.n;
  if (!ctx->draws.push(d)) return false;
  if (!ctx->uploads.empty() && ctx->uploads.back().dst + ctx->uploads.back().count == ctx->n_draws &&
      ctx->uploads.back().staged_off + ctx->uploads.back().count == off)
    ctx->uploads.back().count++;
  else
    ctx->uploads.push_back({ctx->n_draws, 1u, off});
  ctx->n_draws++;
  return true;
}

// The early-outs the reference takes before a quad is emitted (glcontext.nim:1463-1464, :1631-1632, :1305-1310).  The
// device drops such records anyway; between beginMask and endMask the host must not count them as mask draws either.
bool draws_nothing(fdc_ctx* ctx, const fdc_call& d) {
  switch (d.op) {
    case FDC_OP_ROUNDED_RECT: return d.f[2] <= 0.0f || d.f[3] <= 0.0f;
    case FDC_OP_BEZIER: return d.f[2] <= 0.0f || d.f[3] <= 0.0f || d.f[10] <= 0.0f;
    case FDC_OP_IMAGE:
    case FDC_OP_MSDF: return ctx->entries.count((uint64_t)d.u[0] | ((uint64_t)d.u[1] << 32)) == 0;
    default: return false;
  }
}

// Appends one draw record under the current backend state.
int add_draw(fdc_ctx* ctx, const fdc_call& d, uint32_t ordinal) {
  if (!ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "draw outside beginFrame/endFrame");
  if (ctx->mask_begun && draws_nothing(ctx, d)) return FDC_OK;
  const uint32_t idx = ctx->n_draws;
  if (ctx->xform_dirty) current_xform(ctx);
  const bool consecutive = ctx->runs.n > 0 && ordinal == ctx->last_draw_ordinal + 1;
  if (ctx->state_dirty || ctx->begin_pending || !consecutive || ctx->runs.n == 0) {
    RunState rs = current_state(ctx);
    rs.first_draw = idx;
    rs.call_index = ordinal;
    if (ctx->begin_pending) rs.flags |= PF_MASK_BEGIN;
    if (!ctx->runs.push(rs)) return ctx->fail(FDC_ERR_CUDA, "out of pinned memory");
    ctx->state_dirty = ctx->begin_pending;  // the draw after a MASK_BEGIN draw needs its own run
    ctx->begin_pending = false;
  }
  if (!stage_draw(ctx, d)) return ctx->fail(FDC_ERR_CUDA, "out of pinned memory");
  ctx->last_draw_ordinal = ordinal;
  ctx->segments.back().count++;
  if (ctx->mask_begun) {
    MaskLevel& ml = ctx->mask_levels[ctx->mask_write];
    ml.draws.push_back(d);
    ml.states.push_back(ctx->runs.p[ctx->runs.n - 1]);
    if (ml.draws.size() == 1) ml.first_run = (int32_t)ctx->runs.n - 1;
    if (ml.draws.size() == 1 && d.op == FDC_OP_ROUNDED_RECT) {
      ml.mask_draw = (int32_t)idx;
    } else {
      // The level cannot be used as a clip box (several mask draws, or a shape that is not a rounded rect): content under
      // it is binned everywhere, so the level must read 0 wherever no mask draw lands -- GL cleared the whole mask
      // texture at beginMask (glcontext.nim:1901-1902).  The first draw carries that clear over the parent's clip box.
      ml.mask_draw = -1;
      if (ml.first_run >= 0 && !(ctx->runs.p[ml.first_run].flags & PF_MASK_WIDE)) {
        ctx->runs.p[ml.first_run].flags |= PF_MASK_WIDE;
        ml.states[0].flags |= PF_MASK_WIDE;
      }
    }
  }
  return FDC_OK;
}

void fill_rect_radii(fdc_call& c, const float rect[4], const float rx[4], const float ry[4]) {
  memcpy(&c.f[0], rect, 16);
  memcpy(&c.f[4], rx, 16);
  memcpy(&c.f[8], ry, 16);
}

int begin_mask_impl(fdc_ctx* ctx, const float rect[4], const float rx[4], const float ry[4], uint32_t ordinal) {
  if (!ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "ctx.beginFrame has not been called.");
  if (ctx->mask_begun) return ctx->fail(FDC_ERR_STATE, "ctx.beginMask has already been called.");
  if (ctx->mask_write + 1 > kMaxMaskDepth)
    return ctx->fail(FDC_ERR_CAPACITY, "clip masks nest deeper than %d levels", kMaxMaskDepth);
  ctx->mask_begun = true;
  ctx->mask_write++;
  MaskLevel& ml = ctx->mask_levels[ctx->mask_write];
  ml.draws.clear();
  ml.states.clear();
  ml.mask_draw = -1;
  ml.first_run = -1;
  ml.clip = ctx->mask_levels[ctx->mask_write - 1].clip;
  ctx->state_dirty = true;
  ctx->begin_pending = true;
  // drawRoundedRectSdf(clipRect, rgba(255,0,0,255), radii, sdfModeClipAA, 4, 0)   glcontext.nim:1906-1914
  fdc_call c;
  memset(&c, 0, sizeof(c));
  c.op = FDC_OP_ROUNDED_RECT;
  fill_rect_radii(c, rect, rx, ry);
  c.f[12] = 4.0f;
  c.u[0] = FDC_SDF_CLIP_AA;
  c.u[1] = FDC_FILL_COLOR;
  c.u[3] = 0xFF0000FFu;
  // A zero-sized clip rect is dropped (early-out, glcontext.nim:1463-1464) but the level was still cleared: the next mask
  // draw, if any, carries the clear; a level that ends without a draw clips everything away (end_mask_impl).
  return add_draw(ctx, c, ordinal);
}

int end_mask_impl(fdc_ctx* ctx) {
  if (!ctx->mask_begun) return ctx->fail(FDC_ERR_STATE, "ctx.maskBegun has not been called.");
  ctx->mask_begun = false;
  ctx->begin_pending = false;
  MaskLevel& ml = ctx->mask_levels[ctx->mask_write];
  if (ml.draws.empty()) ml.clip = -2;  // cleared and never drawn: "clip to nothing"
  else if (ml.mask_draw >= 0) ml.clip = ml.mask_draw;
  ctx->state_dirty = true;
  return FDC_OK;
}

int pop_mask_impl(fdc_ctx* ctx) {
  if (ctx->mask_write <= 0) return ctx->fail(FDC_ERR_STATE, "popMask without beginMask");
  if (ctx->mask_begun) return ctx->fail(FDC_ERR_STATE, "popMask inside beginMask/endMask");
  ctx->mask_write--;
  ctx->state_dirty = true;
  return FDC_OK;
}

// Conservative pixel bbox of a transformed rect (superset of the ceil'd quad), clipped to the frame.
void host_bbox(fdc_ctx* ctx, const float rect[4], int& x0, int& y0, int& x1, int& y1) {
  const float* m = ctx->mat.m;
  float xs[4] = {rect[0], rect[0] + rect[2], rect[0] + rect[2], rect[0]};
  float ys[4] = {rect[1], rect[1], rect[1] + rect[3], rect[1] + rect[3]};
  float mnx = 1e30f, mny = 1e30f, mxx = -1e30f, mxy = -1e30f;
  for (int k = 0; k < 4; k++) {
    float x = m[0] * xs[k] + m[4] * ys[k] + m[12], y = m[1] * xs[k] + m[5] * ys[k] + m[13];
    mnx = fminf(mnx, x); mxx = fmaxf(mxx, x); mny = fminf(mny, y); mxy = fmaxf(mxy, y);
  }
  x0 = std::max(0, (int)floorf(fmaxf(mnx, -1e6f)) - 1);
  y0 = std::max(0, (int)floorf(fmaxf(mny, -1e6f)) - 1);
  x1 = std::min(ctx->W, (int)ceilf(fminf(mxx, 1e6f)) + 2);
  y1 = std::min(ctx->H, (int)ceilf(fminf(mxy, 1e6f)) + 2);
}

void start_segment(fdc_ctx* ctx) {
  Segment s;
  s.first = ctx->n_draws;
  ctx->segments.push_back(s);
}

// Re-emit the mask primitives of every open texture-mask level at the start of a new segment: the shade kernel
// keeps mask values in registers, so a new launch has to rebuild them (they are pure functions of the pixel).
int reemit_masks(fdc_ctx* ctx) {
  for (int L = 1; L <= ctx->mask_write; L++) {
    MaskLevel& ml = ctx->mask_levels[L];
    for (size_t k = 0; k < ml.draws.size(); k++) {
      RunState rs = ml.states[k];
      rs.first_draw = ctx->n_draws;
      if (k == 0) rs.flags |= PF_MASK_BEGIN;
      if (!ctx->runs.push(rs) || !stage_draw(ctx, ml.draws[k])) return ctx->fail(FDC_ERR_CUDA, "out of pinned memory");
      ctx->segments.back().count++;
    }
    if (ml.draws.empty() && L >= 1) {
      // level cleared but never drawn: nothing to re-emit, content is clipped away anyway
    }
  }
  ctx->state_dirty = true;
  return FDC_OK;
}

// ---------------------------------------------------------------------------------------------- frame execution
cudaEvent_t next_event(fdc_ctx* ctx, int* index) {
  if (ctx->ev_used == ctx->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->ev_pool.push_back(e);
  }
  *index = (int)ctx->ev_used;
  return ctx->ev_pool[ctx->ev_used++];
}

struct Timed {
  fdc_ctx* ctx;
  int a, kind;
  Timed(fdc_ctx* c, int k) : ctx(c), a(-1), kind(k) {
    if (!c->capturing) cudaEventRecord(next_event(c, &a), c->stream);  // (events inside a captured graph cannot be timed)
  }
  ~Timed() {
    if (a < 0) return;
    int b;
    cudaEventRecord(next_event(ctx, &b), ctx->stream);
    ctx->spans.push_back({a, b, kind});
  }
};

int ensure_bin_buffers(fdc_ctx* ctx, uint32_t max_prims) {
  const FrameView& f = ctx->frame;
  const size_t n_bins = (size_t)f.cbx * f.cby;
  const size_t n_chunks = (max_prims + kChunk - 1) / kChunk;
  CK(ctx->d_seg_table.reserve(std::max<size_t>(2, 2 * n_chunks * n_bins)));  // (start, count) per (bin, chunk)
  CK(ctx->d_tile_start.reserve((size_t)f.tiles_x * f.tiles_y));
  CK(ctx->d_tile_count.reserve((size_t)f.tiles_x * f.tiles_y));
  CK(ctx->d_counters.reserve(kNumCounters));
  CK(ctx->d_row_cost.reserve(std::max(1, f.tiles_y)));
  CK(ctx->d_coarse_list.reserve(std::max<size_t>((size_t)max_prims * 3 + n_bins * 4, 1u << 16)));
  CK(ctx->d_tile_list.reserve(std::max<size_t>((size_t)max_prims * 24 + (size_t)f.tiles_x * f.tiles_y * 2, 1u << 20)));
  return FDC_OK;
}

BinBuffers bin_buffers(fdc_ctx* ctx) {
  BinBuffers b;
  b.seg_table = ctx->d_seg_table.p;
  b.coarse_list = ctx->d_coarse_list.p;
  b.coarse_cap = (uint32_t)std::min<size_t>(ctx->d_coarse_list.cap, 0xFFFFFFF0u);
  if (ctx->dbg_coarse_limit) b.coarse_cap = std::min(b.coarse_cap, ctx->dbg_coarse_limit);
  b.tile_start = ctx->d_tile_start.p;
  b.tile_count = ctx->d_tile_count.p;
  b.tile_list = ctx->d_tile_list.p;
  b.tile_cap = (uint32_t)std::min<size_t>(ctx->d_tile_list.cap, 0xFFFFFFF0u);
  if (ctx->dbg_tile_limit) b.tile_cap = std::min(b.tile_cap, ctx->dbg_tile_limit);
  b.counters = ctx->d_counters.p;
  b.row_cost = ctx->d_row_cost.p;
  return b;
}

SetupArgs setup_args(fdc_ctx* ctx, const Segment& s) {
  SetupArgs a;
  a.draws = ctx->d_draws.p;
  a.rects64 = ctx->rec_shared_frame ? reinterpret_cast<const fdc_rect64*>(ctx->fb() + ctx->rec_off) : ctx->d_rects64.p;
  a.runs = ctx->d_runs.p;
  a.n_runs = (int)ctx->runs.n;
  a.xforms = ctx->d_xforms.p;
  a.first = s.first;
  a.count = s.count;
  a.prims = ctx->d_prims.p + s.first;
  a.prim_bins = ctx->d_prim_bins.p + s.first;
  a.geoms = ctx->d_geoms.p + s.first;
  a.exts = ctx->d_exts.p + s.first;
  a.prim_call = ctx->d_prim_call.p + s.first;
  a.atlas = atlas_view(ctx);
  a.frame = ctx->frame;
  a.counters = ctx->d_counters.p;
  a.zero_counters = (int)kCntStickyOverflow;
  a.row_cost = ctx->d_row_cost.p;
  a.n_row_cost = ctx->frame.tiles_y;
  return a;
}

// First tile row that rank r does NOT own (equal bands unless the host set its own).
int band_end_tile_row(const fdc_ctx* ctx, int r) {
  const int tiles_y = ctx->frame.tiles_y;
  if ((int)ctx->band_bounds.size() == ctx->n_ranks + 1 && ctx->band_bounds.back() == tiles_y) return ctx->band_bounds[r + 1];
  const int per = (tiles_y + ctx->n_ranks - 1) / ctx->n_ranks;
  return std::min((r + 1) * per, tiles_y);
}

int ensure_copy_streams(fdc_ctx* ctx) {
  if (ctx->copy_streams[0]) return FDC_OK;
  for (auto& cs : ctx->copy_streams) CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  for (auto& e : ctx->ev_sub) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : ctx->ev_copy) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return FDC_OK;
}

// Launches every kernel of the recorded frame.  `upload`: copy the recording to the device first.  `retry`: the frame is
// being re-run because a bin list overflowed -- pixels a first attempt may already have blended are restored first.
int execute_frame(fdc_ctx* ctx, bool upload, bool retry = false) {
  cudaStream_t st = ctx->stream;
  const uint32_t n_draws = ctx->n_draws;
  int rc = sync_table(ctx);
  if (rc) return rc;
  if (!ctx->capturing) {
    ctx->ev_used = 0;
    ctx->spans.clear();
    cudaEventRecord(ctx->ev_begin, st);
  }
  int launches = 0;
  if (upload) {
    CK(ctx->d_draws.reserve_keep(std::max<uint32_t>(n_draws, 1), n_draws, st));
    CK(ctx->d_runs.reserve(std::max<size_t>(ctx->runs.n, 1)));
    CK(ctx->d_xforms.reserve(std::max<size_t>(ctx->xforms.n, 1)));
    CK(ctx->d_rectmasks.reserve(std::max<size_t>(ctx->rectmasks.n, 1)));
    CK(ctx->d_prims.reserve(std::max<uint32_t>(n_draws, 1)));
    CK(ctx->d_prim_bins.reserve(std::max<uint32_t>(n_draws, 1)));
    CK(ctx->d_geoms.reserve(std::max<uint32_t>(n_draws, 1)));
    CK(ctx->d_exts.reserve(std::max<uint32_t>(n_draws, 1)));
    CK(ctx->d_prim_call.reserve(std::max<uint32_t>(n_draws, 1)));
    for (auto& u : ctx->uploads)
      CK(cudaMemcpyAsync(ctx->d_draws.p + u.dst, ctx->draws.p + u.staged_off, sizeof(fdc_call) * u.count, cudaMemcpyHostToDevice, st));
    ctx->uploads.clear();
    if (ctx->runs.n) CK(cudaMemcpyAsync(ctx->d_runs.p, ctx->runs.p, sizeof(RunState) * ctx->runs.n, cudaMemcpyHostToDevice, st));
    if (ctx->xforms.n) CK(cudaMemcpyAsync(ctx->d_xforms.p, ctx->xforms.p, sizeof(Xform) * ctx->xforms.n, cudaMemcpyHostToDevice, st));
    if (ctx->rectmasks.n)
      CK(cudaMemcpyAsync(ctx->d_rectmasks.p, ctx->rectmasks.p, sizeof(RectMaskRec) * ctx->rectmasks.n, cudaMemcpyHostToDevice, st));
  }
  const bool exchange = upload && !ctx->exchanges.empty();
  uint32_t max_prims = 0;
  for (auto& s : ctx->segments) max_prims = std::max(max_prims, s.count);
  rc = ensure_bin_buffers(ctx, max_prims);
  if (rc) return rc;
  const size_t fb_bytes = (size_t)ctx->W * ctx->H * 4;
  if (ctx->flag_off && fb_bytes > ctx->flag_off)
    return ctx->fail(FDC_ERR_CAPACITY, "frame larger than the framebuffer reserved with fdc_reserve_framebuffer / fdc_bind_shared_framebuffer");
  if (ctx->exported.va && ctx->ext_fb == (uint8_t*)ctx->exported.va && fb_bytes > ctx->exported.size)
    return ctx->fail(FDC_ERR_CAPACITY, "frame larger than the framebuffer exported with fdc_export_framebuffer");
  if (!ctx->ext_fb) {
    const size_t had = ctx->d_fb.cap;
    CK(ctx->d_fb.reserve(fb_bytes));
    // a new framebuffer starts transparent black: a first frame without clearMain blends over defined pixels
    if (ctx->d_fb.cap != had) CK(cudaMemsetAsync(ctx->d_fb.p, 0, ctx->d_fb.cap, st));
  }
  bool any_blur = false;
  for (auto& s : ctx->segments) any_blur = any_blur || s.has_blur;
  if (any_blur) {
    CK(ctx->d_backdrop.reserve(fb_bytes));
    CK(ctx->d_temp.reserve(fb_bytes));
  }
  // Per-frame counters (sticky overflow flags, list-size maxima, entry total): zeroed once here; launch_binning resets
  // only the per-segment words, so an overflow in ANY segment is still visible when the host looks after the frame.
  // (a frame whose first segment has primitives lets that segment's setup kernel zero all the words instead)
  const bool setup_zeroes_all = !ctx->segments.empty() && ctx->segments[0].count > 0;
  if (!setup_zeroes_all) {
    CK(cudaMemsetAsync(ctx->d_counters.p + kCntStickyOverflow, 0, sizeof(uint32_t) * (kNumCounters - kCntStickyOverflow), st));
    if (ctx->d_row_cost.p) CK(cudaMemsetAsync(ctx->d_row_cost.p, 0, sizeof(uint32_t) * (size_t)std::max(1, ctx->frame.tiles_y), st));
  }
  ctx->frame_resolved = false;
  // A frame that blends over the previous pixels (no clearMain) in several segments cannot simply be re-run after an
  // overflow in a later segment -- the earlier segments would be composited twice.  Keep the pre-frame pixels.
  if (!ctx->clear && ctx->segments.size() > 1) {
    if (retry && ctx->snapshot_valid) {
      CK(cudaMemcpyAsync(ctx->fb(), ctx->d_snapshot.p, fb_bytes, cudaMemcpyDeviceToDevice, st));
    } else if (!retry) {
      CK(ctx->d_snapshot.reserve(fb_bytes));
      CK(cudaMemcpyAsync(ctx->d_snapshot.p, ctx->fb(), fb_bytes, cudaMemcpyDeviceToDevice, st));
      ctx->snapshot_valid = true;
    }
  } else if (!retry) {
    ctx->snapshot_valid = false;
  }
  ctx->frame_barrier_base = ctx->barrier_seq;
  const bool banded_blur = ctx->n_ranks > 1 && any_blur;
  const bool end_barrier = ctx->n_ranks > 1 && ctx->frame_barrier && ctx->n_peers == ctx->n_ranks && ctx->flag_off != 0;
  uint32_t* flag_ptrs[kMaxRanks] = {};
  if (banded_blur || end_barrier || exchange) {
    if (ctx->n_peers != ctx->n_ranks || ctx->flag_off == 0)
      return ctx->fail(FDC_ERR_STATE, "backdrop blur under a tile-band partition needs a framebuffer the peers can reach: "
                                      "fdc_reserve_framebuffer + fdc_set_peer_framebuffers, or fdc_bind_shared_framebuffer");
    for (int r = 0; r < ctx->n_ranks; r++) {
      uint8_t* base = (r == ctx->rank || !ctx->h_peers[r]) ? ctx->fb() : ctx->h_peers[r];
      flag_ptrs[r] = reinterpret_cast<uint32_t*>(base + ctx->flag_off);
    }
  }
  if (exchange) {
    // Sharded upload: push the record slices this rank brought over PCIe into every rank's copy, then wait until all
    // ranks have done the same.  (The peers' setup kernels of the PREVIOUS frame are long done: every rank passed that
    // frame's end barrier after its own shade.)
    for (auto& x : ctx->exchanges) {
      const size_t off = ctx->rec_off + (size_t)x.first * sizeof(fdc_rect64);
      launch_push_to_peers(ctx->fb() + off, off, (size_t)x.count * sizeof(fdc_rect64), ctx->mc_fb, ctx->d_peers.p, ctx->n_peers, ctx->rank, st);
      launches++;
    }
    ctx->barrier_seq++;
    launch_signal_flags(flag_ptrs, ctx->n_ranks, ctx->rank, st);
    launch_wait_flags(flag_ptrs[ctx->rank], ctx->n_ranks, st);
    launches += 2;
  }
  uint32_t pending_wait = 0;  // "neighbours finished reading my halo rows" value to wait for before the next shade
  for (size_t si = 0; si < ctx->segments.size(); si++) {
    const Segment& s = ctx->segments[si];
    {
      Timed t(ctx, 0);
      SetupArgs sargs = setup_args(ctx, s);
      if (si == 0 && setup_zeroes_all) sargs.zero_counters = (int)kNumCounters;
      launch_prim_setup(sargs, st);
      launches += s.count ? 1 : 0;
      launch_binning(ctx->d_prim_bins.p + s.first, s.count, ctx->frame, bin_buffers(ctx), st, &launches);
    }
    {
      Timed t(ctx, 1);
      ShadeArgs sa;
      memset(&sa, 0, sizeof(sa));
      sa.prims = ctx->d_prims.p + s.first;
      sa.geoms = ctx->d_geoms.p + s.first;
      sa.exts = ctx->d_exts.p + s.first;
      sa.rectmasks = ctx->d_rectmasks.p;
      sa.tile_start = ctx->d_tile_start.p;
      sa.tile_count = ctx->d_tile_count.p;
      sa.tile_list = ctx->d_tile_list.p;
      sa.counters = ctx->d_counters.p;
      sa.fb = ctx->fb();
      sa.backdrop = ctx->d_backdrop.p;
      sa.atlas = atlas_view(ctx);
      sa.frame = ctx->frame;
      sa.load_dst = (si > 0 || !ctx->clear) ? 1 : 0;
      sa.clear_rgba8 = ctx->clear_rgba8;
      const bool last = si + 1 == ctx->segments.size();
      sa.stats = ctx->want_stats ? ctx->d_stats.p : nullptr;
      // the band reaches the other ranks from the last segment's copy-out: one multicast store per chunk when the
      // framebuffer has an NVSwitch multicast mapping, else one store per peer
      sa.multicast = (last && ctx->n_ranks > 1) ? ctx->mc_fb : nullptr;
      sa.peers = (last && ctx->n_peers > 0 && !sa.multicast) ? ctx->d_peers.p : nullptr;
      sa.n_peers = (last && ctx->n_peers > 0 && !sa.multicast) ? ctx->n_peers : 0;
      sa.fence_at_exit = (sa.n_peers > 0 || sa.multicast) && !end_barrier;
      if (pending_wait) {
        launch_wait_flags(flag_ptrs[ctx->rank], ctx->n_ranks, st);
        pending_wait = 0;
        launches++;
      }
      if (last && ctx->n_peers > 0 && ctx->gather_mode == FDC_GATHER_COPY && !ctx->ext_fb && !ctx->mc_fb) {
        // Copy-engine gather: shade the band slice by slice; a finished slice travels to every peer over NVLink
        // (cudaMemcpyAsync on side streams) while the SMs shade the next one.
        sa.peers = nullptr;
        sa.n_peers = 0;
        rc = ensure_copy_streams(ctx);
        if (rc) return rc;
        const int rows = ctx->frame.ty1 - ctx->frame.ty0;
        const int n_sub = std::max(1, std::min(std::min(ctx->gather_sub_bands, (int)fdc_ctx::kMaxSubBands), rows));
        for (int sb = 0; sb < n_sub; sb++) {
          const int r0 = ctx->frame.ty0 + rows * sb / n_sub, r1 = ctx->frame.ty0 + rows * (sb + 1) / n_sub;
          if (r1 <= r0) continue;
          ShadeArgs ss = sa;
          ss.frame.ty0 = r0;
          ss.frame.ty1 = r1;
          launch_shade(ss, st);
          launches += 2;
          CK(cudaEventRecord(ctx->ev_sub[sb], st));
          const int y0 = r0 * kTileH, y1 = std::min(r1 * kTileH, ctx->H);
          const size_t off = (size_t)y0 * ctx->W * 4, bytes = (size_t)(y1 - y0) * ctx->W * 4;
          int k = 0;
          for (int r = 0; r < ctx->n_peers; r++) {
            uint8_t* peer = ctx->h_peers[r];
            if (!peer || peer == ctx->d_fb.p || r == ctx->rank) continue;
            cudaStream_t cs = ctx->copy_streams[(k++ + sb) % fdc_ctx::kCopyStreams];
            CK(cudaStreamWaitEvent(cs, ctx->ev_sub[sb], 0));
            CK(cudaMemcpyAsync(peer + off, ctx->d_fb.p + off, bytes, cudaMemcpyDeviceToDevice, cs));
          }
        }
        for (int c = 0; c < fdc_ctx::kCopyStreams; c++) {
          CK(cudaEventRecord(ctx->ev_copy[c], ctx->copy_streams[c]));
          CK(cudaStreamWaitEvent(st, ctx->ev_copy[c], 0));
        }
      } else {
        launch_shade(sa, st);
        launches += 2;  // lean + full kernel
      }
    }
    if (s.has_blur) {
      Timed t(ctx, 2);
      BlurArgs ba;
      memset(&ba, 0, sizeof(ba));
      int by0 = s.ry0, by1 = s.ry1;
      uint32_t v_done = 0;
      if (banded_blur) {
        // Halo exchange over peer memory: (1) everyone has finished shading this segment, (2) the H pass reads the rows
        // it needs straight out of the owners' framebuffers, (3) tell everyone the halo has been read so the next
        // segment may overwrite those rows.  Every rank runs the same barrier sequence, with or without work.
        ctx->barrier_seq += 2;
        v_done = 1;
        launch_signal_flags(flag_ptrs, ctx->n_ranks, ctx->rank, st);
        launch_wait_flags(flag_ptrs[ctx->rank], ctx->n_ranks, st);
        launches += 2;
        by0 = std::max(by0, ctx->frame.band_y0);
        by1 = std::min(by1, ctx->frame.band_y1);
        ba.n_src = ctx->n_ranks;
        for (int r = 0; r < ctx->n_ranks; r++) ba.band_end_px[r] = band_end_tile_row(ctx, r) * kTileH;
        for (int r = 0; r < ctx->n_ranks; r++)
          ba.src_rank[r] = (r == ctx->rank || !ctx->h_peers[r]) ? ctx->fb() : ctx->h_peers[r];
      }
      ba.src = ctx->fb();
      ba.temp = ctx->d_temp.p;
      ba.dst = ctx->d_backdrop.p;
      ba.W = ctx->W; ba.H = ctx->H;
      ba.x0 = s.rx0; ba.y0 = by0; ba.x1 = s.rx1; ba.y1 = by1;
      ba.radius = s.blur_radius;
      launch_backdrop_blur(ba, st, &launches);
      if (banded_blur) {
        // (the H pass is the only reader of remote rows and precedes this signal in stream order)
        launch_signal_flags(flag_ptrs, ctx->n_ranks, ctx->rank, st);
        launches++;
        pending_wait = v_done;
      }
    }
  }
  if (pending_wait) {  // do not let the next frame's shade overwrite rows a neighbour may still be reading
    launch_wait_flags(flag_ptrs[ctx->rank], ctx->n_ranks, st);
    launches++;
  }
  if (end_barrier) {
    // Fused gather: once every rank has passed this barrier, every rank's framebuffer holds the whole frame.
    ctx->barrier_seq++;
    launch_signal_flags(flag_ptrs, ctx->n_ranks, ctx->rank, st);
    launch_wait_flags(flag_ptrs[ctx->rank], ctx->n_ranks, st);
    launches += 2;
  }
  if (!ctx->capturing) cudaEventRecord(ctx->ev_end, st);
  CK(cudaGetLastError());
  ctx->stats.n_prims = n_draws;
  ctx->stats.n_segments = (uint32_t)ctx->segments.size();
  ctx->stats.tiles_x = ctx->frame.tiles_x;
  ctx->stats.tiles_y = ctx->frame.tiles_y;
  ctx->stats.tile_w = kTileW;
  ctx->stats.tile_h = kTileH;
  ctx->stats.n_launches = launches;
  ctx->have_frame = true;
  return FDC_OK;
}

void drop_graph(fdc_ctx* ctx) {
  if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
  ctx->graph_exec = nullptr;
}

// After a frame: wait, and if a bin list overflowed in any segment grow the lists and re-run the frame.  Under a
// tile-band partition the re-run cannot be private to this rank (peers gathered or read rows of the aborted frame):
// the lists are regrown and FDC_ERR_RETRY tells the host to call fdc_retry_frame on EVERY rank.
int resolve_frame(fdc_ctx* ctx) {
  if (!ctx->have_frame) return FDC_OK;
  for (int attempt = 0; attempt < 5; attempt++) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->frame_resolved) return FDC_OK;
    uint32_t c[kNumCounters] = {};
    if (!ctx->d_counters.p) return FDC_OK;
    CK(cudaMemcpy(c, ctx->d_counters.p, sizeof(c), cudaMemcpyDeviceToHost));
    ctx->stats.n_tile_entries = c[kCntSumEntries];
    if (ctx->n_ranks > 1 && ctx->flag_off && ctx->barrier_seq != ctx->frame_barrier_base) {
      uint32_t late = 0;
      uint32_t* err = reinterpret_cast<uint32_t*>(ctx->fb() + ctx->flag_off) + kFlagError;
      CK(cudaMemcpy(&late, err, 4, cudaMemcpyDeviceToHost));
      if (late) {
        CK(cudaMemset(err, 0, 4));
        ctx->have_frame = false;
        return ctx->fail(FDC_ERR_STATE, "blur halo barrier timed out waiting for rank mask 0x%x (did every rank submit the frame?)", late);
      }
    }
    if (c[kCntStickyOverflow] == 0) {
      ctx->frame_resolved = true;
      return FDC_OK;
    }
    // overflow in some segment: kCntMaxCoarse / kCntMaxTile hold the largest coarse / tile list any segment needs
    if (attempt == 4) break;
    drop_graph(ctx);  // the lists are about to move
    ctx->dbg_coarse_limit = ctx->dbg_tile_limit = 0;
    if (c[kCntStickyOverflow] & 1u) CK(ctx->d_coarse_list.reserve((size_t)c[kCntMaxCoarse] + (c[kCntMaxCoarse] >> 2) + 1024));
    if (c[kCntStickyOverflow] & 2u) CK(ctx->d_tile_list.reserve((size_t)c[kCntMaxTile] + (c[kCntMaxTile] >> 2) + 1024));
    ctx->n_replays++;
    if (ctx->n_ranks > 1) {
      ctx->frame_resolved = true;  // checked; the verdict is "retry on every rank"
      return ctx->fail(FDC_ERR_RETRY, "a bin list overflowed on rank %d; the lists were regrown -- call fdc_retry_frame on every rank", ctx->rank);
    }
    int rc = execute_frame(ctx, false, true);
    if (rc) return rc;
  }
  return ctx->fail(FDC_ERR_CAPACITY, "bin lists still overflow after regrowing");
}

void compute_frame_view(fdc_ctx* ctx) {
  FrameView& f = ctx->frame;
  f.W = ctx->W; f.H = ctx->H;
  f.tiles_x = (ctx->W + kTileW - 1) / kTileW;
  f.tiles_y = (ctx->H + kTileH - 1) / kTileH;
  if ((int)ctx->band_bounds.size() == ctx->n_ranks + 1 && ctx->band_bounds.back() == f.tiles_y) {
    f.ty0 = ctx->band_bounds[ctx->rank];  // bands chosen by the host (fdc_set_band_tile_rows)
    f.ty1 = ctx->band_bounds[ctx->rank + 1];
  } else {
    const int per = (f.tiles_y + ctx->n_ranks - 1) / ctx->n_ranks;
    f.ty0 = std::min(ctx->rank * per, f.tiles_y);
    f.ty1 = std::min(f.ty0 + per, f.tiles_y);
  }
  f.cty0 = (f.ty0 / kCoarse) * kCoarse;
  f.cbx = (f.tiles_x + kCoarse - 1) / kCoarse;
  f.cby = f.ty1 > f.ty0 ? (f.ty1 - f.cty0 + kCoarse - 1) / kCoarse : 0;
  f.band_y0 = f.ty0 * kTileH;
  f.band_y1 = std::min(f.ty1 * kTileH, ctx->H);
}

uint8_t quant8(float x) {
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  return (uint8_t)floorf(x * 255.0f + 0.5f);
}

}  // namespace

namespace {
// Driver API entry points through the runtime (the library links cudart statically and must load where no driver is
// installed, e.g. for the ABI checks on a CPU-only machine).
template <typename F>
bool driver_fn(const char* name, F* out) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p || q != cudaDriverEntryPointSuccess) return false;
  *out = reinterpret_cast<F>(p);
  return true;
}
void release_export(fdc_ctx* ctx) {
  auto& e = ctx->exported;
  if (!e.va) return;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*addr_free)(CUdeviceptr, size_t) = nullptr;
  if (driver_fn("cuMemUnmap", &unmap) && driver_fn("cuMemRelease", &release) && driver_fn("cuMemAddressFree", &addr_free)) {
    unmap((CUdeviceptr)e.va, e.size);
    release((CUmemGenericAllocationHandle)e.handle);
    addr_free((CUdeviceptr)e.va, e.size);
  }
  if (e.fd >= 0) close(e.fd);
  if (ctx->ext_fb == (uint8_t*)e.va) ctx->ext_fb = nullptr;
  e = {};
}
}  // namespace

// ================================================================================================= C ABI
extern "C" {

int fdc_abi_version(void) { return FDC_ABI_VERSION; }

const char* fdc_last_error(fdc_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int fdc_create(fdc_ctx** out, int device, int atlas_size, float pixel_scale, int rank, int n_ranks) {
  if (!out) return FDC_ERR_INVALID;
  *out = nullptr;
  if (atlas_size < 16 || atlas_size > 16384 || n_ranks < 1 || rank < 0 || rank >= n_ranks) {
    g_create_error = "fdc_create: bad arguments";
    return FDC_ERR_INVALID;
  }
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0 || device < 0 || device >= n_dev) {
    g_create_error = std::string("fdc_create: no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "bad ordinal") +
                     "); there is no CPU fallback";
    return FDC_ERR_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_error = "fdc_create: kernels are built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
    return FDC_ERR_CUDA;
  }
  fdc_ctx* ctx = new fdc_ctx();
  ctx->device = device;
  ctx->pixel_scale = pixel_scale;
  ctx->rank = rank;
  ctx->n_ranks = n_ranks;
  ctx->initial_atlas_size = atlas_size;
  auto bail = [&](cudaError_t err, const char* what) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
    delete ctx;
    return FDC_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
  if ((e = cudaEventCreate(&ctx->ev_begin)) != cudaSuccess) return bail(e, "cudaEventCreate");
  if ((e = cudaEventCreate(&ctx->ev_end)) != cudaSuccess) return bail(e, "cudaEventCreate");
  ctx->mask_levels[0].clip = -1;
  int rc = atlas_alloc(ctx, atlas_size);
  if (rc) {
    g_create_error = ctx->error;
    delete ctx;
    return rc;
  }
  *out = ctx;
  return FDC_OK;
}

void fdc_destroy(fdc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  drop_graph(ctx);
  release_export(ctx);
  for (int l = 0; l < ctx->n_levels; l++) cudaFree(ctx->levels[l]);
  ctx->d_table.release();
  ctx->draws.release(); ctx->runs.release(); ctx->xforms.release(); ctx->rectmasks.release();
  ctx->flat[0].release(); ctx->flat[1].release();
  ctx->d_rects64.release();
  ctx->d_draws.release(); ctx->d_runs.release(); ctx->d_xforms.release(); ctx->d_rectmasks.release();
  ctx->d_prims.release(); ctx->d_prim_bins.release(); ctx->d_geoms.release(); ctx->d_exts.release(); ctx->d_prim_call.release();
  ctx->d_seg_table.release(); ctx->d_coarse_list.release();
  ctx->d_tile_start.release(); ctx->d_tile_count.release(); ctx->d_tile_list.release(); ctx->d_counters.release();
  ctx->d_row_cost.release();
  ctx->d_fb.release(); ctx->d_backdrop.release(); ctx->d_temp.release(); ctx->d_peers.release(); ctx->d_snapshot.release();
  for (auto e : ctx->ev_pool) cudaEventDestroy(e);
  for (auto& cs : ctx->copy_streams) if (cs) cudaStreamDestroy(cs);
  for (auto& e : ctx->ev_sub) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->ev_copy) if (e) cudaEventDestroy(e);
  if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
  if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

// ------------------------------------------------------------------------------------------------- frame
static int sync_frame(fdc_ctx* ctx);

int fdc_begin_frame(fdc_ctx* ctx, int width, int height, int clear_main, const float clear_rgba[4]) {
  if (!ctx) return FDC_ERR_INVALID;
  if (ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "ctx.beginFrame has already been called.");
  if (width <= 0 || height <= 0 || width > 32640 || height > 32640) return ctx->fail(FDC_ERR_INVALID, "bad frame size %dx%d", width, height);
  CK(cudaSetDevice(ctx->device));
  int rc = sync_frame(ctx);  // previous frame (and its read-back) must be complete before its recording is dropped
  if (rc) return rc;
  ctx->have_frame = false;
  drop_graph(ctx);
  if (ctx->ext_fb == nullptr && (width != ctx->W || height != ctx->H)) {
    // new size: the internal framebuffer starts black/transparent like a fresh GL back buffer
    CK(ctx->d_fb.reserve((size_t)width * height * 4));
    CK(cudaMemsetAsync(ctx->d_fb.p, 0, (size_t)width * height * 4, ctx->stream));
  }
  ctx->W = width; ctx->H = height;
  ctx->clear = clear_main != 0;
  if (clear_main && clear_rgba) {
    ctx->clear_rgba8 = (uint32_t)quant8(clear_rgba[0]) | ((uint32_t)quant8(clear_rgba[1]) << 8) | ((uint32_t)quant8(clear_rgba[2]) << 16) |
                       ((uint32_t)quant8(clear_rgba[3]) << 24);
  }
  compute_frame_view(ctx);
  ctx->draws.n = ctx->runs.n = ctx->xforms.n = ctx->rectmasks.n = 0;
  ctx->n_draws = 0;
  ctx->n_rects64 = 0;
  ctx->rec_shared_frame = false;
  ctx->exchanges.clear();
  ctx->uploads.clear();
  ctx->segments.clear();
  start_segment(ctx);
  ctx->frame_begun = true;
  ctx->mask_begun = false;
  ctx->mask_write = 0;
  ctx->mask_levels[0].clip = -1;
  ctx->rm_stack.clear();  // beginFrameProj glcontext.nim:1955
  ctx->state_dirty = ctx->xform_dirty = true;
  ctx->begin_pending = false;
  ctx->call_ordinal = 0;
  ctx->last_draw_ordinal = 0;
  return FDC_OK;
}

int fdc_end_frame(fdc_ctx* ctx) {
  if (!ctx) return FDC_ERR_INVALID;
  if (!ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "ctx.beginFrame was not called first.");
  if (ctx->mask_write != 0) return ctx->fail(FDC_ERR_STATE, "Not all masks have been popped.");
  if (!ctx->rm_stack.empty()) return ctx->fail(FDC_ERR_STATE, "Not all rect masks have been popped.");
  ctx->frame_begun = false;
  CK(cudaSetDevice(ctx->device));
  return execute_frame(ctx, true);
}

// Replays the resident frame.  The first replay of a frame captures its launches (memsets, setup, binning, shade,
// blur, cross-rank barriers) into a CUDA graph; later replays are ONE cudaGraphLaunch -- no per-kernel launch latency,
// which is most of a small frame (cfg1: 8 primitives).  Anything that changes what the launches look like (a new
// recording, regrown lists, another framebuffer) drops the graph.
int fdc_replay_frame(fdc_ctx* ctx) {
  if (!ctx) return FDC_ERR_INVALID;
  if (!ctx->have_frame || ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "no completed frame to replay");
  CK(cudaSetDevice(ctx->device));
  const bool side_streams = ctx->n_peers > 0 && ctx->gather_mode == FDC_GATHER_COPY;
  if (!ctx->graph_enabled || ctx->want_stats || side_streams) return execute_frame(ctx, false);
  int rc = FDC_OK;
  if (ctx->table_dirty) drop_graph(ctx);  // atlas entries (or the atlas itself) changed since the capture
  if (!ctx->graph_exec) {
    rc = resolve_frame(ctx);  // capture against lists known to be large enough (an overflow regrows them and drops the graph)
    if (rc) return rc;
    rc = sync_table(ctx);
    if (rc) return rc;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    ctx->capturing = true;
    rc = execute_frame(ctx, false);
    ctx->capturing = false;
    const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc != FDC_OK || e != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      if (rc != FDC_OK) return rc;
      return execute_frame(ctx, false);  // capture refused (e.g. an allocation happened): plain launches
    }
    const cudaError_t ei = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) {
      ctx->graph_exec = nullptr;
      cudaGetLastError();
      return execute_frame(ctx, false);
    }
  }
  ctx->ev_used = 0;
  ctx->spans.clear();  // per-phase times are not available from inside a graph; gpu_ms is
  CK(cudaEventRecord(ctx->ev_begin, ctx->stream));
  CK(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
  CK(cudaEventRecord(ctx->ev_end, ctx->stream));
  ctx->frame_resolved = false;
  return FDC_OK;
}

// Graph replay on / off (default on).  Off: fdc_replay_frame re-issues the launches one by one, which also times the
// phases (bin_ms / shade_ms / blur_ms of fdc_get_frame_stats).
int fdc_set_replay_graph(fdc_ctx* ctx, int enabled) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->graph_enabled = enabled != 0;
  if (!enabled) drop_graph(ctx);
  return FDC_OK;
}

// Re-runs the last frame on every rank of a tile-band partition after any rank's fdc_sync / fdc_read_pixels returned
// FDC_ERR_RETRY (the host all-reduces the status).  Unlike fdc_replay_frame it restores pixels a first attempt already
// blended; every rank runs the same cross-rank barrier sequence again, so blur halo rows are consistent.
int fdc_retry_frame(fdc_ctx* ctx) {
  if (!ctx) return FDC_ERR_INVALID;
  if (!ctx->have_frame || ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "no completed frame to retry");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return execute_frame(ctx, false, true);
}

// Drops a frame that was begun but cannot be ended (a call in between failed): the recording, the mask / rect-mask
// stacks and the transform stack are reset, the previous frame's pixels stay.  beginFrame may be called again.
int fdc_abort_frame(fdc_ctx* ctx) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->frame_begun = false;
  ctx->mask_begun = false;
  ctx->begin_pending = false;
  ctx->mask_write = 0;
  ctx->rm_stack.clear();
  ctx->mats.clear();
  ctx->mat = mat_identity();
  ctx->draws.n = ctx->runs.n = ctx->xforms.n = ctx->rectmasks.n = 0;
  ctx->n_draws = 0;
  ctx->n_rects64 = 0;
  ctx->uploads.clear();
  ctx->segments.clear();
  ctx->state_dirty = ctx->xform_dirty = true;
  ctx->have_frame = false;  // the resident recording was (partly) overwritten by direct uploads of the aborted frame
  return FDC_OK;
}

int fdc_debug_limit_lists(fdc_ctx* ctx, uint32_t coarse_entries, uint32_t tile_entries) {
  if (!ctx) return FDC_ERR_INVALID;
  drop_graph(ctx);
  ctx->dbg_coarse_limit = coarse_entries;
  ctx->dbg_tile_limit = tile_entries;
  return FDC_OK;
}

static int enqueue_readback(fdc_ctx* ctx) {
  auto& r = ctx->pending_read;
  CK(cudaMemcpy2DAsync(r.out, (size_t)r.w * 4, ctx->fb() + ((size_t)r.y * ctx->W + r.x) * 4, (size_t)ctx->W * 4, (size_t)r.w * 4,
                       (size_t)r.h, cudaMemcpyDeviceToHost, ctx->stream));
  return FDC_OK;
}

// resolve_frame + completion of an asynchronous read-back
static int sync_frame(fdc_ctx* ctx) {
  const uint32_t replays_before = ctx->n_replays;
  int rc = resolve_frame(ctx);
  if (rc == FDC_OK && ctx->pending_read.out) {
    if (ctx->n_replays != replays_before) {  // the frame was re-run after a list regrow: the copy read the aborted one
      rc = enqueue_readback(ctx);
      if (rc == FDC_OK) CK(cudaStreamSynchronize(ctx->stream));
    }
  }
  ctx->pending_read.out = nullptr;
  return rc;
}

int fdc_sync(fdc_ctx* ctx) {
  if (!ctx) return FDC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  return sync_frame(ctx);
}

// Asynchronous readPixels: the device-to-host copy is queued behind the frame on fdc_stream and this call returns;
// `out_rgba` (pinned host memory for a truly asynchronous copy) is valid after the next fdc_sync.
int fdc_read_pixels_async(fdc_ctx* ctx, int x, int y, int w, int h, uint8_t* out_rgba) {
  if (!ctx || !out_rgba) return FDC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->have_frame || !ctx->fb() || ctx->W <= 0) return ctx->fail(FDC_ERR_STATE, "no frame has been rendered");
  if (w <= 0 || h <= 0) { x = 0; y = 0; w = ctx->W; h = ctx->H; }
  if (x < 0 || y < 0 || x + w > ctx->W || y + h > ctx->H) return ctx->fail(FDC_ERR_INVALID, "readPixels rect outside the frame");
  ctx->pending_read.out = out_rgba;
  ctx->pending_read.x = x; ctx->pending_read.y = y; ctx->pending_read.w = w; ctx->pending_read.h = h;
  return enqueue_readback(ctx);
}

int fdc_read_pixels(fdc_ctx* ctx, int x, int y, int w, int h, uint8_t* out_rgba) {
  if (!ctx || !out_rgba) return FDC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  int rc = resolve_frame(ctx);
  if (rc) return rc;
  if (!ctx->fb() || ctx->W <= 0) return ctx->fail(FDC_ERR_STATE, "no frame has been rendered");
  if (w <= 0 || h <= 0) { x = 0; y = 0; w = ctx->W; h = ctx->H; }
  if (x < 0 || y < 0 || x + w > ctx->W || y + h > ctx->H) return ctx->fail(FDC_ERR_INVALID, "readPixels rect outside the frame");
  CK(cudaMemcpy2DAsync(out_rgba, (size_t)w * 4, ctx->fb() + ((size_t)y * ctx->W + x) * 4, (size_t)ctx->W * 4, (size_t)w * 4, (size_t)h,
                       cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return FDC_OK;
}

// ------------------------------------------------------------------------------------------------- transforms
static void mark_xform(fdc_ctx* ctx) { ctx->xform_dirty = true; }

int fdc_translate(fdc_ctx* ctx, float x, float y) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->call_ordinal++;
  Mat4 t = mat_identity();
  t.m[12] = x; t.m[13] = y;
  ctx->mat = mat_mul(ctx->mat, t);
  mark_xform(ctx);
  return FDC_OK;
}
int fdc_rotate(fdc_ctx* ctx, float angle) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->call_ordinal++;
  Mat4 r = mat_identity();  // vmath rotateZ: +angle turns +x toward -y on screen
  const float cs = cosf(angle), sn = sinf(angle);
  r.m[0] = cs; r.m[1] = -sn; r.m[4] = sn; r.m[5] = cs;
  ctx->mat = mat_mul(ctx->mat, r);
  mark_xform(ctx);
  return FDC_OK;
}
int fdc_scale(fdc_ctx* ctx, float sx, float sy) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->call_ordinal++;
  Mat4 s = mat_identity();
  s.m[0] = sx; s.m[5] = sy;
  ctx->mat = mat_mul(ctx->mat, s);
  mark_xform(ctx);
  return FDC_OK;
}
int fdc_apply_transform(fdc_ctx* ctx, const float mat4[16]) {
  if (!ctx || !mat4) return FDC_ERR_INVALID;
  ctx->call_ordinal++;
  Mat4 m;
  memcpy(m.m, mat4, 64);
  ctx->mat = mat_mul(ctx->mat, m);
  mark_xform(ctx);
  return FDC_OK;
}
int fdc_save_transform(fdc_ctx* ctx) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->call_ordinal++;
  ctx->mats.push_back(ctx->mat);
  return FDC_OK;
}
int fdc_restore_transform(fdc_ctx* ctx) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->call_ordinal++;
  if (ctx->mats.empty()) return ctx->fail(FDC_ERR_STATE, "restoreTransform on an empty stack");
  ctx->mat = ctx->mats.back();
  ctx->mats.pop_back();
  mark_xform(ctx);
  return FDC_OK;
}
int fdc_transform_mirrors_y(fdc_ctx* ctx) {
  if (!ctx) return 0;
  const float* m = ctx->mat.m;  // determinant of the 2x2 linear part (glcontext.nim:2019-2024)
  return (m[0] * m[5] - m[1] * m[4]) < 0.0f ? 1 : 0;
}
int fdc_get_transform(fdc_ctx* ctx, float out_mat4[16]) {
  if (!ctx || !out_mat4) return FDC_ERR_INVALID;
  memcpy(out_mat4, ctx->mat.m, 64);
  return FDC_OK;
}

// ------------------------------------------------------------------------------------------------- AA / text flags
float fdc_sdf_aa_factor(fdc_ctx* ctx) { return ctx ? ctx->aa : 0.0f; }
int fdc_set_sdf_aa_factor(fdc_ctx* ctx, float aa) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->call_ordinal++;
  if (ctx->aa == aa) return FDC_OK;
  ctx->aa = aa;
  ctx->state_dirty = true;
  return FDC_OK;
}
int fdc_set_text_subpixel_positioning_enabled(fdc_ctx* ctx, int enabled) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->subpixel_enabled = enabled != 0;
  ctx->state_dirty = true;
  return FDC_OK;
}
int fdc_set_text_subpixel_shift(fdc_ctx* ctx, float shift) {
  if (!ctx) return FDC_ERR_INVALID;
  if (ctx->subpixel_shift != shift) ctx->state_dirty = true;
  ctx->subpixel_shift = shift;
  return FDC_OK;
}
float fdc_pixel_scale(fdc_ctx* ctx) { return ctx ? ctx->pixel_scale : 1.0f; }
// `pixelate` of newContext (glcontext.nim:255-282): the atlas texture's magnification filter becomes GL_NEAREST
// (:165-168); minified sampling stays trilinear.  (Mask and backdrop textures are only ever sampled at texel centres.)
int fdc_set_pixelate(fdc_ctx* ctx, int enabled) {
  if (!ctx) return FDC_ERR_INVALID;
  if (ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "cannot change the atlas filter inside a frame");
  ctx->pixelate = enabled != 0;
  drop_graph(ctx);
  return FDC_OK;
}

// ------------------------------------------------------------------------------------------------- draws
static void put_fill(fdc_call& c, const fdc_fill* fill) {
  c.u[1] = fill->kind;
  c.u[2] = fill->axis;
  memcpy(&c.u[3], fill->c, 16);
  c.f[16] = fill->mid_pos;
}

int fdc_draw_rounded_rect_sdf(fdc_ctx* ctx, const float rect[4], const fdc_fill* fill, const float radii_x[4], const float radii_y[4],
                              int mode, float factor, float spread, const float shape_size[2]) {
  if (!ctx || !rect || !fill || !radii_x || !radii_y) return FDC_ERR_INVALID;
  const uint32_t ord = ctx->call_ordinal++;
  if (rect[2] <= 0.0f || rect[3] <= 0.0f) return FDC_OK;  // glcontext.nim:1463-1464
  fdc_call c;
  memset(&c, 0, sizeof(c));
  c.op = FDC_OP_ROUNDED_RECT;
  fill_rect_radii(c, rect, radii_x, radii_y);
  c.f[12] = factor; c.f[13] = spread;
  if (shape_size) { c.f[14] = shape_size[0]; c.f[15] = shape_size[1]; }
  c.u[0] = (uint32_t)mode;
  put_fill(c, fill);
  return add_draw(ctx, c, ord);
}

int fdc_draw_image(fdc_ctx* ctx, uint64_t key, const float pos[2], const uint32_t colors[4], const float size[2], int flip_y) {
  if (!ctx || !pos || !colors) return FDC_ERR_INVALID;
  const uint32_t ord = ctx->call_ordinal++;
  if (!ctx->entries.count(key)) return ctx->fail(FDC_ERR_MISSING_IMAGE, "missing image in context");
  fdc_call c;
  memset(&c, 0, sizeof(c));
  c.op = FDC_OP_IMAGE;
  c.u[0] = (uint32_t)key; c.u[1] = (uint32_t)(key >> 32);
  memcpy(&c.u[3], colors, 16);
  c.u[7] = flip_y ? 1u : 0u;
  c.f[0] = pos[0]; c.f[1] = pos[1];
  if (size) { c.f[2] = size[0]; c.f[3] = size[1]; }
  return add_draw(ctx, c, ord);
}

int fdc_draw_msdf_image(fdc_ctx* ctx, uint64_t key, const float pos[2], uint32_t color, const float size[2], float px_range,
                        float sd_threshold, float stroke_weight, int flip_y, int is_mtsdf) {
  if (!ctx || !pos || !size) return FDC_ERR_INVALID;
  const uint32_t ord = ctx->call_ordinal++;
  if (!ctx->entries.count(key)) return ctx->fail(FDC_ERR_MISSING_IMAGE, "missing image in context");
  fdc_call c;
  memset(&c, 0, sizeof(c));
  c.op = FDC_OP_MSDF;
  c.u[0] = (uint32_t)key; c.u[1] = (uint32_t)(key >> 32);
  c.u[2] = is_mtsdf ? 1u : 0u;
  c.u[3] = color;
  c.u[7] = flip_y ? 1u : 0u;
  c.f[0] = pos[0]; c.f[1] = pos[1]; c.f[2] = size[0]; c.f[3] = size[1];
  c.f[4] = px_range; c.f[5] = sd_threshold; c.f[6] = stroke_weight;
  return add_draw(ctx, c, ord);
}

int fdc_draw_quadratic_bezier_sdf(fdc_ctx* ctx, const float rect[4], const fdc_fill* fill, const float p0[2], const float p1[2],
                                  const float p2[2], float stroke_weight, int cap) {
  if (!ctx || !rect || !fill || !p0 || !p1 || !p2) return FDC_ERR_INVALID;
  const uint32_t ord = ctx->call_ordinal++;
  if (rect[2] <= 0.0f || rect[3] <= 0.0f || stroke_weight <= 0.0f) return FDC_OK;  // glcontext.nim:1631-1632
  fdc_call c;
  memset(&c, 0, sizeof(c));
  c.op = FDC_OP_BEZIER;
  memcpy(&c.f[0], rect, 16);
  c.f[4] = p0[0]; c.f[5] = p0[1]; c.f[6] = p1[0]; c.f[7] = p1[1]; c.f[8] = p2[0]; c.f[9] = p2[1];
  c.f[10] = stroke_weight;
  c.u[0] = (uint32_t)cap;
  put_fill(c, fill);
  return add_draw(ctx, c, ord);
}

int fdc_draw_filled_quad(fdc_ctx* ctx, const float verts[8], const uint32_t colors[4]) {
  if (!ctx || !verts || !colors) return FDC_ERR_INVALID;
  const uint32_t ord = ctx->call_ordinal++;
  int rc = ensure_rect_image(ctx);
  if (rc) return rc;
  fdc_call c;
  memset(&c, 0, sizeof(c));
  c.op = FDC_OP_FILLED_QUAD;
  memcpy(&c.f[0], verts, 32);
  memcpy(&c.u[3], colors, 16);
  return add_draw(ctx, c, ord);
}

int fdc_draw_rect(fdc_ctx* ctx, const float rect[4], uint32_t color) {
  if (!ctx || !rect) return FDC_ERR_INVALID;
  const uint32_t ord = ctx->call_ordinal++;
  int rc = ensure_rect_image(ctx);
  if (rc) return rc;
  fdc_call c;
  memset(&c, 0, sizeof(c));
  c.op = FDC_OP_RECT;
  memcpy(&c.f[0], rect, 16);
  c.u[3] = color;
  return add_draw(ctx, c, ord);
}

int fdc_draw_backdrop_blur(fdc_ctx* ctx, const float rect[4], const float radii_x[4], const float radii_y[4], float blur_radius) {
  if (!ctx || !rect || !radii_x || !radii_y) return FDC_ERR_INVALID;
  const uint32_t ord = ctx->call_ordinal++;
  if (blur_radius <= 0.0f || rect[2] <= 0.0f || rect[3] <= 0.0f) return FDC_OK;  // glcontext.nim:1791-1792
  if (!ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "draw outside beginFrame/endFrame");
  if (ctx->mask_begun) return ctx->fail(FDC_ERR_STATE, "drawBackdropBlur inside beginMask/endMask is not supported");
  if (ctx->n_ranks > 1 && (ctx->n_peers != ctx->n_ranks || ctx->flag_off == 0))
    return ctx->fail(FDC_ERR_STATE, "backdrop blur under a tile-band partition reads halo rows from the neighbours' framebuffers: "
                                    "call fdc_reserve_framebuffer and fdc_set_peer_framebuffers, or fdc_bind_shared_framebuffer (all ranks) first");
  Segment& s = ctx->segments.back();
  s.has_blur = true;
  s.blur_radius = blur_radius;
  host_bbox(ctx, rect, s.rx0, s.ry0, s.rx1, s.ry1);
  start_segment(ctx);
  int rc = reemit_masks(ctx);
  if (rc) return rc;
  // composite: drawRoundedRectSdf(rect, whiteColor, radii, sdfModeBackdropBlur, factor = blurRadius)  glcontext.nim:1833-1841
  fdc_call c;
  memset(&c, 0, sizeof(c));
  c.op = FDC_OP_ROUNDED_RECT;
  fill_rect_radii(c, rect, radii_x, radii_y);
  c.f[12] = blur_radius;
  c.u[0] = FDC_SDF_BACKDROP_BLUR;
  c.u[1] = FDC_FILL_COLOR;
  c.u[3] = 0xFFFFFFFFu;
  ctx->state_dirty = true;
  return add_draw(ctx, c, ord);
}

// ------------------------------------------------------------------------------------------------- masks
int fdc_begin_mask(fdc_ctx* ctx, const float rect[4], const float radii_x[4], const float radii_y[4]) {
  if (!ctx || !rect || !radii_x || !radii_y) return FDC_ERR_INVALID;
  const uint32_t ord = ctx->call_ordinal++;
  return begin_mask_impl(ctx, rect, radii_x, radii_y, ord);
}
int fdc_end_mask(fdc_ctx* ctx) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->call_ordinal++;
  return end_mask_impl(ctx);
}
int fdc_pop_mask(fdc_ctx* ctx) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->call_ordinal++;
  return pop_mask_impl(ctx);
}
int fdc_begin_rect_mask(fdc_ctx* ctx, const float rect[4], const float radii_x[4], const float radii_y[4]) {
  if (!ctx || !rect || !radii_x || !radii_y) return FDC_ERR_INVALID;
  const uint32_t ord = ctx->call_ordinal++;
  if (!ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "ctx.beginFrame has not been called.");
  if (ctx->mask_begun) return ctx->fail(FDC_ERR_STATE, "ctx.beginRectMask cannot start inside a mask.");
  if (ctx->rm_stack.empty() && rect[2] > 0.0f && rect[3] > 0.0f) {
    // makeRectMask glcontext.nim:831-850
    RectMaskRec rm;
    memset(&rm, 0, sizeof(rm));
    const float hx = rect[2] * 0.5f, hy = rect[3] * 0.5f;
    Mat4 inv;
    if (!mat_inverse(ctx->mat, inv)) inv = mat_identity();
    float rr[4];
    const bool ell = rounded_radii_vec_h(radii_x, radii_y, hx, hy, rr);
    rm.cx = rect[0] + hx; rm.cy = rect[1] + hy; rm.hx = hx; rm.hy = hy;
    rm.r0 = rr[0]; rm.r1 = rr[1]; rm.r2 = rr[2]; rm.r3 = rr[3];
    rm.ax = inv.m[0]; rm.ay = inv.m[4]; rm.az = inv.m[12];
    rm.bx = inv.m[1]; rm.by = inv.m[5]; rm.bz = inv.m[13];
    rm.elliptical = ell ? 1.0f : 0.0f;
    if (!ctx->rectmasks.push(rm)) return ctx->fail(FDC_ERR_CUDA, "out of pinned memory");
    ctx->rm_stack.push_back({true, (uint32_t)ctx->rectmasks.n});
    ctx->state_dirty = true;
    return FDC_OK;
  }
  int rc = begin_mask_impl(ctx, rect, radii_x, radii_y, ord);
  if (rc) return rc;
  rc = end_mask_impl(ctx);
  if (rc) return rc;
  ctx->rm_stack.push_back({false, 0});
  return FDC_OK;
}
int fdc_pop_rect_mask(fdc_ctx* ctx) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->call_ordinal++;
  if (ctx->rm_stack.empty()) return ctx->fail(FDC_ERR_STATE, "No rect mask has been pushed.");
  RectMaskEntry e = ctx->rm_stack.back();
  ctx->rm_stack.pop_back();
  ctx->state_dirty = true;
  if (!e.fast) return pop_mask_impl(ctx);
  return FDC_OK;
}

// ------------------------------------------------------------------------------------------------- display list
static inline bool is_draw_op(uint32_t op) { return op >= FDC_OP_ROUNDED_RECT && op <= FDC_OP_RECT; }
constexpr size_t kDirectRunMin = 2048;  // records; shorter runs are staged
constexpr size_t kShardRunMin = 1024;   // records per rank below which a run of compact records is uploaded whole by every rank

// A long run of draw records issued under one backend state: one RunState, and the records go to the device in a
// single copy straight from the caller's buffer (no host-side staging pass over 128 bytes per draw).
static int add_direct_run(fdc_ctx* ctx, const fdc_call* calls, size_t count, uint32_t first_ordinal) {
  RunState rs = current_state(ctx);
  rs.first_draw = ctx->n_draws;
  rs.call_index = first_ordinal;
  if (!ctx->runs.push(rs)) return ctx->fail(FDC_ERR_CUDA, "out of pinned memory");
  CK(cudaSetDevice(ctx->device));
  CK(ctx->d_draws.reserve_keep((size_t)ctx->n_draws + count, ctx->n_draws, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_draws.p + ctx->n_draws, calls, sizeof(fdc_call) * count, cudaMemcpyHostToDevice, ctx->stream));
  ctx->n_draws += (uint32_t)count;
  ctx->segments.back().count += (uint32_t)count;
  ctx->last_draw_ordinal = first_ordinal + (uint32_t)count - 1;
  ctx->state_dirty = true;  // the next draw starts its own run
  return FDC_OK;
}

int fdc_submit_calls(fdc_ctx* ctx, const fdc_call* calls, size_t n) {
  if (!ctx || (!calls && n)) return FDC_ERR_INVALID;
  size_t short_until = 0;  // records before this index belong to a draw run already found too short
  for (size_t i = 0; i < n; i++) {
    const fdc_call& c = calls[i];
    int rc = FDC_OK;
    if (i >= short_until && is_draw_op(c.op) && ctx->frame_begun && !ctx->mask_begun && n - i >= kDirectRunMin) {
      size_t j = i;
      bool need_rect = false;
      while (j < n && is_draw_op(calls[j].op)) {
        need_rect = need_rect || calls[j].op == FDC_OP_FILLED_QUAD || calls[j].op == FDC_OP_RECT;
        j++;
      }
      if (j - i >= kDirectRunMin) {
        if (need_rect && (rc = ensure_rect_image(ctx)) != FDC_OK) return rc;
        rc = add_direct_run(ctx, calls + i, j - i, ctx->call_ordinal);
        if (rc != FDC_OK) return rc;
        ctx->call_ordinal += (uint32_t)(j - i);
        i = j - 1;
        continue;
      }
      short_until = j;
    }
    switch (c.op) {
      case FDC_OP_NOP: ctx->call_ordinal++; break;
      case FDC_OP_SAVE_TRANSFORM: rc = fdc_save_transform(ctx); break;
      case FDC_OP_RESTORE_TRANSFORM: rc = fdc_restore_transform(ctx); break;
      case FDC_OP_TRANSLATE: rc = fdc_translate(ctx, c.f[0], c.f[1]); break;
      case FDC_OP_ROTATE: rc = fdc_rotate(ctx, c.f[0]); break;
      case FDC_OP_SCALE: rc = fdc_scale(ctx, c.f[0], c.f[1]); break;
      case FDC_OP_APPLY_TRANSFORM: rc = fdc_apply_transform(ctx, c.f); break;
      case FDC_OP_SET_AA: rc = fdc_set_sdf_aa_factor(ctx, c.f[0]); break;
      case FDC_OP_SET_SUBPIXEL:
        ctx->call_ordinal++;
        fdc_set_text_subpixel_positioning_enabled(ctx, (int)c.u[0]);
        fdc_set_text_subpixel_shift(ctx, c.f[0]);
        break;
      case FDC_OP_BEGIN_MASK: rc = fdc_begin_mask(ctx, &c.f[0], &c.f[4], &c.f[8]); break;
      case FDC_OP_END_MASK: rc = fdc_end_mask(ctx); break;
      case FDC_OP_POP_MASK: rc = fdc_pop_mask(ctx); break;
      case FDC_OP_BEGIN_RECT_MASK: rc = fdc_begin_rect_mask(ctx, &c.f[0], &c.f[4], &c.f[8]); break;
      case FDC_OP_POP_RECT_MASK: rc = fdc_pop_rect_mask(ctx); break;
      case FDC_OP_BACKDROP_BLUR: rc = fdc_draw_backdrop_blur(ctx, &c.f[0], &c.f[4], &c.f[8], c.f[12]); break;
      case FDC_OP_FILLED_QUAD:
      case FDC_OP_RECT:
        rc = ensure_rect_image(ctx);
        if (rc) break;
        // fallthrough
      case FDC_OP_ROUNDED_RECT:
      case FDC_OP_IMAGE:
      case FDC_OP_MSDF:
      case FDC_OP_BEZIER: {
        // Draw records are taken verbatim; early-outs and atlas lookups happen in prim_setup_kernel.
        const uint32_t ord = ctx->call_ordinal++;
        rc = add_draw(ctx, c, ord);
        break;
      }
      default: rc = ctx->fail(FDC_ERR_INVALID, "fdc_submit_calls: unknown op %u at record %zu", c.op, i); break;
    }
    if (rc != FDC_OK) return rc;
  }
  return FDC_OK;
}

int fdc_submit_draws(fdc_ctx* ctx, const fdc_call* draws, size_t n) {
  if (!ctx || (!draws && n)) return FDC_ERR_INVALID;
  if (n == 0) return FDC_OK;
  if (!ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "draw outside beginFrame/endFrame");
  if (ctx->mask_begun || n < kDirectRunMin) return fdc_submit_calls(ctx, draws, n);  // small or mask content: per record
  int rc = ensure_rect_image(ctx);  // FILLED_QUAD / RECT records may be inside; the 4x4 white image is cheap
  if (rc) return rc;
  rc = add_direct_run(ctx, draws, n, ctx->call_ordinal);
  if (rc == FDC_OK) ctx->call_ordinal += (uint32_t)n;
  return rc;
}

// fdc_rect64 helpers (pure host) and the compact bulk path.
int fdc_pack_rect64(const fdc_call* in, fdc_rect64* out) {
  if (!in || !out || in->op != FDC_OP_ROUNDED_RECT) return 0;
  const uint32_t mode = in->u[0], kind = in->u[1], axis = in->u[2];
  if (mode > 255u || kind < (uint32_t)FDC_FILL_COLOR || kind > (uint32_t)FDC_FILL_LINEAR3 || axis > 3u) return 0;
  if (memcmp(&in->f[4], &in->f[8], 4 * sizeof(float)) != 0) return 0;  // elliptical corners need the full record
  fdc_rect64 r;
  memset(&r, 0, sizeof(r));
  for (int k = 0; k < 4; k++) { r.rect[k] = in->f[k]; r.radii[k] = in->f[4 + k]; }
  r.factor = in->f[12]; r.spread = in->f[13];
  r.shape_size[0] = in->f[14]; r.shape_size[1] = in->f[15];
  r.c[0] = in->u[3]; r.c[1] = in->u[4]; r.c[2] = in->u[5];
  uint32_t m = 0;
  if (kind == (uint32_t)FDC_FILL_LINEAR3) {
    const float mid = in->f[16];
    if (!(mid >= 0.01f && mid <= 0.99f)) return 0;
    m = (uint32_t)lrintf(mid * 255.0f);
    if (m > 255u) return 0;
  }
  r.packed = mode | (kind << 8) | (axis << 10) | (m << 16);
  // accept only what expands back to the very same 128 bytes (tries the neighbouring uint8 mid positions too)
  for (int dm = 0; dm < 3; dm++) {
    const int mm = (int)m + (dm == 0 ? 0 : (dm == 1 ? -1 : 1));
    if (mm < 0 || mm > 255) continue;
    r.packed = mode | (kind << 8) | (axis << 10) | ((uint32_t)mm << 16);
    fdc_call back;
    expand_rect64_words(r, reinterpret_cast<uint32_t*>(&back));
    if (memcmp(&back, in, sizeof(fdc_call)) == 0) {
      *out = r;
      return 1;
    }
    if (kind != (uint32_t)FDC_FILL_LINEAR3) break;
  }
  return 0;
}

void fdc_expand_rect64(const fdc_rect64* in, fdc_call* out) {
  if (in && out) expand_rect64_words(*in, reinterpret_cast<uint32_t*>(out));
}

int fdc_submit_rects64(fdc_ctx* ctx, const fdc_rect64* rects, size_t n) {
  if (!ctx || (!rects && n)) return FDC_ERR_INVALID;
  if (n == 0) return FDC_OK;
  if (!ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "draw outside beginFrame/endFrame");
  if (ctx->mask_begun) return ctx->fail(FDC_ERR_STATE, "fdc_submit_rects64 inside beginMask/endMask is not supported");
  RunState rs = current_state(ctx);
  rs.first_draw = ctx->n_draws;
  rs.call_index = ctx->call_ordinal;
  rs.compact = 1;
  rs.src_off = ctx->n_rects64;
  if (!ctx->runs.push(rs)) return ctx->fail(FDC_ERR_CUDA, "out of pinned memory");
  CK(cudaSetDevice(ctx->device));
  // the draw index space stays one: d_draws keeps (unused) room for these indices
  CK(ctx->d_draws.reserve_keep((size_t)ctx->n_draws + n, ctx->n_draws, ctx->stream));
  const bool fits_shared = ctx->n_ranks > 1 && ctx->ext_fb && ctx->frame_barrier && ctx->n_peers == ctx->n_ranks &&
                           ctx->rec_bytes >= ((size_t)ctx->n_rects64 + n) * sizeof(fdc_rect64);
  if (ctx->n_rects64 == 0) ctx->rec_shared_frame = fits_shared;  // every rank sees the same stream: same decision
  if (ctx->rec_shared_frame) {
    if (!fits_shared)
      return ctx->fail(FDC_ERR_CAPACITY, "record exchange area of the shared framebuffer is too small for this frame (%zu bytes)", ctx->rec_bytes);
    fdc_rect64* base = reinterpret_cast<fdc_rect64*>(ctx->fb() + ctx->rec_off) + ctx->n_rects64;
    if (n >= kShardRunMin * (size_t)ctx->n_ranks) {
      const size_t lo = n * (size_t)ctx->rank / ctx->n_ranks, hi = n * (size_t)(ctx->rank + 1) / ctx->n_ranks;
      CK(cudaMemcpyAsync(base + lo, rects + lo, sizeof(fdc_rect64) * (hi - lo), cudaMemcpyHostToDevice, ctx->stream));
      ctx->exchanges.push_back({ctx->n_rects64 + (uint32_t)lo, (uint32_t)(hi - lo)});
    } else {
      CK(cudaMemcpyAsync(base, rects, sizeof(fdc_rect64) * n, cudaMemcpyHostToDevice, ctx->stream));
    }
  } else {
    CK(ctx->d_rects64.reserve_keep((size_t)ctx->n_rects64 + n, ctx->n_rects64, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_rects64.p + ctx->n_rects64, rects, sizeof(fdc_rect64) * n, cudaMemcpyHostToDevice, ctx->stream));
  }
  ctx->n_rects64 += (uint32_t)n;
  ctx->n_draws += (uint32_t)n;
  ctx->segments.back().count += (uint32_t)n;
  ctx->last_draw_ordinal = ctx->call_ordinal + (uint32_t)n - 1;
  ctx->call_ordinal += (uint32_t)n;
  ctx->state_dirty = true;  // the next draw starts its own run
  return FDC_OK;
}

// ------------------------------------------------------------------------------------------------- atlas
int fdc_put_image(fdc_ctx* ctx, uint64_t key, int w, int h, const uint8_t* rgba, float out_rect[4], int* out_rebuilt) {
  if (!ctx) return FDC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  return put_image_impl(ctx, key, w, h, rgba, out_rect, out_rebuilt);
}
int fdc_update_image(fdc_ctx* ctx, uint64_t key, int w, int h, const uint8_t* rgba) {
  if (!ctx || !rgba) return FDC_ERR_INVALID;
  auto it = ctx->entries.find(key);
  if (it == ctx->entries.end()) return ctx->fail(FDC_ERR_MISSING_IMAGE, "updateImage: unknown key");
  const float as = (float)ctx->atlas_size;
  if (it->second.w != (float)w / as || it->second.h != (float)h / as) return ctx->fail(FDC_ERR_INVALID, "updateImage: size differs");
  CK(cudaSetDevice(ctx->device));
  return upload_chain(ctx, (int)(it->second.x * as), (int)(it->second.y * as), w, h, rgba);
}
int fdc_has_image(fdc_ctx* ctx, uint64_t key) { return ctx && ctx->entries.count(key) ? 1 : 0; }
int fdc_get_image_rect(fdc_ctx* ctx, uint64_t key, float out_rect[4]) {
  if (!ctx || !out_rect) return FDC_ERR_INVALID;
  auto it = ctx->entries.find(key);
  if (it == ctx->entries.end()) return ctx->fail(FDC_ERR_MISSING_IMAGE, "unknown image key");
  out_rect[0] = it->second.x; out_rect[1] = it->second.y; out_rect[2] = it->second.w; out_rect[3] = it->second.h;
  return FDC_OK;
}
int fdc_remove_image(fdc_ctx* ctx, uint64_t key) {
  if (!ctx) return FDC_ERR_INVALID;
  if (ctx->entries.erase(key)) ctx->table_dirty = true;  // entries.del(key): the texels stay, as in GL
  ctx->entry_info.erase(key);
  return FDC_OK;
}
int fdc_reset_image_atlas(fdc_ctx* ctx, int minimum_size) {
  if (!ctx) return FDC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  int size = std::max(ctx->initial_atlas_size, 1);  // plannedAtlasSize figbackend.nim:223-227
  const int minimum = std::max(minimum_size, size);
  while (size < minimum) size *= 2;
  ctx->entry_info.clear();
  ctx->atlas_generation++;
  ctx->atlas_rebuilds++;
  return atlas_alloc(ctx, size);
}
// Glyph bitmaps rasterised on the device, straight into their atlas slots (fdc_glyph.cu).
int fdc_rasterize_glyphs(fdc_ctx* ctx, const fdc_glyph_job* jobs, size_t n_jobs, const fdc_outline_seg* segs, size_t n_segs,
                         int lcd_filter, int* out_rebuilt) {
  if (!ctx || (!jobs && n_jobs) || (!segs && n_segs)) return FDC_ERR_INVALID;
  if (out_rebuilt) *out_rebuilt = 0;
  if (n_jobs == 0) return FDC_OK;
  CK(cudaSetDevice(ctx->device));
  struct GlyphDev { uint32_t first_seg, n_segs; int32_t w, h, ax, ay; };
  std::vector<GlyphDev> dev(n_jobs);
  size_t valid_from = 0;  // jobs placed before the last regrow lost their slots (unless the atlas replays itself)
  for (size_t i = 0; i < n_jobs; i++) {
    const fdc_glyph_job& j = jobs[i];
    if (j.width <= 0 || j.height <= 0 || j.width > 4096 || j.height > 4096 || (size_t)j.first_seg + j.n_segs > n_segs)
      return ctx->fail(FDC_ERR_INVALID, "rasterizeGlyphs: bad job %zu (%dx%d, segments %u+%u of %zu)", i, j.width, j.height, j.first_seg,
                       j.n_segs, n_segs);
    int rx = 0, ry = 0;
    bool grew = false;
    int rc = find_empty_rect(ctx, j.width, j.height, &rx, &ry, &grew);
    if (rc) return rc;
    if (grew) {
      if (out_rebuilt) *out_rebuilt = 1;
      if (!ctx->atlas_replay) valid_from = i;
      else
        for (size_t k = 0; k < i; k++) {  // re-packed: the earlier jobs' slots moved
          auto it = ctx->entry_info.find(jobs[k].key);
          if (it != ctx->entry_info.end()) { dev[k].ax = it->second.px; dev[k].ay = it->second.py; }
        }
    }
    const float as = (float)ctx->atlas_size;
    ctx->entries[j.key] = {(float)rx / as, (float)ry / as, (float)j.width / as, (float)j.height / as};
    fdc_ctx::EntryInfo& e = ctx->entry_info[j.key];
    e.px = rx; e.py = ry; e.w = j.width; e.h = j.height;
    e.order = ++ctx->put_counter;
    if (e.kind == FDC_ENTRY_UNKNOWN) e.kind = FDC_ENTRY_GLYPH;
    dev[i] = {j.first_seg, j.n_segs, j.width, j.height, rx, ry};
  }
  for (size_t k = 0; k < valid_from; k++) dev[k].w = dev[k].h = 0;
  ctx->table_dirty = true;
  void* d_jobs = nullptr;
  fdc_outline_seg* d_segs = nullptr;
  CK(cudaMalloc(&d_jobs, n_jobs * sizeof(GlyphDev)));
  CK(cudaMalloc(&d_segs, std::max<size_t>(n_segs, 1) * sizeof(fdc_outline_seg)));
  CK(cudaMemcpyAsync(d_jobs, dev.data(), n_jobs * sizeof(GlyphDev), cudaMemcpyHostToDevice, ctx->stream));
  if (n_segs) CK(cudaMemcpyAsync(d_segs, segs, n_segs * sizeof(fdc_outline_seg), cudaMemcpyHostToDevice, ctx->stream));
  launch_glyph_raster(d_jobs, (int)n_jobs, d_segs, ctx->levels[0], ctx->atlas_size, lcd_filter, ctx->stream);
  CK(cudaGetLastError());
  for (size_t k = valid_from; k < n_jobs; k++) {
    if (dev[k].w > 1 && dev[k].h > 1) {
      int rc = build_mips(ctx, dev[k].ax, dev[k].ay, dev[k].w, dev[k].h);
      if (rc) return rc;
    }
  }
  CK(cudaStreamSynchronize(ctx->stream));  // `dev` and the caller's arrays are pageable
  cudaFree(d_jobs);
  cudaFree(d_segs);
  return FDC_OK;
}

// ---- atlas residency bookkeeping for hosts without the reference's Nim tables
// markImageEntry / markGlyphEntry / markGeneratedEntry, figbackend.nim:359-398
int fdc_mark_entry(fdc_ctx* ctx, uint64_t key, int kind, uint64_t id_a, uint64_t id_b) {
  if (!ctx || kind < FDC_ENTRY_UNKNOWN || kind > FDC_ENTRY_GENERATED) return FDC_ERR_INVALID;
  auto it = ctx->entry_info.find(key);
  if (it == ctx->entry_info.end()) return ctx->fail(FDC_ERR_MISSING_IMAGE, "markEntry: unknown key");
  it->second.kind = kind; it->second.a = id_a; it->second.b = id_b;
  return FDC_OK;
}
static int remove_matching(fdc_ctx* ctx, int kind, bool by_b, uint64_t id) {
  int n = 0;
  for (auto it = ctx->entry_info.begin(); it != ctx->entry_info.end();) {
    if (it->second.kind == kind && (by_b ? it->second.b : it->second.a) == id) {
      ctx->entries.erase(it->first);  // removeAtlasEntry, figbackend.nim:355-357
      it = ctx->entry_info.erase(it);
      n++;
    } else {
      ++it;
    }
  }
  if (n) ctx->table_dirty = true;
  return n;
}
// clearFontGlyphs / clearTypefaceGlyphs, figbackend.nim:416-432; return the number of entries removed
int fdc_clear_font_glyphs(fdc_ctx* ctx, uint64_t font_id) { return ctx ? remove_matching(ctx, FDC_ENTRY_GLYPH, false, font_id) : 0; }
int fdc_clear_typeface_glyphs(fdc_ctx* ctx, uint64_t typeface_id) { return ctx ? remove_matching(ctx, FDC_ENTRY_GLYPH, true, typeface_id) : 0; }
// retainImageOwner / retainFontOwner, figbackend.nim:434-437, :452-455
int fdc_retain_owner(fdc_ctx* ctx, int what, uint64_t id, uint64_t token) {
  if (!ctx || (what != FDC_OWNER_IMAGE && what != FDC_OWNER_FONT)) return FDC_ERR_INVALID;
  auto& owners = (what == FDC_OWNER_IMAGE ? ctx->image_owners : ctx->font_owners)[id];
  if (std::find(owners.begin(), owners.end(), token) == owners.end()) owners.push_back(token);
  return FDC_OK;
}
// releaseImageOwner / releaseFontOwner, figbackend.nim:439-450, :457-468.  *out_last = 1 when that was the last owner;
// the entries it kept alive are then evicted here (what the reference's callers do next: removeImage / clearFontGlyphs).
int fdc_release_owner(fdc_ctx* ctx, int what, uint64_t id, uint64_t token, int* out_last) {
  if (!ctx || (what != FDC_OWNER_IMAGE && what != FDC_OWNER_FONT)) return FDC_ERR_INVALID;
  if (out_last) *out_last = 0;
  auto& table = what == FDC_OWNER_IMAGE ? ctx->image_owners : ctx->font_owners;
  auto it = table.find(id);
  if (it == table.end()) return FDC_OK;
  auto& owners = it->second;
  owners.erase(std::remove(owners.begin(), owners.end(), token), owners.end());
  if (!owners.empty()) return FDC_OK;
  table.erase(it);
  if (out_last) *out_last = 1;
  if (what == FDC_OWNER_IMAGE) remove_matching(ctx, FDC_ENTRY_IMAGE, false, id);
  else remove_matching(ctx, FDC_ENTRY_GLYPH, false, id);
  return FDC_OK;
}
// atlasUsage, figbackend.nim:303-333
int fdc_get_atlas_usage(fdc_ctx* ctx, fdc_atlas_usage* out) {
  if (!ctx || !out) return FDC_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  out->atlas_size = ctx->atlas_size;
  out->generation = ctx->atlas_generation;
  out->rebuild_count = ctx->atlas_rebuilds;
  out->atlas_area = (int64_t)ctx->atlas_size * ctx->atlas_size;
  out->entry_count = (int32_t)ctx->entries.size();
  for (auto& kv : ctx->entries) {
    const int w = std::max(0, (int)lroundf(kv.second.w * (float)ctx->atlas_size)), h = std::max(0, (int)lroundf(kv.second.h * (float)ctx->atlas_size));
    out->used_area += (int64_t)w * h;
    auto it = ctx->entry_info.find(kv.first);
    const int kind = it == ctx->entry_info.end() ? FDC_ENTRY_UNKNOWN : it->second.kind;
    if (kind == FDC_ENTRY_IMAGE) out->image_count++;
    else if (kind == FDC_ENTRY_GLYPH) out->glyph_count++;
    else if (kind == FDC_ENTRY_GENERATED) out->generated_count++;
    else out->unknown_count++;
  }
  out->packed_area = std::max<int64_t>(fdc_atlas_packed_area(ctx), out->used_area);
  out->used_area = std::min(out->used_area, out->atlas_area);
  out->packed_area = std::min(out->packed_area, out->atlas_area);
  return FDC_OK;
}
// Native replay on regrow: when the atlas has to double, the live entries are re-packed and their texels carried over
// on the device (out_rebuilt of fdc_put_image still reports that every rect changed).  Default off = the reference's
// protocol, where the host replays its images itself (noteAtlasRebuilt, figbackend.nim:202-207).
int fdc_set_atlas_replay(fdc_ctx* ctx, int enabled) {
  if (!ctx) return FDC_ERR_INVALID;
  ctx->atlas_replay = enabled != 0;
  return FDC_OK;
}
int fdc_atlas_size(fdc_ctx* ctx) { return ctx ? ctx->atlas_size : 0; }
int fdc_atlas_packed_area(fdc_ctx* ctx) {
  if (!ctx) return 0;
  long long a = 0;
  for (uint16_t h : ctx->heights) a += h;
  return (int)std::min<long long>(a, 0x7FFFFFFF);
}

// ------------------------------------------------------------------------------------------------- plumbing
int fdc_bind_framebuffer(fdc_ctx* ctx, void* device_rgba8) {
  if (!ctx) return FDC_ERR_INVALID;
  if (ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "cannot rebind the framebuffer inside a frame");
  int rc = resolve_frame(ctx);
  if (rc) return rc;
  drop_graph(ctx);
  if (ctx->ext_fb && ctx->frame_barrier) {  // leaving a shared framebuffer: forget its flags, peers and exchange area
    ctx->mc_fb = nullptr;
    ctx->frame_barrier = false;
    ctx->flag_off = ctx->rec_off = ctx->rec_bytes = 0;
    ctx->n_peers = 0;
    ctx->h_peers.clear();
  }
  ctx->ext_fb = (uint8_t*)device_rgba8;
  return FDC_OK;
}
void* fdc_framebuffer_ptr(fdc_ctx* ctx) { return ctx ? ctx->fb() : nullptr; }
// Bands chosen by the host instead of equal ones: n_ranks + 1 tile-row boundaries (16-px rows), bounds[0] = 0,
// non-decreasing, bounds[n_ranks] = ceil(H / 16) of the frames to come.  Takes effect at once for replays of the
// recorded frame and for every later frame of that height; n_bounds = 0 restores equal bands.  Every rank must be given
// the same boundaries.  Needs a framebuffer the ranks share or reach (the NCCL / copy-engine gathers want equal bands).
int fdc_set_band_tile_rows(fdc_ctx* ctx, const int* bounds, int n_bounds) {
  if (!ctx) return FDC_ERR_INVALID;
  if (ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "fdc_set_band_tile_rows inside a frame");
  if (n_bounds == 0) {
    ctx->band_bounds.clear();
  } else {
    if (!bounds || n_bounds != ctx->n_ranks + 1 || bounds[0] != 0) return ctx->fail(FDC_ERR_INVALID, "band boundaries: need n_ranks + 1 values starting at 0");
    for (int r = 0; r < ctx->n_ranks; r++)
      if (bounds[r + 1] < bounds[r]) return ctx->fail(FDC_ERR_INVALID, "band boundaries must not decrease");
    ctx->band_bounds.assign(bounds, bounds + n_bounds);
  }
  CK(cudaSetDevice(ctx->device));
  int rc = resolve_frame(ctx);  // nothing of the old partition may still be in flight
  if (rc != FDC_OK && rc != FDC_ERR_RETRY) return rc;
  drop_graph(ctx);
  if (ctx->W > 0 && ctx->H > 0) compute_frame_view(ctx);
  return FDC_OK;
}

// Tile entries per 16-px tile row, summed over the segments of the last completed frame: this rank's rows only (0
// elsewhere), so a sum over the ranks gives the profile of the whole frame.  A host balances the bands with it.
int fdc_get_tile_row_costs(fdc_ctx* ctx, uint32_t* out, int cap, int* n_rows) {
  if (!ctx) return FDC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  int rc = resolve_frame(ctx);
  if (rc) return rc;
  if (!ctx->have_frame) return ctx->fail(FDC_ERR_STATE, "no completed frame");
  const int n = ctx->frame.tiles_y;
  if (n_rows) *n_rows = n;
  if (!out) return FDC_OK;
  if (cap < n) return ctx->fail(FDC_ERR_INVALID, "fdc_get_tile_row_costs: %d rows, room for %d", n, cap);
  if (n > 0) CK(cudaMemcpy(out, ctx->d_row_cost.p, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
  return FDC_OK;
}

int fdc_band_rows(fdc_ctx* ctx, int* y0, int* y1) {
  if (!ctx || !y0 || !y1) return FDC_ERR_INVALID;
  *y0 = ctx->frame.band_y0;
  *y1 = ctx->frame.band_y1;
  return FDC_OK;
}
void* fdc_stream(fdc_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int fdc_set_peer_framebuffers(fdc_ctx* ctx, void* const* device_ptrs, int n) {
  if (!ctx || n < 0 || (n && !device_ptrs)) return FDC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  int rc = resolve_frame(ctx);
  if (rc) return rc;
  drop_graph(ctx);
  ctx->n_peers = n;
  ctx->h_peers.assign(n, nullptr);
  if (n) {
    if (n > kMaxRanks) return ctx->fail(FDC_ERR_CAPACITY, "at most %d ranks", kMaxRanks);
    for (int r = 0; r < n; r++) ctx->h_peers[r] = (uint8_t*)device_ptrs[r];
    // pointers that live on another device of this process need peer access (IPC mappings already have it)
    for (int r = 0; r < n; r++) {
      if (!device_ptrs[r]) continue;
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, device_ptrs[r]) == cudaSuccess && at.device != ctx->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return ctx->cuda_fail(e, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
      }
    }
    CK(ctx->d_peers.reserve((size_t)n));
    CK(cudaMemcpy(ctx->d_peers.p, device_ptrs, sizeof(void*) * n, cudaMemcpyHostToDevice));
  }
  return FDC_OK;
}

// Present without a read-back (replaces readPixels, glcontext.nim:2094-2135, for a presenter on the same machine): the
// framebuffer becomes a CUDA VMM allocation exported as a POSIX file descriptor.  A Vulkan (VK_KHR_external_memory_fd,
// OPAQUE_FD) or OpenGL (EXT_memory_object_fd) presenter -- or another CUDA process, cuMemImportFromShareableHandle --
// imports it once and samples / blits the RGBA8 rows (pitch width*4, top-left origin) after fdc_sync; no pixel crosses
// PCIe.  The descriptor stays owned by the context (dup() it to keep it beyond fdc_destroy).
int fdc_export_framebuffer(fdc_ctx* ctx, int width, int rows, int* out_fd, size_t* out_bytes) {
  if (!ctx || width <= 0 || rows <= 0 || !out_fd) return FDC_ERR_INVALID;
  if (ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "cannot replace the framebuffer inside a frame");
  CK(cudaSetDevice(ctx->device));
  int rc = resolve_frame(ctx);
  if (rc) return rc;
  drop_graph(ctx);
  CUresult (*get_gran)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*export_fd)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  if (!driver_fn("cuMemGetAllocationGranularity", &get_gran) || !driver_fn("cuMemCreate", &create) ||
      !driver_fn("cuMemAddressReserve", &reserve) || !driver_fn("cuMemMap", &map) || !driver_fn("cuMemSetAccess", &set_access) ||
      !driver_fn("cuMemExportToShareableHandle", &export_fd))
    return ctx->fail(FDC_ERR_CUDA, "the CUDA driver does not provide the virtual memory management API");
  release_export(ctx);
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = ctx->device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  if (get_gran(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0)
    return ctx->fail(FDC_ERR_CUDA, "cuMemGetAllocationGranularity failed");
  const size_t size = (((size_t)width * rows * 4) + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h = 0;
  CUdeviceptr va = 0;
  CUresult r = create(&h, size, &prop, 0);
  if (r != CUDA_SUCCESS) return ctx->fail(FDC_ERR_CUDA, "cuMemCreate(%zu bytes, exportable) failed with %d", size, (int)r);
  r = reserve(&va, size, gran, 0, 0);
  if (r == CUDA_SUCCESS) r = map(va, size, 0, h, 0);
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (r == CUDA_SUCCESS) r = set_access(va, size, &acc, 1);
  int fd = -1;
  if (r == CUDA_SUCCESS) r = export_fd(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) return ctx->fail(FDC_ERR_CUDA, "mapping / exporting the framebuffer failed with %d", (int)r);
  ctx->exported.handle = (unsigned long long)h;
  ctx->exported.va = (unsigned long long)va;
  ctx->exported.size = size;
  ctx->exported.fd = fd;
  CK(cudaMemsetAsync((void*)va, 0, size, ctx->stream));  // a fresh back buffer is transparent black
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->ext_fb = (uint8_t*)va;
  ctx->mc_fb = nullptr;
  ctx->frame_barrier = false;
  ctx->flag_off = ctx->rec_off = ctx->rec_bytes = 0;
  *out_fd = fd;
  if (out_bytes) *out_bytes = size;
  return FDC_OK;
}

// A framebuffer the host allocated so that every rank can reach every copy (CUDA VMM / symmetric memory): this rank's
// mapping, the peers' mappings and -- when the GPUs sit behind an NVSwitch -- the multicast mapping through which one
// store lands in all copies.  `bytes` must cover width * rows * 4 pixels + 4096 bytes of cross-rank flags.
int fdc_bind_shared_framebuffer(fdc_ctx* ctx, void* local_ptr, size_t bytes, void* const* peer_ptrs, int n, void* multicast_ptr,
                                int width, int rows) {
  if (!ctx || !local_ptr || width <= 0 || rows <= 0 || n != ctx->n_ranks || (n && !peer_ptrs)) return FDC_ERR_INVALID;
  if (ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "cannot rebind the framebuffer inside a frame");
  CK(cudaSetDevice(ctx->device));
  int rc = resolve_frame(ctx);
  if (rc) return rc;
  drop_graph(ctx);
  const size_t pix = (((size_t)width * rows * 4) + 255) & ~(size_t)255;
  if (bytes < pix + 4096) return ctx->fail(FDC_ERR_INVALID, "shared framebuffer needs %zu bytes (pixels + 4096 bytes of flags), got %zu", pix + 4096, bytes);
  if (n > kMaxRanks) return ctx->fail(FDC_ERR_CAPACITY, "at most %d ranks", kMaxRanks);
  ctx->ext_fb = (uint8_t*)local_ptr;
  ctx->mc_fb = (uint8_t*)multicast_ptr;
  ctx->flag_off = pix;
  ctx->rec_off = pix + 4096;
  ctx->rec_bytes = bytes - ctx->rec_off;  // whatever the host gave beyond pixels + flags is the record exchange area
  ctx->barrier_seq = 0;
  ctx->frame_barrier = true;
  ctx->n_peers = n;
  ctx->h_peers.assign(n, nullptr);
  for (int r = 0; r < n; r++) ctx->h_peers[r] = (r == ctx->rank) ? (uint8_t*)local_ptr : (uint8_t*)peer_ptrs[r];
  CK(ctx->d_peers.reserve((size_t)n));
  CK(cudaMemcpy(ctx->d_peers.p, ctx->h_peers.data(), sizeof(void*) * n, cudaMemcpyHostToDevice));
  CK(cudaMemsetAsync((uint8_t*)local_ptr + pix, 0, 4096, ctx->stream));  // flags start at 0 on every rank
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->n_ranks > 1) {  // blur scratch up front: no allocation while peers spin on our flags
    CK(ctx->d_backdrop.reserve((size_t)width * rows * 4));
    CK(ctx->d_temp.reserve((size_t)width * rows * 4));
  }
  return FDC_OK;
}

// The cross-rank flag barrier that ends every frame on a shared framebuffer (default on).  A host that synchronises the
// ranks itself, or that sizes several contexts of one process one after the other, can switch it off.
int fdc_set_frame_barrier(fdc_ctx* ctx, int enabled) {
  if (!ctx) return FDC_ERR_INVALID;
  int rc = resolve_frame(ctx);
  if (rc) return rc;
  drop_graph(ctx);
  ctx->frame_barrier = enabled != 0;
  return FDC_OK;
}

int fdc_set_peer_gather(fdc_ctx* ctx, int mode, int sub_bands) {
  if (!ctx || (mode != FDC_GATHER_STORES && mode != FDC_GATHER_COPY)) return FDC_ERR_INVALID;
  int rc = resolve_frame(ctx);
  if (rc) return rc;
  drop_graph(ctx);
  ctx->gather_mode = mode;
  if (sub_bands > 0) ctx->gather_sub_bands = sub_bands;
  return FDC_OK;
}

int fdc_reserve_framebuffer(fdc_ctx* ctx, int width, int rows) {
  if (!ctx || width <= 0 || rows <= 0) return FDC_ERR_INVALID;
  if (ctx->frame_begun) return ctx->fail(FDC_ERR_STATE, "cannot resize the framebuffer inside a frame");
  CK(cudaSetDevice(ctx->device));
  int rc = resolve_frame(ctx);
  if (rc) return rc;
  drop_graph(ctx);
  // pixels, then (256-byte aligned) a small array of cross-rank flags that peers write for the blur halo barrier
  const size_t pix = (((size_t)width * rows * 4) + 255) & ~(size_t)255;
  const size_t bytes = pix + 4096;
  ctx->d_fb.release();
  CK(ctx->d_fb.reserve(bytes));
  CK(cudaMemsetAsync(ctx->d_fb.p, 0, ctx->d_fb.cap, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->flag_off = pix;
  ctx->barrier_seq = 0;
  if (ctx->n_ranks > 1) {  // blur scratch up front: no allocation (an implicit device sync) while peers spin on our flags
    CK(ctx->d_backdrop.reserve((size_t)width * rows * 4));
    CK(ctx->d_temp.reserve((size_t)width * rows * 4));
  }
  return FDC_OK;
}

int fdc_framebuffer_ipc_handle(fdc_ctx* ctx, uint8_t out_handle[64]) {
  if (!ctx || !out_handle) return FDC_ERR_INVALID;
  if (!ctx->d_fb.p) return ctx->fail(FDC_ERR_STATE, "no internal framebuffer yet (fdc_reserve_framebuffer)");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  CK(cudaSetDevice(ctx->device));
  CK(cudaIpcGetMemHandle(&h, ctx->d_fb.p));
  memcpy(out_handle, &h, 64);
  return FDC_OK;
}

int fdc_open_peer_framebuffer(fdc_ctx* ctx, const uint8_t handle[64], void** out_device_ptr) {
  if (!ctx || !handle || !out_device_ptr) return FDC_ERR_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CK(cudaSetDevice(ctx->device));
  CK(cudaIpcOpenMemHandle(out_device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return FDC_OK;
}

int fdc_get_frame_stats(fdc_ctx* ctx, fdc_frame_stats* out) {
  if (!ctx || !out) return FDC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  int rc = resolve_frame(ctx);
  if (rc) return rc;
  if (ctx->have_frame) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end) == cudaSuccess) ctx->stats.gpu_ms = ms;
    if (!ctx->spans.empty()) {  // (a graph replay has no per-phase events: the last launch-by-launch values stay)
      float acc[3] = {0, 0, 0};
      for (auto& sp : ctx->spans)
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[sp.a], ctx->ev_pool[sp.b]) == cudaSuccess) acc[sp.kind] += ms;
      ctx->stats.bin_ms = acc[0]; ctx->stats.shade_ms = acc[1]; ctx->stats.blur_ms = acc[2];
    }
  }
  *out = ctx->stats;
  return FDC_OK;
}

int fdc_debug_shade_stats(fdc_ctx* ctx, uint64_t out[8]) {
  if (!ctx || !out) return FDC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  int rc = resolve_frame(ctx);
  if (rc) return rc;
  if (!ctx->have_frame) return ctx->fail(FDC_ERR_STATE, "no completed frame");
  CK(ctx->d_stats.reserve(8));
  CK(cudaMemsetAsync(ctx->d_stats.p, 0, 64, ctx->stream));
  ctx->want_stats = true;
  rc = execute_frame(ctx, false);
  ctx->want_stats = false;
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(out, ctx->d_stats.p, 64, cudaMemcpyDeviceToHost));
  return FDC_OK;
}

// renderFrame (figrender.nim:1960-2002) with the front-end DFS run natively: beginFrame, the flattened scene, endFrame.
int fdc_render_frame(fdc_ctx* ctx, const fdc_scene* scene, float ui_scale, float frame_w, float frame_h, int clear_main,
                     const float clear_rgba[4]) {
  if (!ctx || !scene || (!scene->lists && scene->n_lists)) return FDC_ERR_INVALID;
  if (!(ui_scale > 0.0f)) return ctx->fail(FDC_ERR_INVALID, "ui_scale must be positive");
  std::vector<uint64_t> keys;
  keys.reserve(ctx->entries.size());
  for (auto& kv : ctx->entries) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  fdc_flatten_env env;
  env.ui_scale = ui_scale;
  env.pixel_scale = ctx->pixel_scale;
  env.aa_factor = ctx->aa;
  env.subpixel_enabled = ctx->subpixel_enabled ? 1u : 0u;
  env.image_keys = keys.data();
  env.n_image_keys = keys.size();
  ctx->flat_idx ^= 1;
  PinnedBuf<fdc_call>& buf = ctx->flat[ctx->flat_idx];
  // the buffer keeps its size: a steady scene flattens in one pass
  if (buf.cap < 1024 && !buf.reserve(1024)) return ctx->fail(FDC_ERR_CUDA, "cudaMallocHost failed");
  size_t n_calls = 0;
  for (int pass = 0; pass < 2; pass++) {
    const char* err = fdc::flatten_renders(*scene, env, buf.p, buf.cap, &n_calls);
    if (err) return ctx->fail(FDC_ERR_INVALID, "%s", err);
    if (n_calls <= buf.cap) break;
    buf.n = 0;  // nothing to keep
    if (!buf.reserve(n_calls + n_calls / 8)) return ctx->fail(FDC_ERR_CUDA, "cudaMallocHost failed");
  }
  int rc = fdc_begin_frame(ctx, (int)(frame_w * ui_scale), (int)(frame_h * ui_scale), clear_main, clear_rgba);
  if (rc) return rc;
  rc = fdc_submit_calls(ctx, buf.p, n_calls);
  if (rc == FDC_OK) rc = fdc_end_frame(ctx);
  if (rc != FDC_OK && ctx->frame_begun) {
    // one bad frame must not wedge the context ("beginFrame has already been called" forever): drop it, keep the message
    const std::string msg = ctx->error;
    fdc_abort_frame(ctx);
    ctx->error = msg;
  }
  return rc;
}

int fdc_debug_bins(fdc_ctx* ctx, int segment, uint32_t* tile_offsets, size_t offsets_cap, uint32_t* entries, size_t entries_cap,
                   size_t* n_offsets, size_t* n_entries) {
  if (!ctx) return FDC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  int rc = resolve_frame(ctx);
  if (rc) return rc;
  if (!ctx->have_frame) return ctx->fail(FDC_ERR_STATE, "no completed frame");
  if (segment < 0 || segment >= (int)ctx->segments.size()) return ctx->fail(FDC_ERR_INVALID, "bad segment index");
  const Segment& s = ctx->segments[segment];
  // Re-run setup + binning for this segment (the lists are reused between segments), then read them back.
  int launches = 0;
  launch_prim_setup(setup_args(ctx, s), ctx->stream);
  launch_binning(ctx->d_prim_bins.p + s.first, s.count, ctx->frame, bin_buffers(ctx), ctx->stream, &launches);
  CK(cudaStreamSynchronize(ctx->stream));
  const size_t n_tiles = (size_t)ctx->frame.tiles_x * ctx->frame.tiles_y;
  std::vector<uint32_t> start(n_tiles, 0), count(n_tiles, 0), calls(std::max<uint32_t>(s.count, 1));
  uint32_t c[4];
  CK(cudaMemcpy(c, ctx->d_counters.p, sizeof(c), cudaMemcpyDeviceToHost));
  if (c[kCntOverflow]) return ctx->fail(FDC_ERR_CAPACITY, "bin lists overflowed during debug readback");
  CK(cudaMemcpy(start.data(), ctx->d_tile_start.p, n_tiles * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(count.data(), ctx->d_tile_count.p, n_tiles * 4, cudaMemcpyDeviceToHost));
  if (s.count) CK(cudaMemcpy(calls.data(), ctx->d_prim_call.p + s.first, (size_t)s.count * 4, cudaMemcpyDeviceToHost));
  std::vector<TileEntry> list(std::max<uint32_t>(c[kCntCursor], 1));
  if (c[kCntCursor]) CK(cudaMemcpy(list.data(), ctx->d_tile_list.p, (size_t)c[kCntCursor] * sizeof(TileEntry), cudaMemcpyDeviceToHost));
  size_t total = 0;
  const int ty0 = ctx->frame.ty0, ty1 = ctx->frame.ty1, tx_n = ctx->frame.tiles_x;
  for (int ty = ty0; ty < ty1; ty++)
    for (int tx = 0; tx < tx_n; tx++) total += count[(size_t)ty * tx_n + tx] & kTileCountMask;
  if (n_offsets) *n_offsets = n_tiles + 1;
  if (n_entries) *n_entries = total;
  if (!tile_offsets || !entries) return FDC_OK;
  if (offsets_cap < n_tiles + 1 || entries_cap < total) return ctx->fail(FDC_ERR_INVALID, "debug_bins: buffers too small");
  size_t at = 0;
  for (size_t t = 0; t < n_tiles; t++) {
    tile_offsets[t] = (uint32_t)at;
    const int ty = (int)(t / tx_n);
    if (ty < ty0 || ty >= ty1) continue;
    for (uint32_t k = 0; k < (count[t] & kTileCountMask); k++) entries[at++] = calls[list[start[t] + k].pid];
  }
  tile_offsets[n_tiles] = (uint32_t)at;
  return FDC_OK;
}

}  // extern "C"
