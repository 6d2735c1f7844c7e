// Host-callable launchers of the sm_100a kernels.  All launches are asynchronous on `stream`.
#pragma once
#include <cuda_runtime.h>

#include "fdc_types.h"

namespace fdc {

struct SetupArgs {
  const fdc_call* draws;  // all draws of the frame (device)
  const fdc_rect64* rects64;  // compact records of the runs marked `compact`
  const RunState* runs;
  int n_runs;
  const Xform* xforms;
  uint32_t first, count;  // this segment's draw range [first, first+count)
  Prim* prims;            // [count] output
  PrimBin* prim_bins;     // [count] output: bbox / flags / inner rect for the binning kernels
  QuadGeom* geoms;        // [count] output (only PF_GENERAL entries are meaningful)
  PrimExt* exts;          // [count] output (only gradient PF_FAST entries are meaningful)
  uint32_t* prim_call;    // [count] backend-call ordinal per primitive (debug bins)
  AtlasView atlas;
  FrameView frame;
  uint32_t* counters;     // binning counters: the kernel zeroes the first `zero_counters` words for the binning that follows
  int zero_counters;
  uint32_t* row_cost;     // per-frame tile-entry totals per tile row (fine binner adds; band balancing reads); zeroed with the
  int n_row_cost;         //   per-frame counters (zero_counters == kNumCounters)
};
void launch_prim_setup(const SetupArgs& a, cudaStream_t stream);

struct BinBuffers {
  uint32_t* seg_table;     // [2 * n_chunks * n_cbins]: per (bin, chunk) the segment (start, count) of the coarse list
  uint32_t* coarse_list;   // [coarse_cap]
  uint32_t coarse_cap;
  uint32_t* tile_start;    // [tiles_x * tiles_y]
  uint32_t* tile_count;    // [tiles_x * tiles_y]
  TileEntry* tile_list;    // [tile_cap]
  uint32_t tile_cap;
  uint32_t* counters;      // [kNumCounters], see fdc_types.h (kCnt...)
  uint32_t* row_cost;      // [tiles_y] tile entries per tile row, accumulated over the frame's segments (may be null)
};
void launch_binning(const PrimBin* prim_bins, uint32_t n_prims, const FrameView& frame, const BinBuffers& b,
                    cudaStream_t stream, int* n_launches);

struct ShadeArgs {
  const Prim* prims;
  const QuadGeom* geoms;
  const PrimExt* exts;
  const RectMaskRec* rectmasks;
  const uint32_t* tile_start;
  const uint32_t* tile_count;
  const TileEntry* tile_list;
  const uint32_t* counters;  // kCntStickyOverflow set: the kernels leave the pixels untouched; kCntFullTiles: work for the full loop
  uint8_t* fb;               // RGBA8 W*H, top-left origin
  const uint8_t* backdrop;   // RGBA8 W*H (blurred copy for sdfModeBackdropBlur) or nullptr
  AtlasView atlas;
  FrameView frame;
  int load_dst;              // 0: start from clear colour; 1: read fb
  uint32_t clear_rgba8;
  unsigned long long* stats; // optional [8] debug counters (nullptr in production): visits partial/full/slow, entries, occl steps
  uint8_t* const* peers;     // optional peer framebuffers (device array of n_peers pointers) or nullptr
  int n_peers;
  uint8_t* multicast;        // optional NVSwitch multicast mapping of the framebuffer (all ranks' copies) or nullptr
  int fence_at_exit;         // system-scope fence after the remote stores (no flag barrier follows on the stream)
};
void launch_shade(const ShadeArgs& a, cudaStream_t stream);

constexpr int kMaxRanks = 16;
constexpr int kFlagError = 32;  // word of a rank's flag array that records barrier time-outs (bit r: rank r never arrived)
constexpr int kFlagSignalSeq = 40, kFlagWaitSeq = 41;  // this rank's running signal / wait numbers (local use only)
struct BlurArgs {
  const uint8_t* src;  // framebuffer
  // Tile-band partition: rows [band_end_px[r-1], band_end_px[r]) of the frame live in src_rank[r] (this rank's own framebuffer or
  // a peer's, read over NVLink for the blur halo).  n_src == 0: everything is in `src`.
  const uint8_t* src_rank[kMaxRanks];
  int n_src;
  int band_end_px[kMaxRanks];  // rank r owns rows [band_end_px[r-1], band_end_px[r])
  uint8_t* temp;       // H-pass output
  uint8_t* dst;        // backdrop (V-pass output)
  int W, H;
  int x0, y0, x1, y1;  // region of dst that is needed (composite quad bbox, clipped)
  float radius;        // blurRadius as passed to drawBackdropBlur
};
void launch_backdrop_blur(const BlurArgs& a, cudaStream_t stream, int* n_launches);
// Cross-rank stream-ordered barrier over peer memory: store this rank's next signal number into slot `my_rank` of every
// rank's flag array; (wait) spin until all `n` slots of the local array have reached this rank's next wait number.
// Signals and waits pair up in issue order; every rank issues the same sequence.
void launch_signal_flags(uint32_t* const* flag_arrays, int n, int my_rank, cudaStream_t stream);
void launch_wait_flags(uint32_t* my_flags, int n, cudaStream_t stream);

// Copies `bytes` (a multiple of 16) at `src` (= own copy + byte_off) to the same offset of every rank's copy of a shared
// buffer: one multimem.st per 16 bytes through the multicast mapping, or one store per peer.
void launch_push_to_peers(const uint8_t* src, size_t byte_off, size_t bytes, uint8_t* multicast, uint8_t* const* peers, int n_peers,
                          int my_rank, cudaStream_t stream);

// One CTA per glyph: signed-area accumulation of the outline, alpha into the glyph's atlas slot (fdc_glyph.cu).
// `glyphs_dev`: device array of {first_seg, n_segs, w, h, atlas_x, atlas_y} (six 32-bit words each).
void launch_glyph_raster(const void* glyphs_dev, int n_glyphs, const fdc_outline_seg* segs_dev, uint8_t* atlas_level0, int atlas_size,
                         int lcd_filter, cudaStream_t stream);

void launch_fill_u32(uint32_t* dst, uint32_t value, size_t n, cudaStream_t stream);
// Builds mip level `l+1` region from level `l` (premultiplied 2x2 box, see oracle upload_chain).
void launch_mip_down(const uint8_t* src, int src_size, uint8_t* dst, int dst_size, int sx, int sy, int sw, int sh,
                     int dx, int dy, cudaStream_t stream);

}  // namespace fdc
