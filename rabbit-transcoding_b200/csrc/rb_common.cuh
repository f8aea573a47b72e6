// rb_common.cuh — context, device buffers, launch bookkeeping shared by the sm_100a kernels.
//
// Everything here is plumbing for the C ABI of include/rabbit_b200.h.  No reference code is used; the
// reference (file:line) each kernel restates is cited at the kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "rabbit_b200.h"

#define RB_WARP 32

struct RbBuf {  // grow-only device allocation
  void*  p   = nullptr;
  size_t cap = 0;
  cudaError_t ensure( size_t bytes ) {
    if ( bytes <= cap ) { return cudaSuccess; }
    if ( p ) { cudaFree( p ); }
    p               = nullptr;
    cap             = 0;
    size_t      want = bytes + bytes / 8 + 256;
    cudaError_t e    = cudaMalloc( &p, want );
    if ( e == cudaSuccess ) { cap = want; }
    return e;
  }
  void release() {
    if ( p ) { cudaFree( p ); }
    p   = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>( p );
  }
};

// Device-side patch row: rb200_patch + its frame and index inside the frame.
struct RbPatch {
  int32_t u0, v0, su0, sv0, u1, v1, d1;
  int32_t s2dx, s2dy;
  int16_t frame_patch;  // index inside the frame (partition id)
  int16_t pad0;
  int32_t frame;
  int8_t  normal_axis, tangent_axis, bitangent_axis, mode, orient, lodx, lody, addplane;
};

struct RbTimingEntry {
  std::string name;
  cudaEvent_t a, b;
};

struct rb200_ctx {
  int          device     = 0;
  cudaStream_t stream     = nullptr;
  bool         own_stream = false;
  std::string  err;

  // ---- GOF description ----
  rb200_params P{};
  int          F = 0;
  int          W = 0, H = 0, oW = 0, oH = 0, Wb = 0, Hb = 0, M = 1, R = 16, prec = 4;
  int          bmWords = 0;  // occupancy bitmap words per row
  bool         have_gof = false, uploaded = false, reconstructed = false, geo_smoothed = false;
  bool         colors_transferred = false, color_smoothed = false, rgb_done = false;

  std::vector<rb200_patch>     h_patches;
  std::vector<int32_t>         h_patch_off;
  std::vector<rb200_eom_patch> h_eom;
  std::vector<int32_t>         h_eom_off, h_eom_members;
  std::vector<rb200_raw_patch> h_raw;
  std::vector<int32_t>         h_raw_off;
  int64_t                      nWI = 0;  // work items = patch blocks in emission order

  // ---- device inputs ----
  RbBuf d_occ_video, d_geometry, d_attribute, d_patches;
  RbBuf d_aux_geo, d_aux_attr;  // auxiliary video planes (raw points in a separate video)
  RbBuf d_raw_geo, d_raw_attr;  // decoder-native planes of rb200_gof_upload_yuv420 before the conversion kernels
  RbBuf d_wi_patch, d_wi_local, d_wi_count, d_wi_base, d_wi_eom_count, d_wi_eom_base, d_eom_order, d_wi_eom_slot;
  RbBuf d_frame_wi_off;  // [F+1] first work item of each frame
  RbBuf d_bitmap;        // [F][H][bmWords] uint32 full-resolution occupancy bits
  RbBuf d_b2p;           // [F][Hb][Wb] uint32
  RbBuf d_frame_info;    // per-frame device scalars (RbFrameInfo)
  RbBuf d_raw_desc;
  RbBuf d_plr_modes, d_plr_block_mode, d_plr_block_off;  // rb200_gof_set_plr
  bool  have_plr = false;

  // ---- device outputs: SoA cloud of the whole GOF, frame f = [h_frame_off[f], h_frame_off[f+1]) ----
  RbBuf d_pos;   // short4  {x, y, z, boundaryType}
  RbBuf d_col;   // ushort4 {c0, c1, c2, layer}
  RbBuf d_pix;   // uint32  x | y << 16
  RbBuf d_part;  // uint32  patch index (partition[])
  RbBuf d_rgb;   // uchar4  {r, g, b, 0}
  RbBuf d_pos_pre;  // copy of d_pos before geometry smoothing (tempFrameBuffer, PCCDecoder.cpp:435)
  RbBuf d_pack;     // staging for packed downloads
  // optional per-stage copies (rb200_enable_stage_snapshots): positions after reconstruction / geometry smoothing,
  // colours16 after reconstruction / transfer / colour smoothing — lets a frame-by-frame caller see the state each
  // stage left although every stage runs for the whole GOF at once
  bool  snapshots = false;
  RbBuf d_snap_pos[2], d_snap_col[3];
  bool  have_snap_pos[2] = {false, false}, have_snap_col[3] = {false, false, false};
  RbBuf d_blist, d_blist_n;  // indices of the boundary (type 1) points of the GOF + their count (device)
  RbBuf d_moved_bits;        // one bit per point of the GOF: moved by the geometry filter (type 3); cleared by the reconstruction
  RbBuf d_pbf;               // occupancy synthesis: the patch-local maps of the GOF (rb_pbf.cu)
  RbBuf d_bnd_bitmap;        // occupancy synthesis: [F][H][bmWords] PCCPatch::isBorder of every occupied pixel
  int64_t blist_cap = 0;     // that count on the host
  std::vector<int64_t>            h_frame_off;  // [F+1]
  std::vector<rb200_frame_counts> h_counts;
  RbBuf                           d_frame_off;  // [F+1] int64 on device

  // ---- scratch of the smoothing / transfer / metrics stages (owned by their translation units) ----
  RbBuf d_geo_grid, d_geo_cells, d_geo_cell_ids;
  RbBuf d_col_grid, d_col_cells, d_col_cell_ids, d_col_lum, d_col_lum_off;
  RbBuf d_scratch[8];
  int   geo_grid_w = 0, col_grid_w = 0;
  int   geo_grow = 0, col_grow = 0;  // smoothing tables: times the block pool was quadrupled after an overflow
  int64_t col_lum_want = 0;          // colour smoothing: luma list length asked for by a truncated cell
  int64_t geo_cell_cap = 0, col_cell_cap = 0;
  bool  geo_grid_clean = false, col_grid_clean = false;
  int   geo_grid_frames = 0, col_grid_frames = 0;

  // scratch of the transfer / metrics translation units: created lazily by them, freed by rb200_destroy (opaque here)
  void* transfer_scratch = nullptr;
  void* metrics_scratch  = nullptr;
  bool  pos_pre_valid    = false;  // d_pos_pre holds the pre-smoothing cloud of THIS GOF
  int   test_grid_shrink = 0;      // rb200_debug_set_grid_shrink

  // pinned host staging for small read-backs
  void*  h_pinned     = nullptr;
  size_t h_pinned_cap = 0;
  char*  h_ring       = nullptr;  // pinned ring for small host -> device tables (rb_pinned_ring)
  size_t h_ring_off   = 0;

  // ---- instrumentation ----
  rb200_launch_stats          stats{};
  bool                        timing = false;
  std::vector<RbTimingEntry>  timing_events;
  std::vector<std::string>    timing_names;
  std::vector<double>         timing_ms;
  std::vector<int64_t>        timing_n;
};

struct RbFrameInfo {  // device-resident per-frame scalars
  int32_t  max_coord;       // max over all coordinates of all points (PCCCodec.cpp:68-79)
  int32_t  geo_cells;       // number of marked geometry-smoothing cells
  int32_t  col_cells;       // number of marked colour-smoothing cells
  int32_t  smoothed;        // points moved (type 3)
  int32_t  recolored;       // points changed by colour smoothing
  int32_t  sum_overflow;    // a float accumulator would have left the exact range (App. A.3 guard)
  int32_t  eom_total;
  int32_t  pad;
};

int  rb_fail( rb200_ctx* c, int code, const char* fmt, ... );
int  rb_cuda( rb200_ctx* c, cudaError_t e, const char* what );
void rb_timing_begin( rb200_ctx* c, const char* name );
void rb_timing_end( rb200_ctx* c );
void* rb_pinned( rb200_ctx* c, size_t bytes );
// a fresh slice of a pinned ring for a small table that is copied to the device asynchronously and never read back:
// no wait before the host writes it (the stream is only synchronised when the ring wraps)
void* rb_pinned_ring( rb200_ctx* c, size_t bytes );

#define RB_CUDA( call )                                                \
  do {                                                                 \
    cudaError_t e__ = ( call );                                        \
    if ( e__ != cudaSuccess ) { return rb_cuda( c, e__, #call ); }     \
  } while ( 0 )

// kernel launch with launch counting + optional per-kernel event timing + error check
#define RB_LAUNCH( name, kern, grid, block, smem, ... )                                  \
  do {                                                                                   \
    rb_timing_begin( c, name );                                                          \
    kern<<<( grid ), ( block ), ( smem ), c->stream>>>( __VA_ARGS__ );                   \
    rb_timing_end( c );                                                                  \
    c->stats.kernel_launches++;                                                          \
    cudaError_t e__ = cudaGetLastError();                                                \
    if ( e__ != cudaSuccess ) { return rb_cuda( c, e__, name ); }                \
  } while ( 0 )

static inline int rb_div_up( int64_t a, int64_t b ) { return (int)( ( a + b - 1 ) / b ); }

// stage entry points implemented in the other translation units
int rb_reconstruct_impl( rb200_ctx* c );
int rb_pbf_impl( rb200_ctx* c );
int rb_smooth_geometry_impl( rb200_ctx* c );
int rb_transfer_colors_impl( rb200_ctx* c );
int rb_smooth_radius_impl( rb200_ctx* c );  // the non-grid smoothPointCloud (rb_transfer.cu: it needs the kd forest)
int rb_interleave_colors_impl( rb200_ctx* c );
int rb_smooth_color_impl( rb200_ctx* c );
int rb_convert_rgb8_impl( rb200_ctx* c );
int rb_ingest_yuv420_impl( rb200_ctx* c, int geo_bytes, int attr_bytes, int attr_bitdepth, int filter, int geo_shift, int attr_shift,
                           const int* geo_bitdepth /* in, out, msb */, const int* occ_bitdepth /* out, msb */ );
int rb_gather_nv12_impl( rb200_ctx* c, const rb200_frames_nv12* fr );
int rb_debug_rgb8_impl( rb200_ctx* c, const uint16_t* yuv, int64_t n, uint8_t* rgb, int force_f64 );
void rb_metrics_release( rb200_ctx* c );
void rb_transfer_release( rb200_ctx* c );
// exclusive scan of n uint32 (in == out allowed); `sums` needs rb_scan_scratch_bytes( n ) bytes
size_t rb_scan_scratch_bytes( int64_t n );
int    rb_scan_u32( rb200_ctx* c, const uint32_t* in, uint32_t* out, int64_t n, uint32_t* sums );
