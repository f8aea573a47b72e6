// rb_context.cu — context lifetime, frame ingest, downloads, instrumentation (C ABI plumbing).
#include <stdarg.h>

#include <algorithm>

#include "rb_common.cuh"

int rb_fail( rb200_ctx* c, int code, const char* fmt, ... ) {
  char    buf[512];
  va_list ap;
  va_start( ap, fmt );
  vsnprintf( buf, sizeof( buf ), fmt, ap );
  va_end( ap );
  if ( c ) { c->err = buf; }
  return code;
}

int rb_cuda( rb200_ctx* c, cudaError_t e, const char* what ) {
  return rb_fail( c, RB200_ERR_CUDA, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString( e ), what );
}

void rb_timing_begin( rb200_ctx* c, const char* name ) {
  if ( !c->timing ) { return; }
  RbTimingEntry t;
  t.name = name;
  cudaEventCreate( &t.a );
  cudaEventCreate( &t.b );
  cudaEventRecord( t.a, c->stream );
  c->timing_events.push_back( t );
}
void rb_timing_end( rb200_ctx* c ) {
  if ( !c->timing ) { return; }
  cudaEventRecord( c->timing_events.back().b, c->stream );
}

void* rb_pinned( rb200_ctx* c, size_t bytes ) {
  if ( bytes > c->h_pinned_cap ) {
    if ( c->h_pinned ) { cudaFreeHost( c->h_pinned ); }
    c->h_pinned     = nullptr;
    c->h_pinned_cap = 0;
    if ( cudaMallocHost( &c->h_pinned, bytes + 4096 ) != cudaSuccess ) { return nullptr; }
    c->h_pinned_cap = bytes + 4096;
  }
  return c->h_pinned;
}

void* rb_pinned_ring( rb200_ctx* c, size_t bytes ) {
  constexpr size_t RING = 4u << 20;
  bytes = ( bytes + 255 ) & ~size_t( 255 );
  if ( bytes > RING ) { return nullptr; }
  if ( !c->h_ring ) {
    if ( cudaMallocHost( (void**)&c->h_ring, RING ) != cudaSuccess ) { return nullptr; }
    c->h_ring_off = 0;
  }
  if ( c->h_ring_off + bytes > RING ) {  // every earlier slice was handed to a copy on c->stream
    cudaStreamSynchronize( c->stream );
    c->h_ring_off = 0;
  }
  void* p = c->h_ring + c->h_ring_off;
  c->h_ring_off += bytes;
  return p;
}

static void rb_timing_resolve( rb200_ctx* c ) {
  if ( c->timing_events.empty() ) { return; }
  cudaStreamSynchronize( c->stream );
  for ( auto& t : c->timing_events ) {
    float ms = 0.f;
    cudaEventElapsedTime( &ms, t.a, t.b );
    size_t k = 0;
    for ( ; k < c->timing_names.size(); k++ ) {
      if ( c->timing_names[k] == t.name ) { break; }
    }
    if ( k == c->timing_names.size() ) {
      c->timing_names.push_back( t.name );
      c->timing_ms.push_back( 0 );
      c->timing_n.push_back( 0 );
    }
    c->timing_ms[k] += ms;
    c->timing_n[k] += 1;
    cudaEventDestroy( t.a );
    cudaEventDestroy( t.b );
  }
  c->timing_events.clear();
}

// ------------------------------------------------------------------------------------------------
// pack kernels: device SoA records -> the reference's std::vector element layouts (PCCPointSet.h:520-531)
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_positions( const short4* __restrict__ pos, int64_t n, int16_t* __restrict__ out ) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  short4 p       = pos[i];
  out[i * 3 + 0] = p.x;
  out[i * 3 + 1] = p.y;
  out[i * 3 + 2] = p.z;
}
__global__ void k_pack_types( const short4* __restrict__ pos, int64_t n, uint16_t* __restrict__ out ) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  out[i] = (uint16_t)pos[i].w;
}
__global__ void k_pack_colors16( const ushort4* __restrict__ col, int64_t n, uint16_t* __restrict__ out ) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  ushort4 p      = col[i];
  out[i * 3 + 0] = p.x;
  out[i * 3 + 1] = p.y;
  out[i * 3 + 2] = p.z;
}
__global__ void k_pack_rgb( const uchar4* __restrict__ rgb, int64_t n, uint8_t* __restrict__ out ) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  uchar4 p       = rgb[i];
  out[i * 3 + 0] = p.x;
  out[i * 3 + 1] = p.y;
  out[i * 3 + 2] = p.z;
}
__global__ void k_pack_pixels( const uint32_t* __restrict__ pix,
                               const ushort4* __restrict__ col,
                               int64_t n,
                               uint32_t* __restrict__ out ) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  uint32_t p     = pix[i];
  out[i * 3 + 0] = p & 0xFFFFu;
  out[i * 3 + 1] = p >> 16;
  out[i * 3 + 2] = col[i].w;
}
__global__ void k_expand_bitmap( const uint32_t* __restrict__ bm, int W, int H, int words, uint8_t* __restrict__ out ) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= (int64_t)W * H ) { return; }
  int y = (int)( i / W ), x = (int)( i % W );
  out[i] = ( bm[(size_t)y * words + ( x >> 5 )] >> ( x & 31 ) ) & 1u;
}

// generateOccupancyMap (:1584-1606) for one frame: thread (v, u) of the full-resolution map.  The reference visits the
// pixels in row-major order and writes the 0/1 result back into the video sample (:1599-1600), so the FIRST visit of a
// sample (its top-left pixel) sees the decoded value and every later visit sees the 0/1 left by the first one.
__global__ void k_occupancy_frame( const uint8_t* __restrict__ video, int oW, int oH, int prec, int threshold, int eom,
                                   uint32_t* __restrict__ map, uint8_t* __restrict__ video_out ) {
  const int     W = oW * prec, H = oH * prec;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= (int64_t)W * H ) { return; }
  const int u = (int)( i % W ), v = (int)( i / W );
  const int s = video[( v / prec ) * oW + u / prec];
  if ( eom ) {
    map[i] = (uint32_t)s;
    return;
  }
  const int  first = s > threshold;
  const bool top   = ( u % prec == 0 ) && ( v % prec == 0 );
  const int  val   = top ? first : ( first > threshold );
  map[i]           = (uint32_t)val;
  // the value the sample holds after its last visit (bottom-right pixel)
  if ( ( u % prec == prec - 1 ) && ( v % prec == prec - 1 ) ) { video_out[( v / prec ) * oW + u / prec] = (uint8_t)( prec == 1 ? first : val ); }
}

extern "C" {

int rb200_abi_version( void ) { return RB200_ABI_VERSION; }

void* rb200_host_alloc( size_t bytes ) {
  void* p = nullptr;
  if ( cudaMallocHost( &p, bytes ? bytes : 1 ) != cudaSuccess ) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void rb200_host_free( void* p ) {
  if ( p ) { cudaFreeHost( p ); }
}

int rb200_occupancy_map( rb200_ctx* c, uint8_t* video, int oW, int oH, int prec, int threshold, int eom, uint32_t* map_out ) {
  if ( !c || !video || !map_out || oW <= 0 || oH <= 0 || prec < 1 ) { return rb_fail( c, RB200_ERR_INVALID, "occupancy_map: bad arguments" ); }
  cudaSetDevice( c->device );
  const size_t nv = (size_t)oW * oH, nm = nv * prec * prec;
  RB_CUDA( c->d_pack.ensure( nm * 4 + 2 * nv + 512 ) );
  uint32_t* dMap = c->d_pack.as<uint32_t>();
  uint8_t*  dIn  = c->d_pack.as<uint8_t>() + nm * 4, *dOut = dIn + ( ( nv + 255 ) & ~size_t( 255 ) );
  RB_CUDA( cudaMemcpyAsync( dIn, video, nv, cudaMemcpyDefault, c->stream ) );
  RB_CUDA( cudaMemcpyAsync( dOut, dIn, nv, cudaMemcpyDeviceToDevice, c->stream ) );
  RB_LAUNCH( "occupancy_frame", k_occupancy_frame, rb_div_up( nm, 256 ), 256, 0, dIn, oW, oH, prec, threshold, eom, dMap, dOut );
  RB_CUDA( cudaMemcpyAsync( map_out, dMap, nm * 4, cudaMemcpyDefault, c->stream ) );
  RB_CUDA( cudaMemcpyAsync( video, dOut, nv, cudaMemcpyDefault, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  c->stats.h2d_bytes += (int64_t)nv;
  c->stats.d2h_bytes += (int64_t)( nm * 4 + nv );
  return RB200_OK;
}

int rb200_create( int device, rb200_ctx** out ) {
  if ( !out ) { return RB200_ERR_INVALID; }
  *out = nullptr;
  int n = 0;
  if ( cudaGetDeviceCount( &n ) != cudaSuccess || n <= 0 ) { return RB200_ERR_CUDA; }  // no CPU fallback
  if ( device < 0 || device >= n ) { return RB200_ERR_INVALID; }
  if ( cudaSetDevice( device ) != cudaSuccess ) { return RB200_ERR_CUDA; }
  rb200_ctx* c = new rb200_ctx;
  c->device    = device;
  if ( cudaStreamCreateWithFlags( &c->stream, cudaStreamNonBlocking ) != cudaSuccess ) {
    delete c;
    return RB200_ERR_CUDA;
  }
  c->own_stream = true;
  *out          = c;
  return RB200_OK;
}

void rb200_destroy( rb200_ctx* c ) {
  if ( !c ) { return; }
  cudaSetDevice( c->device );
  cudaStreamSynchronize( c->stream );
  RbBuf* bufs[] = {&c->d_aux_geo, &c->d_aux_attr, &c->d_occ_video, &c->d_geometry, &c->d_attribute, &c->d_raw_geo, &c->d_raw_attr, &c->d_patches, &c->d_wi_patch, &c->d_wi_local,
                   &c->d_wi_count, &c->d_wi_base, &c->d_wi_eom_count, &c->d_wi_eom_base, &c->d_eom_order,
                   &c->d_wi_eom_slot, &c->d_frame_wi_off, &c->d_bitmap, &c->d_b2p, &c->d_frame_info, &c->d_raw_desc,
                   &c->d_plr_modes, &c->d_plr_block_mode, &c->d_plr_block_off,
                   &c->d_pos, &c->d_col, &c->d_pix, &c->d_part, &c->d_rgb, &c->d_pos_pre, &c->d_pack, &c->d_frame_off,
                   &c->d_geo_grid, &c->d_geo_cells, &c->d_geo_cell_ids, &c->d_col_grid, &c->d_col_cells,
                   &c->d_col_cell_ids, &c->d_col_lum, &c->d_col_lum_off, &c->d_blist, &c->d_blist_n, &c->d_moved_bits, &c->d_pbf, &c->d_bnd_bitmap,
                   &c->d_snap_pos[0], &c->d_snap_pos[1], &c->d_snap_col[0], &c->d_snap_col[1], &c->d_snap_col[2]};
  for ( auto* b : bufs ) { b->release(); }
  for ( auto& b : c->d_scratch ) { b.release(); }
  rb_metrics_release( c );
  rb_transfer_release( c );
  if ( c->h_pinned ) { cudaFreeHost( c->h_pinned ); }
  if ( c->h_ring ) { cudaFreeHost( c->h_ring ); }
  for ( auto& t : c->timing_events ) {
    cudaEventDestroy( t.a );
    cudaEventDestroy( t.b );
  }
  if ( c->own_stream ) { cudaStreamDestroy( c->stream ); }
  delete c;
}

const char* rb200_error_string( const rb200_ctx* c ) { return c ? c->err.c_str() : "null context"; }

int rb200_set_stream( rb200_ctx* c, void* s ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  cudaSetDevice( c->device );
  cudaStreamSynchronize( c->stream );
  if ( s == nullptr ) {
    if ( !c->own_stream ) {
      RB_CUDA( cudaStreamCreateWithFlags( &c->stream, cudaStreamNonBlocking ) );
      c->own_stream = true;
    }
  } else {
    if ( c->own_stream ) { cudaStreamDestroy( c->stream ); }
    c->own_stream = false;
    c->stream     = (cudaStream_t)s;
  }
  return RB200_OK;
}

int rb200_synchronize( rb200_ctx* c ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  return RB200_OK;
}

int rb200_gof_begin( rb200_ctx* c, const rb200_params* p, int nFrames ) {
  if ( !c || !p || nFrames <= 0 ) { return rb_fail( c, RB200_ERR_INVALID, "gof_begin: bad arguments" ); }
  cudaSetDevice( c->device );
  const int R = p->occupancy_resolution, pr = p->occupancy_precision;
  if ( p->width <= 0 || p->height <= 0 || R != 16 ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "occupancy_resolution must be 16 (got %d)", R );
  }
  if ( !( pr == 1 || pr == 2 || pr == 4 || pr == 8 || pr == 16 ) ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "occupancy_precision %d unsupported", pr );
  }
  if ( p->width % R || p->height % R || p->width > 65535 || p->height > 65535 ) {
    return rb_fail( c, RB200_ERR_INVALID, "atlas %dx%d must be a multiple of %d and < 65536", p->width, p->height, R );
  }
  if ( p->pbf_enable ) {  // occupancy synthesis (PCCCodec.cpp:541-554, PCCPatch.cpp:797-977)
    if ( p->enhanced_occupancy_map_code || p->single_map_pixel_interleaving || p->point_local_reconstruction ) {
      return rb_fail( c, RB200_ERR_UNSUPPORTED, "occupancy synthesis together with EOM, pixel interleaving or point local "
                                                "reconstruction is not implemented" );
    }
    if ( p->enable_size_quantization ) {  // the reference clears pixels of a map it never allocated there (:571-597)
      return rb_fail( c, RB200_ERR_UNSUPPORTED, "occupancy synthesis together with patch size quantisation is undefined in the reference" );
    }
    if ( p->pbf_passes_count < 1 || p->pbf_filter_size < 1 || p->pbf_log2_threshold < 1 ) {
      return rb_fail( c, RB200_ERR_INVALID, "pbf_enable needs pbf_passes_count, pbf_filter_size and pbf_log2_threshold >= 1" );
    }
  }
  if ( p->single_map_pixel_interleaving || p->point_local_reconstruction ) {
    // generatePoints :350-471 / :472-496 + transferColorWeight (colorPointCloud :1367-1434)
    if ( p->map_count_minus1 != 0 ) {
      return rb_fail( c, RB200_ERR_INVALID, "pixel interleaving / point local reconstruction need a single map" );
    }
    if ( p->single_map_pixel_interleaving && p->surface_thickness < 1 ) {
      return rb_fail( c, RB200_ERR_INVALID, "single_map_pixel_interleaving needs surface_thickness >= 1" );
    }
    if ( p->enhanced_occupancy_map_code || p->multiple_streams ) {
      return rb_fail( c, RB200_ERR_UNSUPPORTED, "pixel interleaving / point local reconstruction together with EOM or "
                                                "multiple streams is not implemented" );
    }
  }
  if ( p->map_count_minus1 < 0 || p->map_count_minus1 > 1 ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "map_count_minus1 must be 0 or 1" );
  }
  if ( p->relative_t1 && ( !p->multiple_streams || p->map_count_minus1 != 1 ) ) {
    return rb_fail( c, RB200_ERR_INVALID, "relative_t1 needs multiple_streams and two maps (PCCCodec.cpp:1387-1416)" );
  }
  if ( p->relative_t1 && ( p->enhanced_occupancy_map_code || p->use_additional_points_patch ) ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "relative_t1 together with EOM or raw patches is not implemented" );
  }
  if ( p->use_aux_separate_video ) {
    if ( ( p->use_additional_points_patch || p->enhanced_occupancy_map_code ) && ( p->aux_width < p->occupancy_resolution || p->aux_height < p->occupancy_resolution ||
                                            p->aux_width % p->occupancy_resolution || p->aux_height % p->occupancy_resolution ||
                                            p->aux_width > 16384 || p->aux_height > 16384 ) ) {
      return rb_fail( c, RB200_ERR_INVALID, "use_aux_separate_video needs aux_width / aux_height (multiples of the occupancy resolution)" );
    }
  }
  if ( p->geometry_bitdepth_3d < 1 || p->geometry_bitdepth_3d > 14 ) {
    return rb_fail( c, RB200_ERR_INVALID, "geometry_bitdepth_3d out of range" );
  }
  c->P       = *p;
  c->F       = nFrames;
  c->W       = p->width;
  c->H       = p->height;
  c->R       = R;
  c->prec    = pr;
  c->oW      = c->W / pr;
  c->oH      = c->H / pr;
  c->Wb      = c->W / R;
  c->Hb      = c->H / R;
  c->M       = p->map_count_minus1 + 1;
  c->bmWords = ( c->W + 31 ) / 32;
  c->have_gof = true;
  c->have_plr = false;
  c->uploaded = c->reconstructed = c->geo_smoothed = c->colors_transferred = c->color_smoothed = c->rgb_done = false;
  const size_t F = nFrames;
  RB_CUDA( c->d_occ_video.ensure( F * c->oW * c->oH ) );
  RB_CUDA( c->d_geometry.ensure( F * c->M * (size_t)c->W * c->H * 2 ) );
  if ( p->attribute_count > 0 ) { RB_CUDA( c->d_attribute.ensure( F * c->M * 3 * (size_t)c->W * c->H * 2 ) ); }
  RB_CUDA( c->d_bitmap.ensure( F * (size_t)c->H * c->bmWords * 4 ) );
  RB_CUDA( c->d_b2p.ensure( F * (size_t)c->Wb * c->Hb * 4 ) );
  RB_CUDA( c->d_frame_info.ensure( F * sizeof( RbFrameInfo ) ) );
  RB_CUDA( c->d_frame_off.ensure( ( F + 1 ) * 8 ) );
  c->h_frame_off.assign( F + 1, 0 );
  c->h_counts.assign( F, rb200_frame_counts{} );
  return RB200_OK;
}

// PCCPatch::patchBlock2CanvasBlock footprint check (PCCPatch.cpp:253-308): every patch block must land
// inside the canvas, otherwise the reference exits 180 (PCCPatch.cpp:237-245).
static bool patch_inside( const rb200_patch& p, int Wb, int Hb ) {
  // SWAP, ROT90, ROT270, MROT90, MROT270 exchange U and V on the canvas (PCCBitstreamCommon.h:120-130)
  const bool sw = ( p.orientation == 1 || p.orientation == 2 || p.orientation == 4 || p.orientation == 6 ||
                    p.orientation == 8 );
  const int  cw = sw ? p.size_v0 : p.size_u0, ch = sw ? p.size_u0 : p.size_v0;
  return p.u0 >= 0 && p.v0 >= 0 && p.size_u0 >= 0 && p.size_v0 >= 0 && p.u0 + cw <= Wb && p.v0 + ch <= Hb;
}

// planes either as the reconstruction reads them (fr) or decoder-native (fy): see rb200_gof_upload_yuv420
static int gof_upload_common( rb200_ctx* c, const rb200_frames* fr, const rb200_frames_yuv420* fy, const rb200_atlas* at,
                              const rb200_frames_nv12* fn = nullptr ) {
  if ( !c || ( !fr && !fy ) || !at ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload: null argument" ); }
  if ( !c->have_gof ) { return rb_fail( c, RB200_ERR_STATE, "gof_upload before gof_begin" ); }
  cudaSetDevice( c->device );
  const size_t F = c->F;
  if ( fr && ( !fr->occupancy || !fr->geometry || ( c->P.attribute_count > 0 && !fr->attribute ) || !at->patch_offset ) ) {
    return rb_fail( c, RB200_ERR_INVALID, "gof_upload: missing plane or patch table" );
  }
  if ( fy ) {
    if ( !fn && ( !fy->occupancy || !fy->geometry || ( c->P.attribute_count > 0 && !fy->attribute ) ) ) {
      return rb_fail( c, RB200_ERR_INVALID, "gof_upload_yuv420: missing plane or patch table" );
    }
    if ( !at->patch_offset ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload_yuv420: missing plane or patch table" ); }
    if ( ( fy->geometry_sample_bytes != 1 && fy->geometry_sample_bytes != 2 ) ||
         ( fy->attribute_sample_bytes != 1 && fy->attribute_sample_bytes != 2 ) ||
         ( fy->attribute_bitdepth != 8 && fy->attribute_bitdepth != 10 ) || fy->upsampling_filter < 0 || fy->upsampling_filter > 7 ) {
      return rb_fail( c, RB200_ERR_INVALID, "gof_upload_yuv420: sample bytes must be 1 or 2, bit depth 8 or 10, filter 0..7" );
    }
    if ( fy->geometry_bitdepth_out < 0 || fy->geometry_bitdepth_out > 16 || fy->occupancy_bitdepth_out < 0 || fy->occupancy_bitdepth_out > 8 ||
         ( fy->geometry_bitdepth_out > 0 && ( fy->geometry_bitdepth_in < 1 || fy->geometry_bitdepth_in > 16 ) ) ) {
      // PCCImage.cpp:262-269: a bit depth beyond the sample type prints and exits
      return rb_fail( c, RB200_ERR_INVALID, "gof_upload_yuv420: wrong bitdepth parameter (geometry in/out 1..16, occupancy out 1..8)" );
    }
    if ( fy->geometry_shift < 0 || fy->geometry_shift > 9 || fy->attribute_shift < 0 || fy->attribute_shift > 9 ) {
      return rb_fail( c, RB200_ERR_INVALID, "gof_upload_yuv420: shift must be 0..9 (PCCImage.h:118-119)" );
    }
    if ( c->P.attribute_rgb444 ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload_yuv420: RGB444 attributes are not 4:2:0 video" ); }
    if ( ( c->W & 15 ) || ( c->H & 1 ) ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload_yuv420: width must be a multiple of 16, height even" ); }
  }
  // ---- patch tables (host) ----
  const int nPatches = at->patch_offset[F];
  if ( nPatches < 0 || at->patch_offset[0] != 0 || ( nPatches > 0 && !at->patches ) ) {
    return rb_fail( c, RB200_ERR_INVALID, "gof_upload: bad patch offsets" );
  }
  c->h_patches.assign( at->patches, at->patches + nPatches );
  c->h_patch_off.assign( at->patch_offset, at->patch_offset + F + 1 );
  std::vector<RbPatch> dp( nPatches );
  for ( size_t f = 0; f < F; f++ ) {
    if ( c->h_patch_off[f + 1] < c->h_patch_off[f] ) { return rb_fail( c, RB200_ERR_INVALID, "patch offsets not monotone" ); }
    if ( c->h_patch_off[f + 1] - c->h_patch_off[f] > 32000 ) {
      return rb_fail( c, RB200_ERR_UNSUPPORTED, "more than 32000 patches in a frame" );
    }
    for ( int i = c->h_patch_off[f]; i < c->h_patch_off[f + 1]; i++ ) {
      const rb200_patch& s = c->h_patches[i];
      if ( s.orientation < 0 || s.orientation > 8 || s.normal_axis < 0 || s.normal_axis > 2 || s.tangent_axis < 0 ||
           s.tangent_axis > 2 || s.bitangent_axis < 0 || s.bitangent_axis > 2 || s.axis_of_additional_plane < 0 ||
           s.axis_of_additional_plane > 3 ) {
        return rb_fail( c, RB200_ERR_INVALID, "frame %zu patch %d: bad orientation / axes", f, i - c->h_patch_off[f] );
      }
      if ( !patch_inside( s, c->Wb, c->Hb ) ) {
        return rb_fail( c, RB200_ERR_PATCH_OUT_OF_CANVAS,
                        "patch2Canvas (x,y) is out of boundary : frame %zu patch %d canvassize %dx%d", f,
                        i - c->h_patch_off[f], c->W, c->H );
      }
      RbPatch& d       = dp[i];
      d.u0             = s.u0;
      d.v0             = s.v0;
      d.su0            = s.size_u0;
      d.sv0            = s.size_v0;
      d.u1             = s.u1;
      d.v1             = s.v1;
      d.d1             = s.d1;
      d.s2dx           = s.size2d_x_px;
      d.s2dy           = s.size2d_y_px;
      d.frame_patch    = (int16_t)( i - c->h_patch_off[f] );
      d.pad0           = 0;
      d.frame          = (int32_t)f;
      d.normal_axis    = (int8_t)s.normal_axis;
      d.tangent_axis   = (int8_t)s.tangent_axis;
      d.bitangent_axis = (int8_t)s.bitangent_axis;
      d.mode           = (int8_t)s.projection_mode;
      d.orient         = (int8_t)s.orientation;
      d.lodx           = (int8_t)s.lod_x;
      d.lody           = (int8_t)s.lod_y;
      d.addplane       = (int8_t)s.axis_of_additional_plane;
    }
  }
  // work items: patch blocks in the reference's emission order (PCCCodec.cpp:628-650):
  // patch index order (reversed under patchPrecedenceOrderFlag, :629), then v0, then u0.
  std::vector<int32_t> wiPatch, wiLocal, frameWiOff( F + 1, 0 );
  for ( size_t f = 0; f < F; f++ ) {
    frameWiOff[f] = (int32_t)wiPatch.size();
    const int b = c->h_patch_off[f], e = c->h_patch_off[f + 1];
    for ( int k = 0; k < e - b; k++ ) {
      const int i  = c->P.patch_precedence_reverse ? ( e - 1 - k ) : ( b + k );
      const int nb = c->h_patches[i].size_u0 * c->h_patches[i].size_v0;
      for ( int j = 0; j < nb; j++ ) {
        wiPatch.push_back( i );
        wiLocal.push_back( j );
      }
    }
  }
  frameWiOff[F] = (int32_t)wiPatch.size();
  c->nWI        = (int64_t)wiPatch.size();

  // EOM / raw tables
  c->h_eom.clear();
  c->h_eom_off.assign( F + 1, 0 );
  c->h_eom_members.clear();
  if ( c->P.enhanced_occupancy_map_code && at->eom_patches && at->eom_offset ) {
    const int n = at->eom_offset[F];
    if ( at->eom_offset[0] != 0 || n < 0 ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload: bad EOM patch offsets" ); }
    for ( size_t f = 0; f < F; f++ ) {
      if ( at->eom_offset[f + 1] < at->eom_offset[f] ) { return rb_fail( c, RB200_ERR_INVALID, "EOM patch offsets not monotone" ); }
    }
    c->h_eom.assign( at->eom_patches, at->eom_patches + n );
    c->h_eom_off.assign( at->eom_offset, at->eom_offset + F + 1 );
    int nm = 0;
    for ( auto& e : c->h_eom ) {
      if ( e.member_begin < 0 || e.member_count < 0 || e.member_begin > ( 1 << 30 ) - e.member_count ) {
        return rb_fail( c, RB200_ERR_INVALID, "EOM patch member range is negative" );
      }
      nm = std::max( nm, e.member_begin + e.member_count );
    }
    if ( nm > 0 ) {
      if ( !at->eom_members ) { return rb_fail( c, RB200_ERR_INVALID, "eom_members missing" ); }
      c->h_eom_members.assign( at->eom_members, at->eom_members + nm );
    }
  }
  c->h_raw.clear();
  c->h_raw_off.assign( F + 1, 0 );
  if ( c->P.use_additional_points_patch && at->raw_patches && at->raw_offset ) {
    const int n = at->raw_offset[F];
    if ( at->raw_offset[0] != 0 || n < 0 ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload: bad raw patch offsets" ); }
    for ( size_t f = 0; f < F; f++ ) {
      if ( at->raw_offset[f + 1] < at->raw_offset[f] ) { return rb_fail( c, RB200_ERR_INVALID, "raw patch offsets not monotone" ); }
    }
    c->h_raw.assign( at->raw_patches, at->raw_patches + n );
    c->h_raw_off.assign( at->raw_offset, at->raw_offset + F + 1 );
    for ( auto& r : c->h_raw ) {
      // the patch stores X, then Y, then Z: 3 * numberOfRawPoints samples must fit its rectangle (the reference's fill
      // loop is bounded by sizeU * sizeV, PCCCodec.cpp:913-928, and never leaves the patch)
      // (in the auxiliary video the rectangle lies in the auxiliary frame)
      const int wb = c->P.use_aux_separate_video ? c->P.aux_width / c->R : c->Wb, hb = c->P.use_aux_separate_video ? c->P.aux_height / c->R : c->Hb;
      if ( r.u0 < 0 || r.v0 < 0 || r.size_u0 < 0 || r.size_v0 < 0 || r.num_points < 0 || r.u0 + r.size_u0 > wb ||
           r.v0 + r.size_v0 > hb || 3 * (int64_t)r.num_points > (int64_t)r.size_u0 * r.size_v0 * c->R * c->R ) {
        return rb_fail( c, RB200_ERR_INVALID, "raw patch outside the canvas or too small for its 3 x %d samples", r.num_points );
      }
    }
  }

  // ---- device copies ----
  const size_t occBytes = F * (size_t)c->oW * c->oH;
  const size_t geoBytes = F * c->M * (size_t)c->W * c->H * 2;
  const size_t attBytes = c->P.attribute_count > 0 ? F * c->M * 3 * (size_t)c->W * c->H * 2 : 0;
  if ( c->P.use_aux_separate_video && ( c->P.use_additional_points_patch || c->P.enhanced_occupancy_map_code ) ) {
    if ( !fr || ( c->P.use_additional_points_patch && !fr->aux_geometry ) || ( c->P.attribute_count > 0 && !fr->aux_attribute ) ) {
      return rb_fail( c, RB200_ERR_UNSUPPORTED, "gof_upload: the auxiliary video planes are taken through rb200_gof_upload (aux_geometry / aux_attribute)" );
    }
    const size_t ap = (size_t)c->P.aux_width * c->P.aux_height;
    if ( fr->aux_geometry ) {
      RB_CUDA( c->d_aux_geo.ensure( F * ap * 2 ) );
      RB_CUDA( cudaMemcpyAsync( c->d_aux_geo.p, fr->aux_geometry, F * ap * 2, cudaMemcpyDefault, c->stream ) );
      c->stats.h2d_bytes += (int64_t)( F * ap * 2 );
    }
    if ( c->P.attribute_count > 0 ) {
      RB_CUDA( c->d_aux_attr.ensure( F * ap * 6 ) );
      RB_CUDA( cudaMemcpyAsync( c->d_aux_attr.p, fr->aux_attribute, F * ap * 6, cudaMemcpyDefault, c->stream ) );
      c->stats.h2d_bytes += (int64_t)( F * ap * 6 );
    }
  }
  if ( fr ) {
    RB_CUDA( cudaMemcpyAsync( c->d_occ_video.p, fr->occupancy, occBytes, cudaMemcpyDefault, c->stream ) );
    RB_CUDA( cudaMemcpyAsync( c->d_geometry.p, fr->geometry, geoBytes, cudaMemcpyDefault, c->stream ) );
    if ( attBytes ) { RB_CUDA( cudaMemcpyAsync( c->d_attribute.p, fr->attribute, attBytes, cudaMemcpyDefault, c->stream ) ); }
    c->stats.h2d_bytes += (int64_t)( occBytes + geoBytes + attBytes );
  } else if ( fn ) {  // pitched decoder surfaces in device / pinned memory: gathered by a kernel, then the same conversion
    int r = rb_gather_nv12_impl( c, fn );
    if ( r ) { return r; }
    const int gbd[3] = {fy->geometry_bitdepth_in, fy->geometry_bitdepth_out, fy->geometry_msb_align};
    const int obd[2] = {fy->occupancy_bitdepth_out, fy->occupancy_msb_align};
    r = rb_ingest_yuv420_impl( c, fy->geometry_sample_bytes, fy->attribute_sample_bytes, fy->attribute_bitdepth, fy->upsampling_filter,
                               fy->geometry_shift, fy->attribute_shift, gbd, obd );
    if ( r ) { return r; }
  } else {  // decoder-native planes: a quarter of the bytes cross PCIe, the conversion runs on the device
    const size_t plane  = (size_t)c->W * c->H;
    const size_t rawGeo = F * c->M * plane * fy->geometry_sample_bytes;
    const size_t rawAtt = attBytes ? F * c->M * ( plane + plane / 2 ) * fy->attribute_sample_bytes : 0;
    RB_CUDA( cudaMemcpyAsync( c->d_occ_video.p, fy->occupancy, occBytes, cudaMemcpyDefault, c->stream ) );
    if ( fy->geometry_sample_bytes == 2 ) {
      RB_CUDA( cudaMemcpyAsync( c->d_geometry.p, fy->geometry, rawGeo, cudaMemcpyDefault, c->stream ) );
    } else {
      RB_CUDA( c->d_raw_geo.ensure( rawGeo + 64 ) );
      RB_CUDA( cudaMemcpyAsync( c->d_raw_geo.p, fy->geometry, rawGeo, cudaMemcpyDefault, c->stream ) );
    }
    if ( rawAtt ) {
      RB_CUDA( c->d_raw_attr.ensure( rawAtt + 64 ) );
      RB_CUDA( cudaMemcpyAsync( c->d_raw_attr.p, fy->attribute, rawAtt, cudaMemcpyDefault, c->stream ) );
    }
    c->stats.h2d_bytes += (int64_t)( occBytes + rawGeo + rawAtt );
    const int gbd[3] = {fy->geometry_bitdepth_in, fy->geometry_bitdepth_out, fy->geometry_msb_align};
    const int obd[2] = {fy->occupancy_bitdepth_out, fy->occupancy_msb_align};
    int r = rb_ingest_yuv420_impl( c, fy->geometry_sample_bytes, fy->attribute_sample_bytes, fy->attribute_bitdepth, fy->upsampling_filter,
                                   fy->geometry_shift, fy->attribute_shift, gbd, obd );
    if ( r ) { return r; }
  }
  RB_CUDA( c->d_patches.ensure( std::max<size_t>( 1, nPatches ) * sizeof( RbPatch ) ) );
  RB_CUDA( c->d_wi_patch.ensure( std::max<int64_t>( 1, c->nWI ) * 4 ) );
  RB_CUDA( c->d_wi_local.ensure( std::max<int64_t>( 1, c->nWI ) * 4 ) );
  RB_CUDA( c->d_wi_count.ensure( ( c->nWI + 1 ) * 4 ) );
  RB_CUDA( c->d_wi_base.ensure( ( c->nWI + 1 ) * 8 ) );
  RB_CUDA( c->d_frame_wi_off.ensure( ( F + 1 ) * 4 ) );
  // the small tables go through one pinned staging block so the copies are truly asynchronous
  const size_t tb = nPatches * sizeof( RbPatch ) + c->nWI * 8 + ( F + 1 ) * 4;
  char*        st = (char*)rb_pinned( c, tb + 64 );
  if ( !st ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned staging allocation failed" ); }
  RB_CUDA( cudaStreamSynchronize( c->stream ) );  // staging block may still be in flight from a previous GOF
  size_t o = 0;
  if ( nPatches ) { memcpy( st + o, dp.data(), nPatches * sizeof( RbPatch ) ); }
  RB_CUDA( cudaMemcpyAsync( c->d_patches.p, st + o, nPatches * sizeof( RbPatch ), cudaMemcpyHostToDevice, c->stream ) );
  o += nPatches * sizeof( RbPatch );
  if ( c->nWI ) {
    memcpy( st + o, wiPatch.data(), c->nWI * 4 );
    RB_CUDA( cudaMemcpyAsync( c->d_wi_patch.p, st + o, c->nWI * 4, cudaMemcpyHostToDevice, c->stream ) );
    o += c->nWI * 4;
    memcpy( st + o, wiLocal.data(), c->nWI * 4 );
    RB_CUDA( cudaMemcpyAsync( c->d_wi_local.p, st + o, c->nWI * 4, cudaMemcpyHostToDevice, c->stream ) );
    o += c->nWI * 4;
  }
  memcpy( st + o, frameWiOff.data(), ( F + 1 ) * 4 );
  RB_CUDA( cudaMemcpyAsync( c->d_frame_wi_off.p, st + o, ( F + 1 ) * 4, cudaMemcpyHostToDevice, c->stream ) );
  c->stats.h2d_bytes += (int64_t)tb;
  c->uploaded      = true;
  c->pos_pre_valid = false;
  c->have_plr      = false;  // rb200_gof_set_plr follows the upload
  c->reconstructed = c->geo_smoothed = c->colors_transferred = c->color_smoothed = c->rgb_done = false;
  return RB200_OK;
}

int rb200_gof_set_plr( rb200_ctx* c, const rb200_plr* plr ) {
  if ( !c || !plr || !plr->modes || !plr->block_mode || !plr->block_offset || plr->n_modes < 1 || plr->n_modes > 256 ) {
    return rb_fail( c, RB200_ERR_INVALID, "gof_set_plr: bad arguments" );
  }
  if ( !c->uploaded ) { return rb_fail( c, RB200_ERR_STATE, "gof_set_plr: call rb200_gof_upload first" ); }
  cudaSetDevice( c->device );
  const size_t np = c->h_patches.size();
  for ( size_t i = 0; i < np; i++ ) {
    const int64_t nb = (int64_t)c->h_patches[i].size_u0 * c->h_patches[i].size_v0;
    if ( plr->block_offset[i] < 0 || plr->block_offset[i + 1] - plr->block_offset[i] < nb ) {
      return rb_fail( c, RB200_ERR_INVALID, "gof_set_plr: patch %zu has %lld blocks but %lld modes", i, (long long)nb,
                      (long long)( plr->block_offset[i + 1] - plr->block_offset[i] ) );
    }
  }
  const int64_t total = plr->block_offset[np];
  for ( int64_t i = 0; i < total; i++ ) {
    if ( plr->block_mode[i] >= plr->n_modes ) { return rb_fail( c, RB200_ERR_INVALID, "gof_set_plr: block mode %d out of range", plr->block_mode[i] ); }
  }
  RB_CUDA( c->d_plr_modes.ensure( 256 * sizeof( rb200_plr_mode ) ) );
  RB_CUDA( c->d_plr_block_mode.ensure( (size_t)std::max<int64_t>( total, 1 ) ) );
  RB_CUDA( c->d_plr_block_off.ensure( ( np + 1 ) * 8 ) );
  RB_CUDA( cudaMemcpyAsync( c->d_plr_modes.p, plr->modes, plr->n_modes * sizeof( rb200_plr_mode ), cudaMemcpyHostToDevice, c->stream ) );
  if ( total > 0 ) { RB_CUDA( cudaMemcpyAsync( c->d_plr_block_mode.p, plr->block_mode, (size_t)total, cudaMemcpyHostToDevice, c->stream ) ); }
  RB_CUDA( cudaMemcpyAsync( c->d_plr_block_off.p, plr->block_offset, ( np + 1 ) * 8, cudaMemcpyHostToDevice, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );  // the caller's arrays are pageable and may go away
  c->stats.h2d_bytes += (int64_t)( total + ( np + 1 ) * 8 + plr->n_modes * sizeof( rb200_plr_mode ) );
  c->have_plr = true;
  return RB200_OK;
}

int rb200_gof_upload( rb200_ctx* c, const rb200_frames* fr, const rb200_atlas* at ) {
  if ( !fr ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload: null argument" ); }
  return gof_upload_common( c, fr, nullptr, at );
}
int rb200_gof_upload_yuv420( rb200_ctx* c, const rb200_frames_yuv420* fy, const rb200_atlas* at ) {
  if ( !fy ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload_yuv420: null argument" ); }
  return gof_upload_common( c, nullptr, fy, at );
}
int rb200_gof_upload_nv12( rb200_ctx* c, const rb200_frames_nv12* fn, const rb200_atlas* at ) {
  if ( !c || !fn || !fn->occupancy || !fn->geometry ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload_nv12: null argument" ); }
  if ( !c->have_gof ) { return rb_fail( c, RB200_ERR_STATE, "gof_upload before gof_begin" ); }
  if ( fn->sample_bytes != 1 && fn->sample_bytes != 2 ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload_nv12: sample_bytes must be 1 or 2" ); }
  if ( fn->sample_lsb_shift < 0 || fn->sample_lsb_shift > 15 || ( fn->sample_bytes == 1 && fn->sample_lsb_shift ) ) {
    return rb_fail( c, RB200_ERR_INVALID, "gof_upload_nv12: sample_lsb_shift out of range" );
  }
  if ( c->P.attribute_count > 0 && !fn->attribute ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload_nv12: missing attribute surfaces" ); }
  const int nGA = c->F * c->M;
  for ( int i = 0; i < c->F; i++ ) {
    if ( !fn->occupancy[i].luma || fn->occupancy[i].pitch_luma < c->oW ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload_nv12: bad occupancy surface %d", i ); }
  }
  for ( int i = 0; i < nGA; i++ ) {
    if ( !fn->geometry[i].luma || fn->geometry[i].pitch_luma < c->W * fn->sample_bytes ) { return rb_fail( c, RB200_ERR_INVALID, "gof_upload_nv12: bad geometry surface %d", i ); }
    if ( c->P.attribute_count > 0 && ( !fn->attribute[i].luma || !fn->attribute[i].chroma || fn->attribute[i].pitch_luma < c->W * fn->sample_bytes ||
                                       fn->attribute[i].pitch_chroma < c->W * fn->sample_bytes ) ) {
      return rb_fail( c, RB200_ERR_INVALID, "gof_upload_nv12: bad attribute surface %d", i );
    }
  }
  rb200_frames_yuv420 fy    = fn->conversion;
  fy.geometry_sample_bytes  = fn->sample_bytes;
  fy.attribute_sample_bytes = fn->sample_bytes;
  return gof_upload_common( c, nullptr, &fy, at, fn );
}
// the planes the reconstruction reads, as they are in HBM after an upload (tests of the ingest conversion)
int rb200_download_planes( rb200_ctx* c, int frame, int map, uint16_t* geometry, uint16_t* attribute ) {
  if ( !c || !c->uploaded || frame < 0 || frame >= c->F || map < 0 || map >= c->M ) { return rb_fail( c, RB200_ERR_INVALID, "download_planes: bad frame / map or nothing uploaded" ); }
  cudaSetDevice( c->device );
  const size_t plane = (size_t)c->W * c->H, fm = (size_t)frame * c->M + map;
  if ( geometry ) { RB_CUDA( cudaMemcpyAsync( geometry, c->d_geometry.as<uint16_t>() + fm * plane, plane * 2, cudaMemcpyDeviceToHost, c->stream ) ); }
  if ( attribute && c->P.attribute_count > 0 ) {
    RB_CUDA( cudaMemcpyAsync( attribute, c->d_attribute.as<uint16_t>() + fm * 3 * plane, plane * 6, cudaMemcpyDeviceToHost, c->stream ) );
  }
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  return RB200_OK;
}

static int take_snapshot( rb200_ctx* c, RbBuf& dst, bool& have, const RbBuf& src ) {
  const size_t bytes = (size_t)c->h_frame_off[c->F] * 8;
  if ( !c->snapshots || bytes == 0 ) { return RB200_OK; }
  RB_CUDA( dst.ensure( bytes ) );
  RB_CUDA( cudaMemcpyAsync( dst.p, src.p, bytes, cudaMemcpyDeviceToDevice, c->stream ) );
  have = true;
  return RB200_OK;
}

int rb200_reconstruct( rb200_ctx* c ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  if ( !c->uploaded ) { return rb_fail( c, RB200_ERR_STATE, "reconstruct before gof_upload" ); }
  cudaSetDevice( c->device );
  c->pos_pre_valid = false;
  int r = rb_reconstruct_impl( c );
  if ( r == RB200_OK ) {
    c->reconstructed = true;
    c->geo_smoothed = c->colors_transferred = c->color_smoothed = c->rgb_done = false;
    c->have_snap_pos[0] = c->have_snap_pos[1] = c->have_snap_col[0] = c->have_snap_col[1] = c->have_snap_col[2] = false;
    r = take_snapshot( c, c->d_snap_pos[0], c->have_snap_pos[0], c->d_pos );
    if ( r == RB200_OK ) { r = take_snapshot( c, c->d_snap_col[0], c->have_snap_col[0], c->d_col ); }
  }
  return r;
}

int rb200_smooth_geometry( rb200_ctx* c ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  if ( !c->reconstructed ) { return rb_fail( c, RB200_ERR_STATE, "smooth_geometry before reconstruct" ); }
  cudaSetDevice( c->device );
  int r = rb_smooth_geometry_impl( c );
  if ( r == RB200_OK ) {
    c->geo_smoothed = true;
    r               = take_snapshot( c, c->d_snap_pos[1], c->have_snap_pos[1], c->d_pos );
  }
  return r;
}

int rb200_transfer_colors( rb200_ctx* c ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  if ( !c->geo_smoothed ) { return rb_fail( c, RB200_ERR_STATE, "transfer_colors before smooth_geometry" ); }
  cudaSetDevice( c->device );
  int r = rb_transfer_colors_impl( c );
  if ( r == RB200_OK ) {
    c->colors_transferred = true;
    r                     = take_snapshot( c, c->d_snap_col[1], c->have_snap_col[1], c->d_col );
  }
  return r;
}

int rb200_smooth_color( rb200_ctx* c ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  if ( !c->reconstructed ) { return rb_fail( c, RB200_ERR_STATE, "smooth_color before reconstruct" ); }
  cudaSetDevice( c->device );
  int r = rb_smooth_color_impl( c );
  if ( r == RB200_OK ) {
    c->color_smoothed = true;
    r                 = take_snapshot( c, c->d_snap_col[2], c->have_snap_col[2], c->d_col );
  }
  return r;
}

int rb200_convert_rgb8( rb200_ctx* c ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  if ( !c->reconstructed ) { return rb_fail( c, RB200_ERR_STATE, "convert_rgb8 before reconstruct" ); }
  cudaSetDevice( c->device );
  int r = rb_convert_rgb8_impl( c );
  if ( r == RB200_OK ) { c->rgb_done = true; }
  return r;
}

int rb200_debug_yuv16_to_rgb8( rb200_ctx* c, const uint16_t* yuv, int64_t n, uint8_t* rgb, int force_f64 ) {
  if ( !c || !yuv || !rgb || n < 0 ) { return RB200_ERR_INVALID; }
  if ( n == 0 ) { return RB200_OK; }
  cudaSetDevice( c->device );
  return rb_debug_rgb8_impl( c, yuv, n, rgb, force_f64 );
}

// The decoder's per-frame sequence, PCCDecoder.cpp:330-508, for the whole GOF.
int rb200_decode_gof( rb200_ctx* c ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  int r = rb200_reconstruct( c );
  if ( r ) { return r; }
  const rb200_params& p = c->P;
  if ( p.apply_geo_smoothing && p.flag_geometry_smoothing ) {  // :434
    if ( p.grid_smoothing || p.neighbor_count_smoothing > 0 ) {  // (> 0 without the grid: the encoder-side call, rabbit_b200.h)
      r = rb200_smooth_geometry( c );  // :437
      if ( r ) { return r; }
    }
    if ( p.attribute_count > 0 && p.attr_transfer_filter_type != 0 ) {  // :439-465
      if ( p.attr_transfer_filter_type != 1 ) {
        return rb_fail( c, RB200_ERR_UNSUPPORTED, "attrTransferFilterType %d not implemented (only 0 and 1)",
                        p.attr_transfer_filter_type );
      }
      // gridSmoothing_ == 0: tempFrameBuffer == reconstruct and no point is of type 3, so transferColors16bitBP
      // changes nothing (PCCPointSet.cpp:1163-1164) and is skipped
      if ( p.grid_smoothing && !p.pbf_enable ) {  // :446
        r = rb200_transfer_colors( c );
        if ( r ) { return r; }
      }
    }
  }
  if ( p.attribute_count > 0 ) {
    if ( p.apply_attr_smoothing && p.flag_color_smoothing ) {  // :496-499
      r = rb200_smooth_color( c );
      if ( r ) { return r; }
    }
    r = rb200_convert_rgb8( c );  // :500-507
    if ( r ) { return r; }
  }
  return RB200_OK;
}

int rb200_frame_counts_get( rb200_ctx* c, rb200_frame_counts* out ) {
  if ( !c || !out ) { return RB200_ERR_INVALID; }
  if ( !c->reconstructed ) { return rb_fail( c, RB200_ERR_STATE, "frame_counts before reconstruct" ); }
  cudaSetDevice( c->device );
  RbFrameInfo* hi = (RbFrameInfo*)rb_pinned( c, c->F * sizeof( RbFrameInfo ) );
  if ( !hi ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  RB_CUDA( cudaMemcpyAsync( hi, c->d_frame_info.p, c->F * sizeof( RbFrameInfo ), cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  c->stats.d2h_bytes += c->F * sizeof( RbFrameInfo );
  for ( int f = 0; f < c->F; f++ ) {
    c->h_counts[f].smoothed  = hi[f].smoothed;
    c->h_counts[f].recolored = hi[f].recolored;
    out[f]                   = c->h_counts[f];
  }
  return RB200_OK;
}

// packs the points [b, b+n) of the GOF arena into the reference's std::vector layouts and copies them out;
// every requested field gets its own slice of the staging buffer so one synchronisation covers them all
static int download_range( rb200_ctx* c, int64_t b, int64_t n, const rb200_cloud_host* dst, const short4* srcPos = nullptr,
                           const ushort4* srcCol = nullptr ) {
  if ( n == 0 ) { return RB200_OK; }
  if ( !srcPos ) { srcPos = c->d_pos.as<short4>(); }
  if ( !srcCol ) { srcCol = c->d_col.as<ushort4>(); }
  auto         al = []( size_t x ) { return ( x + 255 ) & ~size_t( 255 ); };
  const size_t oPos = 0, oTyp = oPos + ( dst->positions ? al( n * 6 ) : 0 ), oC16 = oTyp + ( dst->boundary_types ? al( n * 2 ) : 0 ),
               oRgb = oC16 + ( dst->colors16 ? al( n * 6 ) : 0 ), oPix = oRgb + ( dst->colors && c->rgb_done ? al( n * 3 ) : 0 ),
               total = oPix + ( dst->point_to_pixel ? al( n * 12 ) : 0 );
  RB_CUDA( c->d_pack.ensure( total + 256 ) );
  char*     pk = c->d_pack.as<char>();
  const int T = 256, G = rb_div_up( n, T );
  if ( dst->positions ) {
    RB_LAUNCH( "pack_positions", k_pack_positions, G, T, 0, srcPos + b, n, (int16_t*)( pk + oPos ) );
    RB_CUDA( cudaMemcpyAsync( dst->positions, pk + oPos, n * 6, cudaMemcpyDeviceToHost, c->stream ) );
    c->stats.d2h_bytes += n * 6;
  }
  if ( dst->boundary_types ) {
    RB_LAUNCH( "pack_types", k_pack_types, G, T, 0, srcPos + b, n, (uint16_t*)( pk + oTyp ) );
    RB_CUDA( cudaMemcpyAsync( dst->boundary_types, pk + oTyp, n * 2, cudaMemcpyDeviceToHost, c->stream ) );
    c->stats.d2h_bytes += n * 2;
  }
  if ( dst->colors16 ) {
    RB_LAUNCH( "pack_colors16", k_pack_colors16, G, T, 0, srcCol + b, n, (uint16_t*)( pk + oC16 ) );
    RB_CUDA( cudaMemcpyAsync( dst->colors16, pk + oC16, n * 6, cudaMemcpyDeviceToHost, c->stream ) );
    c->stats.d2h_bytes += n * 6;
  }
  if ( dst->colors ) {
    if ( c->rgb_done ) {
      RB_LAUNCH( "pack_rgb", k_pack_rgb, G, T, 0, c->d_rgb.as<uchar4>() + b, n, (uint8_t*)( pk + oRgb ) );
      RB_CUDA( cudaMemcpyAsync( dst->colors, pk + oRgb, n * 3, cudaMemcpyDeviceToHost, c->stream ) );
      c->stats.d2h_bytes += n * 3;
    } else {
      memset( dst->colors, 0, n * 3 );  // colorPointCloud's fillColor(0), PCCCodec.cpp:1319
    }
  }
  if ( dst->partition ) {
    RB_CUDA( cudaMemcpyAsync( dst->partition, c->d_part.as<uint32_t>() + b, n * 4, cudaMemcpyDeviceToHost, c->stream ) );
    c->stats.d2h_bytes += n * 4;
  }
  if ( dst->point_to_pixel ) {
    RB_LAUNCH( "pack_pixels", k_pack_pixels, G, T, 0, c->d_pix.as<uint32_t>() + b, c->d_col.as<ushort4>() + b, n,
               (uint32_t*)( pk + oPix ) );
    RB_CUDA( cudaMemcpyAsync( dst->point_to_pixel, pk + oPix, n * 12, cudaMemcpyDeviceToHost, c->stream ) );
    c->stats.d2h_bytes += n * 12;
  }
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  return RB200_OK;
}

int rb200_download_frame( rb200_ctx* c, int f, const rb200_cloud_host* dst ) {
  if ( !c || !dst ) { return RB200_ERR_INVALID; }
  if ( !c->reconstructed ) { return rb_fail( c, RB200_ERR_STATE, "download before reconstruct" ); }
  if ( f < 0 || f >= c->F ) { return rb_fail( c, RB200_ERR_INVALID, "frame index out of range" ); }
  cudaSetDevice( c->device );
  return download_range( c, c->h_frame_off[f], c->h_frame_off[f + 1] - c->h_frame_off[f], dst );
}

int rb200_enable_stage_snapshots( rb200_ctx* c, int enable ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  c->snapshots = enable != 0;
  return RB200_OK;
}

// stage: 0 after reconstruction (+ colour fetch), 1 after geometry smoothing, 2 after the attribute re-transfer,
// 3 after colour smoothing, 4 (or more) current state.  Needs rb200_enable_stage_snapshots before the GOF is decoded.
int rb200_download_frame_stage( rb200_ctx* c, int f, int stage, const rb200_cloud_host* dst ) {
  if ( !c || !dst ) { return RB200_ERR_INVALID; }
  if ( !c->reconstructed ) { return rb_fail( c, RB200_ERR_STATE, "download before reconstruct" ); }
  if ( f < 0 || f >= c->F ) { return rb_fail( c, RB200_ERR_INVALID, "frame index out of range" ); }
  if ( stage < 4 && !c->snapshots ) { return rb_fail( c, RB200_ERR_STATE, "stage snapshots are not enabled" ); }
  cudaSetDevice( c->device );
  const short4*  sp = nullptr;
  const ushort4* sc = nullptr;
  if ( stage < 4 ) {
    // the state after `stage` is the LATEST snapshot taken at or before it (a stage that did not run changes nothing)
    if ( stage >= 1 && c->have_snap_pos[1] ) {
      sp = c->d_snap_pos[1].as<short4>();
    } else if ( c->have_snap_pos[0] ) {
      sp = c->d_snap_pos[0].as<short4>();
    }
    if ( stage >= 3 && c->have_snap_col[2] ) {
      sc = c->d_snap_col[2].as<ushort4>();
    } else if ( stage >= 2 && c->have_snap_col[1] ) {
      sc = c->d_snap_col[1].as<ushort4>();
    } else if ( c->have_snap_col[0] ) {
      sc = c->d_snap_col[0].as<ushort4>();
    }
  }
  return download_range( c, c->h_frame_off[f], c->h_frame_off[f + 1] - c->h_frame_off[f], dst, sp, sc );
}

int rb200_download_gof( rb200_ctx* c, const rb200_cloud_host* dst ) {
  if ( !c || !dst ) { return RB200_ERR_INVALID; }
  if ( !c->reconstructed ) { return rb_fail( c, RB200_ERR_STATE, "download before reconstruct" ); }
  cudaSetDevice( c->device );
  return download_range( c, 0, c->h_frame_off[c->F], dst );
}

int rb200_download_block_to_patch( rb200_ctx* c, int f, uint32_t* dst ) {
  if ( !c || !dst ) { return RB200_ERR_INVALID; }
  if ( !c->reconstructed ) { return rb_fail( c, RB200_ERR_STATE, "download before reconstruct" ); }
  if ( f < 0 || f >= c->F ) { return rb_fail( c, RB200_ERR_INVALID, "frame index out of range" ); }
  cudaSetDevice( c->device );
  const size_t n = (size_t)c->Wb * c->Hb;
  RB_CUDA( cudaMemcpyAsync( dst, c->d_b2p.as<uint32_t>() + f * n, n * 4, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  c->stats.d2h_bytes += n * 4;
  return RB200_OK;
}

int rb200_download_occupancy( rb200_ctx* c, int f, uint8_t* dst ) {
  if ( !c || !dst ) { return RB200_ERR_INVALID; }
  if ( !c->reconstructed ) { return rb_fail( c, RB200_ERR_STATE, "download before reconstruct" ); }
  if ( f < 0 || f >= c->F ) { return rb_fail( c, RB200_ERR_INVALID, "frame index out of range" ); }
  cudaSetDevice( c->device );
  const size_t n = (size_t)c->W * c->H;
  RB_CUDA( c->d_pack.ensure( n ) );
  RB_LAUNCH( "expand_bitmap", k_expand_bitmap, rb_div_up( n, 256 ), 256, 0,
             c->d_bitmap.as<uint32_t>() + (size_t)f * c->H * c->bmWords, c->W, c->H, c->bmWords, c->d_pack.as<uint8_t>() );
  RB_CUDA( cudaMemcpyAsync( dst, c->d_pack.p, n, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  c->stats.d2h_bytes += n;
  return RB200_OK;
}

int rb200_debug_set_grid_shrink( rb200_ctx* c, int shrink ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  c->test_grid_shrink = std::max( 0, std::min( 16, shrink ) );
  return RB200_OK;
}

int rb200_stats_get( rb200_ctx* c, rb200_launch_stats* out, int reset ) {
  if ( !c || !out ) { return RB200_ERR_INVALID; }
  *out = c->stats;
  if ( reset ) { c->stats = rb200_launch_stats{}; }
  return RB200_OK;
}

int rb200_timing_enable( rb200_ctx* c, int enable ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  cudaSetDevice( c->device );
  rb_timing_resolve( c );
  c->timing = enable != 0;
  if ( enable ) {
    c->timing_names.clear();
    c->timing_ms.clear();
    c->timing_n.clear();
  }
  return RB200_OK;
}

int rb200_timing_get( rb200_ctx* c, int index, char* name, int cap, double* ms, int64_t* n ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  cudaSetDevice( c->device );
  rb_timing_resolve( c );
  if ( index < 0 || index >= (int)c->timing_names.size() ) { return RB200_ERR_INVALID; }
  if ( name && cap > 0 ) {
    strncpy( name, c->timing_names[index].c_str(), cap - 1 );
    name[cap - 1] = 0;
  }
  if ( ms ) { *ms = c->timing_ms[index]; }
  if ( n ) { *n = c->timing_n[index]; }
  return RB200_OK;
}

}  // extern "C"
