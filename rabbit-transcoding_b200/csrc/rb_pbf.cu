// rb_pbf.cu — Rec-2 occupancy synthesis ("patch border filtering", PBF) for every frame of the GOF (sm_100a).
//
// Restates PatchBlockFiltering::patchBorderFiltering and the PCCPatch helpers it drives
// (PccLibCommon/source/PCCPatch.cpp:797-977: setLocalData, generateBorderPoints3D, filtering; isBorder :858-867) as
// generatePointCloud runs them when the occupancy-synthesis SEI is present (PccLibCommon/source/PCCCodec.cpp:541-554),
// and the two places the result is read: the per-pixel occupancy of the main loop (:662-664) and the boundary type of
// the created points (:664, :809).  Everything happens in PATCH-LOCAL space — a patch only sees the blocks it owns,
// surrounded by a border of `border_` empty pixels — so every patch of the GOF gets a local map
//   w = sizeU0 * 16 + 2 border, h = sizeV0 * 16 + 2 border
// in one arena: occupancy (two copies, the passes ping-pong), depth, and per pixel the best neighbouring depth.
//   k_pbf_local     setLocalData: occupancy = video > threshold at the occupancy precision, depth = geometry map 0
//   k_pbf_border    generateBorderPoints3D: occupied pixels with an empty pixel among 12 neighbours -> flag + the 3-D
//                   bounding box of the patch's border points
//   k_pbf_neighbors the first half of filtering(): border points of the patches whose boxes intersect, inside this
//                   patch's box grown by 8, projected into this patch; per pixel the depth closest to the patch's own
//                   (ties: the first in patch order, then scan order — the order the reference walks them in) wins.
//                   One 64-bit atomicMin per candidate: | |d - depth| : 8 | patch : 16 | scan position : 24 | d : 16 |
//   k_pbf_filter    one pass of the second half: a pixel with 1..3 occupied 4-neighbours stays when no neighbouring
//                   depth is in its window or the summed distances say the surface continues (sumE >= sumP)
//   k_pbf_scatter   the final local maps back to canvas space: the occupancy bitmap the reprojection kernels read, and a
//                   second bitmap with isBorder() (any empty pixel in the 5x5 window) = the points' boundary type
// The depth of a point is unchanged by all this (getDepthMap( u, v ) is the geometry sample of the pixel), so the
// reprojection kernels run as they are on the two bitmaps.
#include "rb_common.cuh"

namespace {

constexpr int TPB = 256;

struct PbfPatch {
  int64_t off;   // first element of the patch's local maps in the arena
  int32_t w, h;  // local map size including the border
};

struct PbfArgs {
  const RbPatch*  patches;
  const PbfPatch* local;
  const int32_t*  patch_off;  // [F + 1] first patch of every frame
  int             nPatches, F;
  int64_t         total;      // arena elements
  int             border, W, H, oW, oH, M, prec, Wb, Hb, bmWords, threshold;
  const uint8_t*  occ_video;
  const uint16_t* geometry;
  const uint32_t* b2p;
  uint8_t *       occA, *occB, *flag;
  int16_t*        depth;
  unsigned long long* key;
  int32_t*        box;  // [nPatches][6] min xyz, max xyz of the border points
  int             dist2, filterSize;
  uint32_t *      bitmap, *bnd;
};

constexpr unsigned long long KEY_NONE = ~0ull;

// g_orientation (PCCPatch.cpp:40-47): direction of the surface normal in the 2-D occupancy pattern of the 8 neighbours
// (index = the 8 neighbours as bits tl t tr l r bl b br, :50-54; 40 of the 256 patterns have a direction)
__constant__ uint8_t c_orientation[256] = {
    0, 0, 6, 0, 0, 0, 0, 6, 4, 0, 0, 5, 0, 0, 0, 5,  //   0..
    0, 0, 0, 0, 0, 0, 7, 7, 0, 0, 0, 0, 0, 0, 0, 6,  //  16..
    0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 0, 5, 0, 0, 0, 5,  //  32..
    0, 0, 0, 0, 0, 0, 7, 0, 0, 0, 0, 0, 0, 0, 0, 5,  //  48..
    2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  //  64..
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  //  80..
    0, 0, 0, 0, 0, 0, 0, 0, 3, 3, 0, 4, 3, 0, 0, 5,  //  96..
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  // 112..
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 5, 0, 0, 0, 0,  // 128..
    0, 0, 0, 0, 0, 0, 7, 7, 0, 0, 0, 0, 0, 0, 0, 7,  // 144..
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  // 160..
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6,  // 176..
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  // 192..
    1, 1, 0, 0, 1, 0, 0, 7, 0, 0, 0, 0, 0, 0, 0, 0,  // 208..
    2, 0, 0, 0, 0, 0, 0, 0, 3, 3, 0, 3, 0, 0, 0, 4,  // 224..
    1, 0, 0, 0, 1, 0, 1, 0, 2, 3, 0, 0, 1, 2, 0, 0};  // 240..
// g_dilate (PCCPatch.cpp:48)
__constant__ int8_t c_dilate[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};

// element -> ( patch, local position ): binary search over the patches' arena offsets
__device__ __forceinline__ int patch_of( const PbfPatch* __restrict__ local, int n, int64_t e ) {
  int lo = 0, hi = n - 1;
  while ( lo < hi ) {
    const int mid = ( lo + hi + 1 ) >> 1;
    if ( local[mid].off <= e ) {
      lo = mid;
    } else {
      hi = mid - 1;
    }
  }
  return lo;
}

// PCCPatch::patch2Canvas (PCCPatch.cpp:192-250) for a pixel of the patch
__device__ __forceinline__ void patch_to_canvas( const RbPatch& p, int u, int v, int& x, int& y ) {
  const int su = p.su0 * 16, sv = p.sv0 * 16, x0 = p.u0 * 16, y0 = p.v0 * 16;
  switch ( p.orient ) {
    default:
    case 0: x = u + x0; y = v + y0; break;
    case 1: x = v + x0; y = u + y0; break;
    case 2: x = ( sv - 1 - v ) + x0; y = u + y0; break;
    case 3: x = ( su - 1 - u ) + x0; y = ( sv - 1 - v ) + y0; break;
    case 4: x = v + x0; y = ( su - 1 - u ) + y0; break;
    case 5: x = ( su - 1 - u ) + x0; y = v + y0; break;
    case 6: x = ( sv - 1 - v ) + x0; y = ( su - 1 - u ) + y0; break;
    case 7: x = u + x0; y = ( sv - 1 - v ) + y0; break;
    case 8: x = v + x0; y = u + y0; break;
  }
}

// does the patch own the block of its pixel ( u, v )?  (blockToPatch == index + 1, PCCPatch.cpp:820-822)
__device__ __forceinline__ bool owns_pixel( const PbfArgs& a, const RbPatch& p, int x, int y ) {
  return a.b2p[( (size_t)p.frame * a.Hb + ( y >> 4 ) ) * a.Wb + ( x >> 4 )] == (uint32_t)p.frame_patch + 1u;
}

// PCCPatch::generatePoint (PCCPatch.h:201-207; level of detail 1, no additional plane: checked on the host)
__device__ __forceinline__ void pbf_point( const RbPatch& p, int u, int v, int depth, int ( &P )[3] ) {
  const int d  = (uint16_t)depth;  // generatePoint takes a uint16_t
  const int nn = p.mode == 0 ? d + p.d1 : max( p.d1 - d, 0 );
  P[p.tangent_axis]   = (int16_t)( u + p.u1 );
  P[p.bitangent_axis] = (int16_t)( v + p.v1 );
  P[p.normal_axis]    = (int16_t)nn;
}

// setLocalData (:797-855): one thread per element of the arena
__global__ void __launch_bounds__( TPB ) k_pbf_local( const PbfArgs a ) {
  const int64_t e = blockIdx.x * (int64_t)TPB + threadIdx.x;
  if ( e >= a.total ) { return; }
  const int      g  = patch_of( a.local, a.nPatches, e );
  const PbfPatch L  = a.local[g];
  const RbPatch  p  = a.patches[g];
  const int      cl = (int)( e - L.off ), u = cl % L.w - a.border, v = cl / L.w - a.border;
  uint8_t        occ = 0;
  int16_t        d   = 0;
  if ( u >= 0 && v >= 0 && u < p.su0 * 16 && v < p.sv0 * 16 ) {
    int x, y;
    patch_to_canvas( p, u, v, x, y );
    if ( owns_pixel( a, p, x, y ) &&
         (int)a.occ_video[( (size_t)p.frame * a.oH + y / a.prec ) * a.oW + x / a.prec] > a.threshold ) {
      occ = 1;
      d   = (int16_t)a.geometry[( (size_t)p.frame * a.M * a.H + y ) * a.W + x];  // map 0 of the frame (:547)
    }
  }
  a.occA[e]  = occ;
  a.occB[e]  = 0;
  a.flag[e]  = 0;
  a.depth[e] = d;
  a.key[e]   = KEY_NONE;
  if ( e < (int64_t)a.nPatches * 6 ) { a.box[e] = ( e % 6 ) < 3 ? 32767 : -32768; }
}

// generateBorderPoints3D (:868-889)
__global__ void __launch_bounds__( TPB ) k_pbf_border( const PbfArgs a ) {
  const int64_t e = blockIdx.x * (int64_t)TPB + threadIdx.x;
  if ( e >= a.total ) { return; }
  const int      g  = patch_of( a.local, a.nPatches, e );
  const PbfPatch L  = a.local[g];
  const RbPatch  p  = a.patches[g];
  const int      cl = (int)( e - L.off ), u = cl % L.w - a.border, v = cl / L.w - a.border;
  if ( u < 0 || v < 0 || u >= p.su0 * 16 || v >= p.sv0 * 16 ) { return; }
  const uint8_t* o = a.occA + e;
  const int      w = L.w;
  if ( !o[0] ) { return; }
  if ( o[-1] && o[1] && o[-w] && o[w] && o[-2] && o[2] && o[-2 * w] && o[2 * w] && o[w - 1] && o[w + 1] && o[-w - 1] && o[-w + 1] ) {
    return;
  }
  a.flag[e] = 1;
  int P[3];
  pbf_point( p, u, v, a.depth[e], P );
  int32_t* bx = a.box + (size_t)g * 6;
  for ( int k = 0; k < 3; k++ ) {
    atomicMin( bx + k, P[k] );
    atomicMax( bx + 3 + k, P[k] );
  }
}

// filtering(), first loop (:902-913): one thread per border point, over the other patches of its frame
__global__ void __launch_bounds__( TPB ) k_pbf_neighbors( const PbfArgs a ) {
  const int64_t e = blockIdx.x * (int64_t)TPB + threadIdx.x;
  if ( e >= a.total || !a.flag[e] ) { return; }
  const int      g  = patch_of( a.local, a.nPatches, e );
  const PbfPatch L  = a.local[g];
  const RbPatch  p  = a.patches[g];
  const int      cl = (int)( e - L.off ), u = cl % L.w - a.border, v = cl / L.w - a.border;
  int            P[3];
  pbf_point( p, u, v, a.depth[e], P );
  const int32_t* mybox = a.box + (size_t)g * 6;
  const int      f = p.frame, g0 = a.patch_off[f], g1 = a.patch_off[f + 1];
  // position of the point among the border points of its patch = its scan position ( v, u ): (:872-873)
  const unsigned long long scanpos = (unsigned long long)v * (unsigned)( p.su0 * 16 ) + (unsigned)u;
  for ( int i = g0; i < g1; i++ ) {
    if ( i == g ) { continue; }
    const int32_t* bx = a.box + (size_t)i * 6;
    // neighboringPatches_ of patch i holds this patch when the boxes intersect (:969-975, PCCMath.h:266-269)
    if ( !( bx[3] >= mybox[0] && bx[0] <= mybox[3] && bx[4] >= mybox[1] && bx[1] <= mybox[4] && bx[5] >= mybox[2] && bx[2] <= mybox[5] ) ) {
      continue;
    }
    // boundingBox grown by 8 in int16 arithmetic (:899-900), contains (PCCMath.h:244-247)
    bool in = true;
    for ( int k = 0; k < 3; k++ ) {
      const int lo = (int16_t)( bx[k] - 8 ), hi = (int16_t)( bx[3 + k] + 8 );
      in           = in && !( P[k] < lo || P[k] > hi );
    }
    if ( !in ) { continue; }
    const RbPatch  q  = a.patches[i];
    const PbfPatch Lq = a.local[i];
    const int      d  = (int16_t)( q.mode == 0 ? P[q.normal_axis] - q.d1 : q.d1 - P[q.normal_axis] );  // generateDepth
    const int      cu = P[q.tangent_axis] - q.u1 + a.border, cv = P[q.bitangent_axis] - q.v1 + a.border;  // `shift`, :901
    if ( cu < 0 || cv < 0 || cu >= Lq.w || cv >= Lq.h ) { continue; }  // (cannot happen: the grown box lies inside the border)
    const int64_t c    = Lq.off + (int64_t)cv * Lq.w + cu;
    const int     diff = abs( d - (int)a.depth[c] );
    if ( diff <= a.dist2 ) {
      const unsigned long long k = ( (unsigned long long)diff << 56 ) | ( (unsigned long long)( g - g0 ) << 40 ) | ( scanpos << 16 ) |
                                   (unsigned long long)(uint16_t)d;
      atomicMin( a.key + c, k );
    }
  }
}

// filtering(), one pass of the second loop (:915-957)
__global__ void __launch_bounds__( TPB ) k_pbf_filter( const PbfArgs a, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst ) {
  const int64_t e = blockIdx.x * (int64_t)TPB + threadIdx.x;
  if ( e >= a.total ) { return; }
  const int      g  = patch_of( a.local, a.nPatches, e );
  const PbfPatch L  = a.local[g];
  const RbPatch  p  = a.patches[g];
  const int      cl = (int)( e - L.off ), u = cl % L.w - a.border, v = cl / L.w - a.border;
  if ( u < 0 || v < 0 || u >= p.su0 * 16 || v >= p.sv0 * 16 ) { return; }
  const int      w = L.w;
  const uint8_t* s = src + e;
  uint8_t        out;
  if ( s[0] == 0 ) {
    out = 0;
  } else {
    const int nn = s[-1] + s[1] + s[-w] + s[w];
    if ( nn == 0 ) {
      out = 0;
    } else if ( nn == 4 ) {
      out = 1;
    } else {
      const int pat = ( s[-w - 1] << 7 ) | ( s[-w] << 6 ) | ( s[-w + 1] << 5 ) | ( s[-1] << 4 ) | ( s[1] << 3 ) | ( s[w - 1] << 2 ) |
                      ( s[w] << 1 ) | s[w + 1];
      const int orX = c_orientation[pat], orY = ( orX + 2 ) % 8;
      const int dX0 = c_dilate[orX][0], dX1 = c_dilate[orX][1], dY0 = c_dilate[orY][0], dY1 = c_dilate[orY][1];
      const int shiftX = dX0 + dX1 * w, shiftY = dY0 + dY1 * w;
      const int16_t* dep = a.depth + e;
      const int      dE = dep[-shiftX], dP = dep[0];
      const int      wu = a.filterSize, wv = a.filterSize >> 1;
      float          sumE = 0.f, sumP = 0.f;
      int            count = 0;
      const int64_t  lo = L.off, hi = L.off + (int64_t)L.w * L.h;
      int64_t        cx  = e - (int64_t)wu * shiftX - (int64_t)wv * shiftY;
      int            du1 = -wu * dX0 - wv * dY0, dv1 = -wu * dX1 - wv * dY1;
      for ( int dx = -wu; dx <= wu; dx++, cx += shiftX, du1 += dX0, dv1 += dX1 ) {
        int64_t cn = cx;
        int     du = du1, dv = dv1;
        for ( int dy = -wv; dy <= wv; dy++, cn += shiftY, du += dY0, dv += dY1 ) {
          if ( cn < lo || cn >= hi ) { continue; }  // (the window stays inside the border for the supported filter sizes)
          const unsigned long long k = a.key[cn];
          if ( k != KEY_NONE ) {
            const int nd = (int16_t)( k & 0xFFFFu );
            // sqrt( int ) is the double overload; the sums are float (:934-937)
            sumP = (float)( (double)sumP + sqrt( (double)( du * du + dv * dv + ( nd - dP ) * ( nd - dP ) ) ) );
            sumE = (float)( (double)sumE +
                            sqrt( (double)( ( du + dX0 ) * ( du + dX0 ) + ( dv + dX1 ) * ( dv + dX1 ) + ( nd - dE ) * ( nd - dE ) ) ) );
            count++;
          }
        }
      }
      out = ( count == 0 || sumE >= sumP ) ? 1 : 0;
    }
  }
  dst[e] = out;
}

// the final local occupancy -> canvas bitmaps (zeroed before): occupancy (:662-663) and isBorder (:858-867, :664)
__global__ void __launch_bounds__( TPB ) k_pbf_scatter( const PbfArgs a, const uint8_t* __restrict__ occ ) {
  const int64_t e = blockIdx.x * (int64_t)TPB + threadIdx.x;
  if ( e >= a.total || !occ[e] ) { return; }
  const int      g  = patch_of( a.local, a.nPatches, e );
  const PbfPatch L  = a.local[g];
  const RbPatch  p  = a.patches[g];
  const int      cl = (int)( e - L.off ), u = cl % L.w - a.border, v = cl / L.w - a.border;
  if ( u < 0 || v < 0 || u >= p.su0 * 16 || v >= p.sv0 * 16 ) { return; }
  int x, y;
  patch_to_canvas( p, u, v, x, y );
  if ( !owns_pixel( a, p, x, y ) ) { return; }
  const size_t   word = ( (size_t)p.frame * a.H + y ) * a.bmWords + ( x >> 5 );
  const uint32_t bit  = 1u << ( x & 31 );
  atomicOr( a.bitmap + word, bit );
  bool border = false;
  for ( int dy = -2; dy <= 2; dy++ ) {
    for ( int dx = -2; dx <= 2; dx++ ) { border = border || occ[e + (int64_t)dy * L.w + dx] == 0; }
  }
  if ( border ) { atomicOr( a.bnd + word, bit ); }
}

}  // namespace

// Called by rb_reconstruct_impl between block-to-patch and the reprojection kernels when P.pbf_enable is set:
// replaces the occupancy bitmap by the synthesised one and fills c->d_bnd_bitmap.
int rb_pbf_impl( rb200_ctx* c ) {
  const rb200_params& P = c->P;
  const int F = c->F, nPatches = (int)c->h_patches.size();
  if ( P.pbf_passes_count < 1 || P.pbf_passes_count > 16 || P.pbf_filter_size < 1 || P.pbf_log2_threshold < 1 ) {
    return rb_fail( c, RB200_ERR_INVALID, "pbf: passes %d, filter size %d, log2 threshold %d", P.pbf_passes_count, P.pbf_filter_size,
                    P.pbf_log2_threshold );
  }
  const int border = c->prec >= 8 ? 16 : 8;  // :804
  const int dist2  = P.pbf_log2_threshold * P.pbf_log2_threshold;
  // the filter window reaches filterSize * ( 1, 1 ) + ( filterSize / 2 ) * ( 1, 1 ) + 1 pixels from a patch pixel
  if ( P.pbf_filter_size + ( P.pbf_filter_size >> 1 ) + 1 > border || dist2 > 255 ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "pbf: filter size %d / threshold %d reach outside the patch border of %d (the "
                    "reference reads outside its maps there)", P.pbf_filter_size, dist2, border );
  }
  const size_t bmBytes = (size_t)F * c->H * c->bmWords * 4;
  RB_CUDA( c->d_bnd_bitmap.ensure( bmBytes ) );
  RB_CUDA( cudaMemsetAsync( c->d_bitmap.p, 0, bmBytes, c->stream ) );
  RB_CUDA( cudaMemsetAsync( c->d_bnd_bitmap.p, 0, bmBytes, c->stream ) );
  if ( nPatches == 0 ) { return RB200_OK; }
  std::vector<PbfPatch> hl( nPatches );
  int64_t               total = 0;
  for ( int f = 0; f < F; f++ ) {
    if ( c->h_patch_off[f + 1] - c->h_patch_off[f] > 65535 ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "pbf: more than 65535 patches in a frame" ); }
    for ( int g = c->h_patch_off[f]; g < c->h_patch_off[f + 1]; g++ ) {
      const rb200_patch& q = c->h_patches[g];
      if ( q.lod_x != 1 || q.lod_y != 1 || q.axis_of_additional_plane != 0 ) {
        return rb_fail( c, RB200_ERR_UNSUPPORTED, "pbf: patch %d of frame %d has a level of detail / an additional plane (the "
                        "reference indexes its depth map with unscaled 3-D coordinates)", g - c->h_patch_off[f], f );
      }
      if ( (int64_t)q.size_u0 * 16 * q.size_v0 * 16 >= ( 1ll << 24 ) ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "pbf: patch too large" ); }
      hl[g].off = total;
      hl[g].w   = q.size_u0 * 16 + 2 * border;
      hl[g].h   = q.size_v0 * 16 + 2 * border;
      total += (int64_t)hl[g].w * hl[g].h;
    }
  }
  total = std::max<int64_t>( total, (int64_t)nPatches * 6 );  // k_pbf_local also initialises the boxes
  // arena: occA, occB, flag (1 byte each), depth (2), key (8), + tables
  RbBuf& A = c->d_pbf;
  auto   al = []( size_t x ) { return ( x + 255 ) & ~size_t( 255 ); };
  const size_t oOccA = 0, oOccB = oOccA + al( total ), oFlag = oOccB + al( total ), oDepth = oFlag + al( total ),
               oKey = oDepth + al( total * 2 ), oBox = oKey + al( total * 8 ), oLocal = oBox + al( (size_t)nPatches * 24 ),
               oPoff = oLocal + al( (size_t)nPatches * sizeof( PbfPatch ) ), end = oPoff + al( (size_t)( F + 1 ) * 4 );
  RB_CUDA( A.ensure( end ) );
  {
    const size_t tb = al( (size_t)nPatches * sizeof( PbfPatch ) ) + (size_t)( F + 1 ) * 4;
    char*        h  = (char*)rb_pinned_ring( c, tb );
    if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
    memcpy( h, hl.data(), (size_t)nPatches * sizeof( PbfPatch ) );
    memcpy( h + al( (size_t)nPatches * sizeof( PbfPatch ) ), c->h_patch_off.data(), (size_t)( F + 1 ) * 4 );
    RB_CUDA( cudaMemcpyAsync( A.as<char>() + oLocal, h, tb, cudaMemcpyHostToDevice, c->stream ) );
    c->stats.h2d_bytes += (int64_t)tb;
  }
  PbfArgs a{};
  a.patches   = c->d_patches.as<RbPatch>();
  a.local     = (const PbfPatch*)( A.as<char>() + oLocal );
  a.patch_off = (const int32_t*)( A.as<char>() + oPoff );
  a.nPatches  = nPatches;
  a.F         = F;
  a.total     = total;
  a.border    = border;
  a.W = c->W, a.H = c->H, a.oW = c->oW, a.oH = c->oH, a.M = c->M, a.prec = c->prec, a.Wb = c->Wb, a.Hb = c->Hb, a.bmWords = c->bmWords;
  a.threshold = P.enhanced_occupancy_map_code ? 0 : P.threshold_lossy_om;  // :551
  a.occ_video = c->d_occ_video.as<uint8_t>();
  a.geometry  = c->d_geometry.as<uint16_t>();
  a.b2p       = c->d_b2p.as<uint32_t>();
  a.occA      = A.as<uint8_t>() + oOccA;
  a.occB      = A.as<uint8_t>() + oOccB;
  a.flag      = A.as<uint8_t>() + oFlag;
  a.depth     = (int16_t*)( A.as<char>() + oDepth );
  a.key       = (unsigned long long*)( A.as<char>() + oKey );
  a.box       = (int32_t*)( A.as<char>() + oBox );
  a.dist2     = dist2;
  a.filterSize = P.pbf_filter_size;
  a.bitmap    = c->d_bitmap.as<uint32_t>();
  a.bnd       = c->d_bnd_bitmap.as<uint32_t>();
  const int G = rb_div_up( total, TPB );
  RB_LAUNCH( "pbf_local", k_pbf_local, G, TPB, 0, a );
  RB_LAUNCH( "pbf_border", k_pbf_border, G, TPB, 0, a );
  RB_LAUNCH( "pbf_neighbors", k_pbf_neighbors, G, TPB, 0, a );
  const uint8_t* src = a.occA;
  uint8_t*       dst = a.occB;
  for ( int it = 0; it < P.pbf_passes_count; it++ ) {  // :915-917: the maps swap roles every pass
    RB_LAUNCH( "pbf_filter", k_pbf_filter, G, TPB, 0, a, src, dst );
    const uint8_t* t = src;
    src              = dst;
    dst              = const_cast<uint8_t*>( t );
  }
  RB_LAUNCH( "pbf_scatter", k_pbf_scatter, G, TPB, 0, a, src );
  return RB200_OK;
}
