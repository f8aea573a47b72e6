// rb_reconstruct.cu — patch-to-3D reprojection of a whole GOF (sm_100a).
//
// Restates, as batched data-parallel kernels, the reference functions
//   PCCCodec::generateOccupancyMap                       PccLibCommon/source/PCCCodec.cpp:1584-1606
//   PCCCodec::generateBlockToPatchFromOccupancyMapVideo  :1725-1763
//   PCCCodec::generatePointCloud (+ generatePoints)      :517-978 (:327-515)
//   PCCCodec::identifyBoundaryPoints                     :266-325
//   PCCCodec::colorPointCloud (single-stream branch)     :1308-1449
//   PCCPatch::generatePoint / patch2Canvas / patchBlock2CanvasBlock   PCCPatch.h:177-207, PCCPatch.cpp:192-308
//
// Design (B200, HBM-bound integer work, no tensor cores):
//  * unit of work = one 16x16 patch block, in the reference's emission order (patch -> v0 -> u0); one warp per
//    block, lane l owns the 8 consecutive patch-local pixels (v1 = l/2, u1 = 8*(l%2)..+7), so lane order ==
//    emission order and an in-warp shuffle scan gives each point its ordered output slot;
//  * every HBM read of a frame plane is a 16-byte vector load of 8 consecutive canvas pixels (two lanes = one
//    32-byte sector); the canvas tile is staged in shared memory and re-addressed through the patch
//    orientation there, so rotated / mirrored patches cost no uncoalesced traffic;
//  * occupancy lives as a 1-bit-per-pixel bitmap (L2-resident, 1/16 of a byte map); the 3x3 / 5x5 boundary
//    stencils are bit tests on 20 staged row masks per tile;
//  * count -> exclusive scan -> emit keeps the point order of the reference exactly (ordered MD5 parity).
#include <cuda.h>  // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no link against libcuda)

#include <algorithm>

#include "rb_common.cuh"

namespace {

// ---- TMA (cp.async.bulk.tensor) + mbarrier: the canvas tiles of a patch block go global -> shared memory as two bulk
// tensor copies issued by one lane, instead of sixteen 16-byte loads and stores per lane ----
__device__ __forceinline__ uint32_t smem_u32( const void* p ) { return (uint32_t)__cvta_generic_to_shared( p ); }
__device__ __forceinline__ void mbar_init( uint64_t* bar, uint32_t count ) {
  asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( smem_u32( bar ) ), "r"( count ) : "memory" );
  // visible to the async proxy (the TMA unit) before the copy is issued; CTA scope: a cluster-scope fence would compile to
  // CCTL.IVALL and drop the SM's L1
  asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" );
}
__device__ __forceinline__ void mbar_expect_tx( uint64_t* bar, uint32_t bytes ) {
  asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( smem_u32( bar ) ), "r"( bytes ) : "memory" );
}
__device__ __forceinline__ void mbar_wait( uint64_t* bar, uint32_t parity ) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"( smem_u32( bar ) ),
      "r"( parity )
      : "memory" );
}
// box of the tensor map at (x, y, plane) -> dst; completion is counted in bytes on `bar`
__device__ __forceinline__ void tma_load_3d( void* dst, const CUtensorMap* map, int x, int y, int plane, uint64_t* bar ) {
  asm volatile( "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                    smem_u32( dst ) ),
                "l"( map ), "r"( x ), "r"( y ), "r"( plane ), "r"( smem_u32( bar ) )
                : "memory" );
}

struct ReprojArgs {
  const RbPatch*  patches;
  const int32_t*  wi_patch;
  const int32_t*  wi_local;
  int64_t         nWI;
  const uint8_t*  occ_video;
  const uint16_t* geo;
  const uint16_t* attr;
  const uint32_t* bitmap;
  const uint32_t* bnd_bitmap;  // occupancy synthesis: the boundary type of a pixel's points (PCCPatch::isBorder), else null
  const uint32_t* b2p;
  int             W, H, oW, oH, Wb, Hb, M, prec, bmWords;
  int             absolute_d1, remove_dup, eom_fix_bits, classify, attr_count, bitdepth3d;
  int             surface_thickness;
  const rb200_plr_mode* plr_modes;  // point local reconstruction (rb200_gof_set_plr)
  const uint8_t*  plr_block_mode;
  const int64_t*  plr_block_off;
  int             t1_bits;  // multi-stream attribute: 0 = map 1 is absolute, 8 / 16 = map 1 is a delta on map 0
  int32_t*        wi_count;
  uint32_t*       wi_cnt4;  // [nWI][32] per-lane packed pixel counts of the counting pass, reused by the emitting pass
  int32_t*        wi_eom_count;
  const int64_t*  wi_base;
  const int64_t*  wi_eom_base;
  const int32_t*  frame_wi_off;
  const int64_t*  frame_off;
  short4*         pos;
  ushort4*        col;
  uint32_t*       pix;
  uint32_t*       part;
  RbFrameInfo*    finfo;
  short4*         eom_stage;
  uint32_t*       blist;    // indices of the points classified as boundary (type 1)
  uint32_t*       blist_n;
};

// PCCPatch::patch2Canvas restricted to one 16x16 block (PCCPatch.cpp:192-251)
__device__ __forceinline__ void tile_coord( int orient, int u1, int v1, int& tx, int& ty ) {
  switch ( orient ) {  // enum PCCPatchOrientation, PccLibBitstreamCommon/include/PCCBitstreamCommon.h:120-130
    default:
    case 0: tx = u1; ty = v1; break;            // DEFAULT
    case 1: tx = v1; ty = u1; break;            // SWAP
    case 2: tx = 15 - v1; ty = u1; break;       // ROT90
    case 3: tx = 15 - u1; ty = 15 - v1; break;  // ROT180
    case 4: tx = v1; ty = 15 - u1; break;       // ROT270
    case 5: tx = 15 - u1; ty = v1; break;       // MIRROR
    case 6: tx = 15 - v1; ty = 15 - u1; break;  // MROT90
    case 7: tx = u1; ty = 15 - v1; break;       // MROT180
    case 8: tx = v1; ty = u1; break;            // MROT270
  }
}

// the same map as two affine forms, set up once per patch block: tx = au u1 + av v1 + a0, ty = bu u1 + bv v1 + b0
struct TileMap {
  int au, av, a0, bu, bv, b0;
};
__device__ __forceinline__ TileMap tile_map( int orient ) {
  switch ( orient ) {
    default:
    case 0: return {1, 0, 0, 0, 1, 0};
    case 1: return {0, 1, 0, 1, 0, 0};
    case 2: return {0, -1, 15, 1, 0, 0};
    case 3: return {-1, 0, 15, 0, -1, 15};
    case 4: return {0, 1, 0, -1, 0, 15};
    case 5: return {-1, 0, 15, 0, 1, 0};
    case 6: return {0, -1, 15, -1, 0, 15};
    case 7: return {1, 0, 0, 0, -1, 15};
    case 8: return {0, 1, 0, 1, 0, 0};
  }
}
#define TILE_XY( tm, u1, v1, tx, ty )                        \
  do {                                                        \
    tx = ( tm ).au * ( u1 ) + ( tm ).av * ( v1 ) + ( tm ).a0; \
    ty = ( tm ).bu * ( u1 ) + ( tm ).bv * ( v1 ) + ( tm ).b0; \
  } while ( 0 )

// PCCPatch::patchBlock2CanvasBlock (PCCPatch.cpp:253-308); patches are validated inside the canvas at upload
__device__ __forceinline__ void canvas_block( const RbPatch& p, int ub, int vb, int& bx, int& by ) {
  switch ( p.orient ) {
    default:
    case 0: bx = ub + p.u0; by = vb + p.v0; break;
    case 1: bx = vb + p.u0; by = ub + p.v0; break;
    case 2: bx = ( p.sv0 - 1 - vb ) + p.u0; by = ub + p.v0; break;
    case 3: bx = ( p.su0 - 1 - ub ) + p.u0; by = ( p.sv0 - 1 - vb ) + p.v0; break;
    case 4: bx = vb + p.u0; by = ( p.su0 - 1 - ub ) + p.v0; break;
    case 5: bx = ( p.su0 - 1 - ub ) + p.u0; by = vb + p.v0; break;
    case 6: bx = ( p.sv0 - 1 - vb ) + p.u0; by = ( p.su0 - 1 - ub ) + p.v0; break;
    case 7: bx = ub + p.u0; by = ( p.sv0 - 1 - vb ) + p.v0; break;
    case 8: bx = vb + p.u0; by = ub + p.v0; break;
  }
}

// PCCPatch::generateNormalCoordinate (PCCPatch.h:177-186); the double temporaries are always integral
__device__ __forceinline__ int normal_coord( const RbPatch& p, int depth ) {
  return p.mode == 0 ? depth + p.d1 : max( p.d1 - depth, 0 );
}

__device__ __forceinline__ void set_axis( int16_t ( &P )[3], int axis, int v ) {
  if ( axis == 0 ) {
    P[0] = (int16_t)v;
  } else if ( axis == 1 ) {
    P[1] = (int16_t)v;
  } else {
    P[2] = (int16_t)v;
  }
}
__device__ __forceinline__ int get_axis( const int16_t ( &P )[3], int axis ) {
  return axis == 0 ? P[0] : ( axis == 1 ? P[1] : P[2] );
}

// PCCCodec::inverseRotatePosition45DegreeOnAxis (PCCCodec.cpp:2503-2524) followed by addPoint( PCCVector3D )
// (PCCPointSet.h:419-426: truncation to int16).  The reference adds a size_t, so a negative intermediate wraps
// to a huge double whose int16 cast yields 0 on x86-64; mirrored here.
__device__ __forceinline__ int16_t half_trunc( int v ) { return v < 0 ? (int16_t)0 : (int16_t)( v / 2 ); }
__device__ __forceinline__ void   inverse_rotate45( int axis, int lod, int16_t ( &P )[3] ) {
  const int s = ( 1 << ( lod - 1 ) ) - 1;
  const int x = P[0], y = P[1], z = P[2];
  if ( axis == 1 ) {
    P[0] = half_trunc( x - z + s );
    P[2] = half_trunc( x + z - s );
  } else if ( axis == 2 ) {
    P[2] = half_trunc( z - y + s );
    P[1] = half_trunc( z + y - s );
  } else if ( axis == 3 ) {
    P[1] = half_trunc( y - x + s );
    P[0] = half_trunc( y + x - s );
  }
}

// ---------------------------------------------------------------------------------------------------
// K0: occupancy video -> full-resolution bitmap.  generateOccupancyMap (:1584-1606) and the upsampling
// loop of generatePointCloud (:557-570): O[v][u] = video[v/p][u/p] > threshold (EOM: raw symbol != 0).
// One thread per 32-pixel word.
// ---------------------------------------------------------------------------------------------------
__global__ void k_occupancy_bitmap( const uint8_t* __restrict__ video, uint32_t* __restrict__ bitmap, int F, int W,
                                    int H, int oW, int oH, int prec, int words, int threshold, int eom ) {
  const int64_t total = (int64_t)F * H * words;
  int64_t       i     = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= total ) { return; }
  const int      w    = (int)( i % words );
  const int      y    = (int)( ( i / words ) % H );
  const int      f    = (int)( i / ( (int64_t)words * H ) );
  const uint8_t* row  = video + ( (size_t)f * oH + y / prec ) * oW;
  uint32_t       bits = 0;
  const int      x0   = w * 32;
  const uint32_t run  = prec >= 32 ? 0xFFFFFFFFu : ( ( 1u << prec ) - 1u );
  for ( int s = 0; s * prec < 32; s++ ) {
    const int x = x0 + s * prec;
    if ( x >= W ) { break; }
    const int v  = row[x / prec];
    // generateOccupancyMap thresholds the video sample IN PLACE once per full-resolution pixel (:1597-1600), i.e.
    // p*p times per sample: with threshold >= 1 and p > 1 the second visit sees the already binarised 0/1 and
    // clears it, so the re-upsampled map of generatePointCloud (:557-570) is empty.  Reproduced, not "fixed".
    const int on = eom ? ( v != 0 ) : ( ( prec == 1 || threshold == 0 ) ? ( v > threshold ) : 0 );
    if ( on ) { bits |= run << ( s * prec ); }
  }
  if ( x0 + 32 > W ) { bits &= ( 1u << ( W - x0 ) ) - 1u; }
  bitmap[i] = bits;
}

// ---------------------------------------------------------------------------------------------------
// K1: block-to-patch (:1725-1763): the highest patch index with an occupied pixel in the block wins.
// One thread per patch block; 16 row-word tests.
// ---------------------------------------------------------------------------------------------------
__global__ void k_block_to_patch( const RbPatch* __restrict__ patches, const int32_t* __restrict__ wi_patch,
                                  const int32_t* __restrict__ wi_local, int64_t nWI,
                                  const uint32_t* __restrict__ bitmap, uint32_t* __restrict__ b2p, int H, int Wb,
                                  int Hb, int words ) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= nWI ) { return; }
  const RbPatch p  = patches[wi_patch[i]];
  const int     lb = wi_local[i];
  int           bx, by;
  canvas_block( p, lb % p.su0, lb / p.su0, bx, by );
  const uint32_t* bm  = bitmap + ( (size_t)p.frame * H + by * 16 ) * words + ( bx >> 1 );
  const uint32_t  m   = ( bx & 1 ) ? 0xFFFF0000u : 0x0000FFFFu;
  uint32_t        any = 0;
#pragma unroll
  for ( int r = 0; r < 16; r++ ) { any |= bm[(size_t)r * words] & m; }
  if ( any ) { atomicMax( &b2p[( (size_t)p.frame * Hb + by ) * Wb + bx], (uint32_t)p.frame_patch + 1u ); }
}

// size quantisation (:571-597): pixels of an owned block beyond the quantised patch size are cleared
__global__ void k_size_quantization( const RbPatch* __restrict__ patches, const int32_t* __restrict__ wi_patch,
                                     const int32_t* __restrict__ wi_local, int64_t nWI, uint32_t* __restrict__ bitmap,
                                     const uint32_t* __restrict__ b2p, int H, int Wb, int Hb, int words, int log2qx,
                                     int log2qy ) {
  const int64_t i    = blockIdx.x;
  const RbPatch p    = patches[wi_patch[i]];
  const int     lb   = wi_local[i];
  const int     ub   = lb % p.su0, vb = lb / p.su0;
  int           bx, by;
  canvas_block( p, ub, vb, bx, by );
  if ( b2p[( (size_t)p.frame * Hb + by ) * Wb + bx] != (uint32_t)p.frame_patch + 1u ) { return; }
  const int sx = ( p.s2dx >> log2qx ) << log2qx, sy = ( p.s2dy >> log2qy ) << log2qy;
  const int u1 = threadIdx.x & 15, v1 = threadIdx.x >> 4;
  const int u = ub * 16 + u1, v = vb * 16 + v1;
  if ( u >= sx || v >= sy ) {
    int tx, ty;
    tile_coord( p.orient, u1, v1, tx, ty );
    const int x = bx * 16 + tx, y = by * 16 + ty;
    atomicAnd( &bitmap[( (size_t)p.frame * H + y ) * words + ( x >> 5 )], ~( 1u << ( x & 31 ) ) );
  }
}

// ---------------------------------------------------------------------------------------------------
// K2 / K4: the per-pixel reprojection, as a counting pass (EMIT=false) and an emitting pass (EMIT=true).
// 8 warps per CTA, one 16x16 patch block per warp.
// ---------------------------------------------------------------------------------------------------
constexpr int WARPS = 4;

struct __align__( 128 ) TileSmem {  // g and a are the destinations of the bulk tensor copies (128-byte aligned)
  uint16_t g[2][256];     // geometry D0 / D1 tile
  uint16_t a[2][3][256];  // attribute tiles
  uint32_t rows[20];      // occupancy bits of canvas rows Y0-2..Y0+17, bit k <-> x = X0-2+k
  uint16_t bnd[16];       // per tile row: bit tx = the pixel is a boundary pixel (identifyBoundaryPoints)
  uint16_t desc[512];     // emission-ordered point descriptors: u1 | v1 << 4 | layer << 8 | boundary << 15
  uint64_t bar;           // mbarrier the tile copies complete on
};

// STD: the CTC configuration (two maps, absolute D1, attributes, boundary classification, no delta-coded T1) with
// its switches resolved at compile time
template <bool EMIT, bool EOM, bool STD>
__global__ void __launch_bounds__( WARPS * 32, 8 ) k_reproject( const ReprojArgs a, const __grid_constant__ CUtensorMap tmGeo,
                                                                const __grid_constant__ CUtensorMap tmAttr ) {
  __shared__ TileSmem sm[WARPS];
  const int     lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int     kM   = STD ? 2 : a.M;
  const bool    kAbs = STD ? true : a.absolute_d1 != 0, kAttr = STD ? true : a.attr_count > 0, kCls = STD ? true : a.classify != 0;
  const int64_t wi   = (int64_t)blockIdx.x * WARPS + wq;
  if ( wi >= a.nWI ) { return; }
  TileSmem&     S  = sm[wq];
  const RbPatch p  = a.patches[a.wi_patch[wi]];
  const int     lb = a.wi_local[wi];
  const int     ub = lb % p.su0, vb = lb / p.su0;
  int           bx, by;
  canvas_block( p, ub, vb, bx, by );
  const int  f     = p.frame;
  const bool owned = a.b2p[( (size_t)f * a.Hb + by ) * a.Wb + bx] == (uint32_t)p.frame_patch + 1u;  // :650
  if ( !owned ) {
    if ( !EMIT && lane == 0 ) {
      a.wi_count[wi] = 0;
      if ( EOM ) { a.wi_eom_count[wi] = 0; }
    }
    return;
  }
  const int     X0 = bx * 16, Y0 = by * 16;
  const TileMap tm = tile_map( p.orient );
  // ---- the geometry (and attribute) tiles of the block: one bulk tensor copy each (box 16 x 16 x planes of the frame),
  // issued by lane 0 and landing on the warp's mbarrier while the other lanes stage the occupancy rows ----
  if ( lane == 0 ) { mbar_init( &S.bar, 1 ); }
  __syncwarp();
  if ( lane == 0 ) {
    const bool attrTiles = EMIT && kAttr;
    mbar_expect_tx( &S.bar, (uint32_t)( kM * 512 + ( attrTiles ? kM * 3 * 512 : 0 ) ) );
    tma_load_3d( &S.g[0][0], &tmGeo, X0, Y0, f * kM, &S.bar );
    if ( attrTiles ) { tma_load_3d( &S.a[0][0][0], &tmAttr, X0, Y0, f * kM * 3, &S.bar ); }
  }
  // ---- stage occupancy row masks ----
  if ( lane < 20 ) {
    const int y    = Y0 - 2 + lane;
    uint32_t  bits = 0xFFFFFu;  // outside the image: treated as occupied, borders are handled explicitly
    if ( y >= 0 && y < a.H ) {
      const uint32_t* row = a.bitmap + ( (size_t)f * a.H + y ) * a.bmWords;
      const int       xs  = X0 - 2;  // first x of the window (may be -2)
      // assemble bits xs..xs+19 from up to two words
      uint64_t  win = 0;
      const int w0  = ( xs < 0 ) ? -1 : ( xs >> 5 );
      uint32_t  lo  = ( w0 >= 0 ) ? row[w0] : 0u;
      uint32_t  hi  = ( w0 + 1 < a.bmWords ) ? row[w0 + 1] : 0u;
      win           = ( (uint64_t)hi << 32 ) | lo;
      const int sh  = ( xs < 0 ) ? ( xs + 32 ) : ( xs & 31 );
      bits          = (uint32_t)( win >> sh ) & 0xFFFFFu;
      // pixels outside [0,W) read as occupied
      if ( xs < 0 ) { bits |= 0x3u; }
      if ( xs + 20 > a.W ) { bits |= ( 0xFFFFFu << ( a.W - xs ) ) & 0xFFFFFu; }
    }
    S.rows[lane] = bits;
  }
  mbar_wait( &S.bar, 0 );  // the tiles have landed (the wait also orders the async writes before the reads below)
  __syncwarp();
  if ( EMIT && !STD && a.t1_bits && kAttr && kM > 1 ) {
    // multiple streams with a delta-coded second map (colorPointCloud, PCCCodec.cpp:1387-1416): the tile of map 1
    // is reconstructed in shared memory once, T1 = clip( T0 + clamp( T1 - offset, -offset, offset - 1 ), 0, max )
    const int offset = 1 << ( a.t1_bits - 1 ), maxv = ( 1 << a.t1_bits ) - 1;
#pragma unroll
    for ( int ch = 0; ch < 3; ch++ ) {
#pragma unroll
      for ( int j = 0; j < 8; j++ ) {
        const int i  = lane * 8 + j;
        int       nv = (int)S.a[1][ch][i] - offset;
        nv           = min( max( nv, -offset ), offset - 1 ) + (int)S.a[0][ch][i];
        S.a[1][ch][i] = (uint16_t)min( max( nv, 0 ), maxv );
      }
    }
    __syncwarp();
  }

  if ( EMIT && kCls ) {
    // identifyBoundaryPoints (:266-325) for the whole tile at once, 16 rows in 16 lanes: a pixel is type 1 when it is
    // on / next to the image border or any pixel of its 5x5 neighbourhood is unoccupied (the 3x3 test of :274-305 is
    // implied: a full 5x5 contains a full 3x3).  Pixels outside the image read as occupied (staging above).
    if ( !STD && a.bnd_bitmap ) {
      // occupancy synthesis: reconstruct.setBoundaryPointType( isBoundary ) with patch.isBorder( u, v ) (:664, :809)
      if ( lane < 16 ) {
        const uint32_t wv = a.bnd_bitmap[( (size_t)f * a.H + Y0 + lane ) * a.bmWords + ( X0 >> 5 )];
        S.bnd[lane]       = (uint16_t)( wv >> ( X0 & 31 ) );
      }
    } else if ( lane < 16 ) {
      const int      r  = lane + 2, y = Y0 + lane;
      const uint32_t v5 = S.rows[r - 2] & S.rows[r - 1] & S.rows[r] & S.rows[r + 1] & S.rows[r + 2];
      const uint32_t h5 = v5 & ( v5 >> 1 ) & ( v5 << 1 ) & ( v5 >> 2 ) & ( v5 << 2 );  // bit kx: columns kx-2..kx+2 full
      uint32_t       b  = ( ~h5 >> 2 ) & 0xFFFFu;                                        // bit tx <-> kx = tx + 2
      if ( y <= 1 || y >= a.H - 2 ) { b = 0xFFFFu; }
      if ( X0 == 0 ) { b |= 0x3u; }
      if ( X0 + 16 >= a.W ) { b |= ( X0 + 16 == a.W ) ? 0xC000u : 0xFFFFu; }
      S.bnd[lane] = (uint16_t)b;
    }
    __syncwarp();
  }
  const int v1 = lane >> 1, ubase = 8 * ( lane & 1 );
  // ---- pass A: per-pixel point counts (4 bits per pixel: regular; EOM extras separately) ----
  const int nsgnA = p.mode == 0 ? 1 : -1, nloA = p.mode == 0 ? -( 1 << 30 ) : 0;  // generateNormalCoordinate, see pass B
  uint32_t  cnt4 = 0, eom4 = 0;
  const bool reuse = EMIT && !EOM && a.wi_cnt4 != nullptr;  // the counting pass left this lane's counts behind
  if ( reuse ) { cnt4 = a.wi_cnt4[wi * 32 + lane]; }
#pragma unroll
  for ( int j = 0; j < 8; j++ ) {
    if ( reuse ) { break; }
    int tx, ty;
    TILE_XY( tm, ubase + j, v1, tx, ty );
    const uint32_t occ = ( S.rows[ty + 2] >> ( tx + 2 ) ) & 1u;
    if ( !occ ) { continue; }
    const int d0 = S.g[0][ty * 16 + tx];
    const int g1 = ( kM > 1 ) ? S.g[1][ty * 16 + tx] : 0;
    int       c  = 1, e = 0;
    if ( EOM ) {
      // :686-714 eomCode from the D1-D0 difference and the occupancy symbol
      const int sym = a.occ_video[( (size_t)f * a.oH + ( Y0 + ty ) / a.prec ) * a.oW + ( X0 + tx ) / a.prec];
      uint32_t  code;
      if ( kM > 1 ) {
        const int diff = kAbs ? (int)(int16_t)g1 - (int)(int16_t)d0 : (int)(int16_t)g1;
        if ( diff == 1 ) {
          code = 1;
        } else if ( diff > 1 ) {
          const uint32_t top = 1u << ( ( diff - 1 ) & 31 );
          code               = ( ( ( top - (uint32_t)sym ) & 0xFFFFu ) | top ) & 0xFFFFu;
        } else {
          code = 0;
        }
      } else {
        code = ( ( 1u << ( a.eom_fix_bits & 31 ) ) - (uint32_t)sym ) & 0xFFFFu;
      }
      if ( code == 0 ) {
        if ( !a.remove_dup ) { c = 2; }
      } else {
        const int nb = __popc( code & 0x3FFu );
        if ( kM > 1 && nb > 0 ) {
          c = 2;
          e = nb - 1;
        } else {
          e = nb;
        }
      }
    } else if ( kM > 1 ) {
      // :497-512 far layer; :794-795 duplicate removal compares the int16 points
      const int16_t n0 = (int16_t)max( p.d1 + nsgnA * d0, nloA );
      const int16_t n1 = (int16_t)( kAbs ? max( p.d1 + nsgnA * g1, nloA ) : (int)n0 + nsgnA * g1 );
      c                = ( a.remove_dup && n1 == n0 ) ? 1 : 2;
    }
    cnt4 |= (uint32_t)c << ( 4 * j );
    eom4 |= (uint32_t)e << ( 4 * j );
  }
  int myCount = 0, myEom = 0;
#pragma unroll
  for ( int j = 0; j < 8; j++ ) {
    myCount += ( cnt4 >> ( 4 * j ) ) & 15;
    myEom += ( eom4 >> ( 4 * j ) ) & 15;
  }
  // ---- ordered in-warp exclusive scan (lane order == emission order v1, u1) ----
  int incl = myCount, inclE = myEom;
#pragma unroll
  for ( int d = 1; d < 32; d <<= 1 ) {
    const int t  = __shfl_up_sync( 0xFFFFFFFFu, incl, d );
    const int tE = __shfl_up_sync( 0xFFFFFFFFu, inclE, d );
    if ( lane >= d ) {
      incl += t;
      inclE += tE;
    }
  }
  if ( !EMIT ) {
    if ( lane == 31 ) {
      a.wi_count[wi] = incl;
      if ( EOM ) { a.wi_eom_count[wi] = inclE; }
    }
    if ( !EOM && a.wi_cnt4 ) { a.wi_cnt4[wi * 32 + lane] = cnt4; }
    return;
  }

  // ---- pass B: emit ----
  const int64_t base  = a.frame_off[f] + ( a.wi_base[wi] - a.wi_base[a.frame_wi_off[f]] ) + ( incl - myCount );
  int     maxc = 0;
  if ( !EOM ) {
    // every lane lists its points (emission order = lane order) as descriptors in shared memory; then the warp walks
    // the list 32 points at a time, so consecutive lanes compute and store consecutive points: full-line stores
    {
      int o = incl - myCount;
#pragma unroll
      for ( int j = 0; j < 8; j++ ) {
        const int c = ( cnt4 >> ( 4 * j ) ) & 15;
        if ( c > 0 ) { S.desc[o++] = (uint16_t)( ( ubase + j ) | ( v1 << 4 ) ); }
        if ( c > 1 ) { S.desc[o++] = (uint16_t)( ( ubase + j ) | ( v1 << 4 ) | 0x100 ); }
      }
    }
    const int     total = __shfl_sync( 0xFFFFFFFFu, incl, 31 );
    const int64_t wbase = a.frame_off[f] + ( a.wi_base[wi] - a.wi_base[a.frame_wi_off[f]] );
    __syncwarp();
    int nb = 0;
    // generateNormalCoordinate (PCCPatch.h:177-186) as max( d1 + nsgn * depth, nlo )
    const int nsgn = p.mode == 0 ? 1 : -1, nlo = p.mode == 0 ? -( 1 << 30 ) : 0;
    // bit c of each nibble: output coordinate c takes the tangent (bits 0-2) / bitangent (4-6) / normal (8-10) value
    uint32_t axw = 0;
    {
      const uint32_t wn = 1u << p.normal_axis, wb = ( 1u << p.bitangent_axis ) & ~wn, wt = ( 1u << p.tangent_axis ) & ~wn & ~wb;
      axw               = wt | ( wb << 4 ) | ( wn << 8 );
    }
    for ( int k0 = 0; k0 < total; k0 += 32 ) {
      const int k     = k0 + lane;
      int       btype = 0;
      if ( k < total ) {
        const int d  = S.desc[k];
        const int u1 = d & 15, vv1 = ( d >> 4 ) & 15, layer = ( d >> 8 ) & 1;
        int       tx, ty;
        TILE_XY( tm, u1, vv1, tx, ty );
        const int x = X0 + tx, y = Y0 + ty;
        const int u = ub * 16 + u1, v = vb * 16 + vv1;
        const int d0 = S.g[0][ty * 16 + tx];
        // PCCPatch::generatePoint (PCCPatch.h:201-207), branch-free: the axis permutation is three 0/1 weights per
        // output coordinate set up once per block (later writes win: tangent, bitangent, normal); the normal
        // coordinate of both layers is formed and selected
        const int g1 = kM > 1 ? (int)S.g[1][ty * 16 + tx] : 0;
        const int n0 = (int16_t)max( p.d1 + nsgn * d0, nlo );   // generateNormalCoordinate( D0 )
        int       n1 = max( p.d1 + nsgn * g1, nlo );            // generatePoint( u, v, frame1 ), :503
        if ( !kAbs ) { n1 = n0 + nsgn * g1; }          // :505-509
        const int nn = layer ? n1 : n0;
        const int tt = u * p.lodx + p.u1, bb = v * p.lody + p.v1;
        int16_t   Q[3];
#pragma unroll
        for ( int cdx = 0; cdx < 3; cdx++ ) {
          Q[cdx] = (int16_t)( ( ( axw >> cdx ) & 1 ) * tt + ( ( axw >> ( 4 + cdx ) ) & 1 ) * bb + ( ( axw >> ( 8 + cdx ) ) & 1 ) * nn );
        }
        if ( p.addplane ) { inverse_rotate45( p.addplane, a.bitdepth3d, Q ); }
        if ( kCls ) { btype = ( S.bnd[ty] >> tx ) & 1; }  // identifyBoundaryPoints, per-tile masks above
        const int64_t o = wbase + k;
        a.pos[o]        = make_short4( Q[0], Q[1], Q[2], (short)btype );
        ushort4 cv      = make_ushort4( 0, 0, 0, (unsigned short)layer );
        if ( kAttr ) {
          cv = make_ushort4( S.a[layer][0][ty * 16 + tx], S.a[layer][1][ty * 16 + tx], S.a[layer][2][ty * 16 + tx],
                             (unsigned short)layer );
        }
        a.col[o]  = cv;
        a.pix[o]  = (uint32_t)x | ( (uint32_t)y << 16 );
        a.part[o] = (uint32_t)p.frame_patch;
        maxc      = max( maxc, max( (int)Q[0], max( (int)Q[1], (int)Q[2] ) ) );
        if ( btype ) { S.desc[k] = (uint16_t)( d | 0x8000 ); }
      }
      nb += __popc( __ballot_sync( 0xFFFFFFFFu, btype != 0 ) );
    }
    if ( nb > 0 ) {  // boundary list: one atomic per patch block, entries written 32 at a time
      uint32_t lb = 0;
      if ( lane == 0 ) { lb = atomicAdd( a.blist_n, (uint32_t)nb ); }
      lb = __shfl_sync( 0xFFFFFFFFu, lb, 0 );
      __syncwarp();
      for ( int k0 = 0; k0 < total; k0 += 32 ) {
        const int      k    = k0 + lane;
        const bool     isb  = k < total && ( S.desc[k] & 0x8000 );
        const uint32_t bal  = __ballot_sync( 0xFFFFFFFFu, isb );
        if ( isb ) { a.blist[lb + __popc( bal & ( ( 1u << lane ) - 1u ) )] = (uint32_t)( wbase + k ); }
        lb += __popc( bal );
      }
    }
  } else {
  int64_t ebase = a.wi_eom_base[wi] + ( inclE - myEom );
  int     k = 0, ke = 0;
#pragma unroll
  for ( int j = 0; j < 8; j++ ) {
    const int c = ( cnt4 >> ( 4 * j ) ) & 15;
    if ( c == 0 ) { continue; }
    int       tx, ty;
    const int u1 = ubase + j;
    TILE_XY( tm, u1, v1, tx, ty );
    const int x = X0 + tx, y = Y0 + ty;
    const int u = ub * 16 + u1, v = vb * 16 + v1;
    const int d0 = S.g[0][ty * 16 + tx];
    const int g1 = ( kM > 1 ) ? S.g[1][ty * 16 + tx] : 0;
    // PCCPatch::generatePoint (PCCPatch.h:201-207)
    int16_t P0[3] = {0, 0, 0};
    set_axis( P0, p.normal_axis, normal_coord( p, d0 ) );
    set_axis( P0, p.tangent_axis, u * p.lodx + p.u1 );
    set_axis( P0, p.bitangent_axis, v * p.lody + p.v1 );
    // boundary classification (identifyBoundaryPoints, :266-325) on the staged row masks
    int btype = 0;
    if ( kCls ) {
      const int      r  = ty + 2, kx = tx + 2;
      const uint32_t m3 = 7u << ( kx - 1 ), m5 = 31u << ( kx - 2 );
      if ( x == 0 || y == 0 || x == a.W - 1 || y == a.H - 1 ) {
        btype = 1;
      } else if ( ( S.rows[r - 1] & m3 ) != m3 || ( S.rows[r] & m3 ) != m3 || ( S.rows[r + 1] & m3 ) != m3 ) {
        btype = 1;
      } else if ( ( S.rows[r - 2] & m5 ) != m5 || ( S.rows[r - 1] & m5 ) != m5 || ( S.rows[r] & m5 ) != m5 ||
                  ( S.rows[r + 1] & m5 ) != m5 || ( S.rows[r + 2] & m5 ) != m5 ) {
        btype = 1;
      } else if ( x == 1 || y == 1 || x == a.W - 2 || y == a.H - 2 ) {
        btype = 1;
      }
    }
    const uint32_t pixv = (uint32_t)x | ( (uint32_t)y << 16 );
    // layer 0
    {
      int16_t Q[3] = {P0[0], P0[1], P0[2]};
      if ( p.addplane ) { inverse_rotate45( p.addplane, a.bitdepth3d, Q ); }
      const int64_t o = base + k;
      a.pos[o]        = make_short4( Q[0], Q[1], Q[2], (short)btype );
      ushort4 cv      = make_ushort4( 0, 0, 0, 0 );
      if ( kAttr ) { cv = make_ushort4( S.a[0][0][ty * 16 + tx], S.a[0][1][ty * 16 + tx], S.a[0][2][ty * 16 + tx], 0 ); }
      a.col[o]  = cv;
      a.pix[o]  = pixv;
      a.part[o] = (uint32_t)p.frame_patch;
      maxc      = max( maxc, max( (int)Q[0], max( (int)Q[1], (int)Q[2] ) ) );
      k++;
    }
    if ( !EOM ) {
      if ( c > 1 ) {
        int16_t P1[3] = {P0[0], P0[1], P0[2]};
        if ( kAbs ) {
          set_axis( P1, p.normal_axis, normal_coord( p, g1 ) );  // generatePoint( u, v, frame1 ), :503
        } else {
          const int n0 = get_axis( P0, p.normal_axis );
          set_axis( P1, p.normal_axis, p.mode == 0 ? n0 + g1 : n0 - g1 );  // :505-509
        }
        if ( p.addplane ) { inverse_rotate45( p.addplane, a.bitdepth3d, P1 ); }
        const int64_t o = base + k;
        a.pos[o]        = make_short4( P1[0], P1[1], P1[2], (short)btype );
        ushort4 cv      = make_ushort4( 0, 0, 0, 1 );
        if ( kAttr ) {
          cv = make_ushort4( S.a[1][0][ty * 16 + tx], S.a[1][1][ty * 16 + tx], S.a[1][2][ty * 16 + tx], 1 );
        }
        a.col[o]  = cv;
        a.pix[o]  = pixv;
        a.part[o] = (uint32_t)p.frame_patch;
        maxc      = max( maxc, max( (int)P1[0], max( (int)P1[1], (int)P1[2] ) ) );
        k++;
      }
    } else {
      // EOM branch (:669-779): recompute the code, emit the in-place D1 and stage the extra points
      const int sym = a.occ_video[( (size_t)f * a.oH + y / a.prec ) * a.oW + x / a.prec];
      uint32_t  code;
      if ( kM > 1 ) {
        const int diff = kAbs ? (int)(int16_t)g1 - (int)(int16_t)d0 : (int)(int16_t)g1;
        if ( diff == 1 ) {
          code = 1;
        } else if ( diff > 1 ) {
          const uint32_t top = 1u << ( ( diff - 1 ) & 31 );
          code               = ( ( ( top - (uint32_t)sym ) & 0xFFFFu ) | top ) & 0xFFFFu;
        } else {
          code = 0;
        }
      } else {
        code = ( ( 1u << ( a.eom_fix_bits & 31 ) ) - (uint32_t)sym ) & 0xFFFFu;
      }
      const int n0 = get_axis( P0, p.normal_axis );
      if ( code == 0 ) {
        if ( c > 1 ) {  // !removeDuplicatePoints: D1 == D0 (:717-732)
          int16_t Q[3] = {P0[0], P0[1], P0[2]};
          if ( p.addplane ) { inverse_rotate45( p.addplane, a.bitdepth3d, Q ); }
          const int64_t o = base + k;
          a.pos[o]        = make_short4( Q[0], Q[1], Q[2], (short)btype );
          ushort4 cv      = make_ushort4( 0, 0, 0, 1 );
          if ( kAttr && kM > 1 ) {
            cv = make_ushort4( S.a[1][0][ty * 16 + tx], S.a[1][1][ty * 16 + tx], S.a[1][2][ty * 16 + tx], 1 );
          } else if ( kAttr ) {
            cv = make_ushort4( 0, 0, 0, 1 );
          }
          a.col[o]  = cv;
          a.pix[o]  = pixv;
          a.part[o] = (uint32_t)p.frame_patch;
          k++;
        }
      } else {
        const uint32_t low   = code & 0x3FFu;
        const int      d1pos = 31 - __clz( low | 0u );  // highest set bit among bits 0..9 (:736-738); -1 if none
        for ( int i = 0; i < 10; i++ ) {
          if ( !( low & ( 1u << i ) ) ) { continue; }
          int16_t Q[3] = {P0[0], P0[1], P0[2]};
          set_axis( Q, p.normal_axis, p.mode == 0 ? n0 + ( i + 1 ) : n0 - ( i + 1 ) );  // :741-748
          if ( p.addplane ) { inverse_rotate45( p.addplane, a.bitdepth3d, Q ); }
          if ( ( code == 1 || i == d1pos ) && kM > 1 ) {  // in-place D1 (:749-762)
            const int64_t o = base + k;
            a.pos[o]        = make_short4( Q[0], Q[1], Q[2], (short)btype );
            ushort4 cv      = make_ushort4( 0, 0, 0, 1 );
            if ( kAttr ) {
              cv = make_ushort4( S.a[1][0][ty * 16 + tx], S.a[1][1][ty * 16 + tx], S.a[1][2][ty * 16 + tx], 1 );
            }
            a.col[o]  = cv;
            a.pix[o]  = pixv;
            a.part[o] = (uint32_t)p.frame_patch;
            maxc      = max( maxc, max( (int)Q[0], max( (int)Q[1], (int)Q[2] ) ) );
            k++;
          } else {  // eomPointsPerPatch[patchIndex].push_back (:763-771)
            a.eom_stage[ebase + ke] = make_short4( Q[0], Q[1], Q[2], 0 );
            ke++;
          }
        }
      }
    }
  }
  }
  // per-frame max coordinate, feeds the geometry-smoothing grid width (:68-79)
#pragma unroll
  for ( int d = 16; d > 0; d >>= 1 ) { maxc = max( maxc, __shfl_xor_sync( 0xFFFFFFFFu, maxc, d ) ); }
  // (a stale cached value only costs a redundant atomic; without the test every warp hits the same address)
  if ( lane == 0 && maxc > 0 && maxc > a.finfo[f].max_coord ) { atomicMax( &a.finfo[f].max_coord, maxc ); }
}

// ---------------------------------------------------------------------------------------------------
// singleMapPixelInterleaving (generatePoints, PCCCodec.cpp:350-471): one map whose pixels alternate between the near
// and the far layer on a checkerboard.  Every occupied pixel yields its coded point, the other layer interpolated
// from the 4-neighbours of the same patch, and the fill points strictly between the two (:463-468).  Same work
// items, counting / emitting passes and point order as k_reproject; the pixels are independent, so every lane walks
// its 8 pixels on its own (this is not the CTC path: plain global loads, no staging).
// ---------------------------------------------------------------------------------------------------
struct IlvPixel {
  int n0, n1;  // normal coordinate of the coded point / of the interpolated one
  int two;     // the interpolated point exists (count != 0, :434)
  int xmin, nfill;
};

__device__ __forceinline__ IlvPixel interleave_pixel( const ReprojArgs& a, const RbPatch& p, int f, int x, int y ) {
  const size_t    plane = (size_t)a.W * a.H;
  const uint16_t* g     = a.geo + (size_t)f * a.M * plane;
  const uint32_t* bm    = a.bitmap + (size_t)f * a.H * a.bmWords;
  const uint32_t* b2p   = a.b2p + (size_t)f * a.Hb * a.Wb;
  IlvPixel        r{};
  r.n0 = (int16_t)normal_coord( p, g[(size_t)y * a.W + x] );
  // neighbours in the reference's order: left, right, top, bottom (:380-433)
  const int nx[4] = {x - 1, x + 1, x, x}, ny[4] = {y, y, y - 1, y + 1};
  double    dn[4] = {0.0, 0.0, 0.0, 0.0}, mn = (double)r.n0, mx = (double)r.n0;
  int       count = 0;
#pragma unroll
  for ( int k = 0; k < 4; k++ ) {
    if ( nx[k] < 0 || ny[k] < 0 || nx[k] >= a.W || ny[k] >= a.H ) { continue; }           // :362-365
    if ( !( ( bm[(size_t)ny[k] * a.bmWords + ( nx[k] >> 5 ) ] >> ( nx[k] & 31 ) ) & 1u ) ) { continue; }
    if ( b2p[( ny[k] >> 4 ) * a.Wb + ( nx[k] >> 4 )] != (uint32_t)p.frame_patch + 1u ) { continue; }
    const long long v = g[(size_t)ny[k] * a.W + nx[k]];
    // size_t arithmetic: d1 - value wraps for value > d1 and becomes a huge double (:385-389)
    dn[k] = p.mode == 0 ? (double)( v + p.d1 ) : (double)(unsigned long long)( (long long)p.d1 - v );
    count++;
    mn = fmin( mn, dn[k] );
    mx = fmax( mx, dn[k] );
  }
  if ( count == 0 ) { return r; }
  const double st = (double)a.surface_thickness, own = (double)r.n0;
  double       other;
  if ( ( x + y ) & 1 ) {  // the coded point is D1, D0 is interpolated (:435-446)
    other = p.mode == 0 ? round( fmin( fmax( mn, own - st ), own ) ) : round( fmax( fmin( mx, own + st ), own ) );
  } else {                // the coded point is D0, D1 is the clamped neighbour mean (:447-462)
    const double avg = ( ( ( dn[0] + dn[1] ) + dn[2] ) + dn[3] ) / (double)count;
    other = p.mode == 0 ? round( fmax( fmin( avg, own + st ), own ) ) : round( fmin( fmax( avg, own - st ), own ) );
  }
  r.two  = 1;
  r.n1   = (int16_t)(int)other;
  const int lo = min( r.n0, r.n1 ), hi = max( r.n0, r.n1 );
  r.xmin  = lo;
  r.nfill = max( hi - lo - 1, 0 );  // for ( step = 1; step < xmax - xmin; ++step ), :463-468
  return r;
}

// pointLocalReconstruction (generatePoints :472-496, getDeltaNeighbors :238-264): the coded point, a second point
// deltaDepth along the normal (the largest depth step within the threshold inside a (2n+1)^2 window of the geometry
// frame, occupied or not, never less than minD1), and optionally the points in between.
__device__ __forceinline__ IlvPixel plr_pixel( const ReprojArgs& a, const RbPatch& p, int f, int x, int y, const rb200_plr_mode m ) {
  const size_t    plane = (size_t)a.W * a.H;
  const uint16_t* g     = a.geo + (size_t)f * a.M * plane;
  IlvPixel        r{};
  const int dOrg = normal_coord( p, g[(size_t)y * a.W + x] );  // generateNormalCoordinate: a double holding an integer
  r.n0           = (int16_t)dOrg;
  int delta = 0;
  if ( m.interpolate ) {
    const int nb = m.neighbor, thr = 4;  // g_neighborThreshold, PCCCommon.h:127
    const int x0 = max( 0, x - nb ), x1 = min( x + nb, a.W ), y0 = max( 0, y - nb ), y1 = min( y + nb, a.H );
    for ( int xx = x0; xx <= x1; xx++ ) {    // both bounds are inclusive in the reference (:249-252): x == W reads
      for ( int yy = y0; yy <= y1; yy++ ) {  // the first sample of the next row; past the last row it reads beyond
        const size_t idx = (size_t)yy * a.W + xx;  // the plane, which has no defined value: skipped here
        if ( idx >= plane ) { continue; }
        const int d = normal_coord( p, g[idx] ) - dOrg;
        if ( p.mode == 0 ) {
          if ( d <= thr && d > delta ) { delta = d; }
        } else {
          if ( d >= -thr && d < delta ) { delta = d; }
        }
      }
    }
    if ( delta != 0 ) { delta += p.mode == 0 ? -1 : 1; }  // :261
  }
  delta = p.mode == 0 ? max( delta, (int)m.min_d1 ) : min( delta, -(int)m.min_d1 );  // :478-482
  if ( delta == 0 ) { return r; }
  r.two = 1;
  r.n1  = (int16_t)( r.n0 + delta );
  if ( m.filling ) {  // size_t xmin / xmax (:487-488): a negative minimum wraps, and -1 + 1 wraps back to 0
    const int lo = min( r.n0, r.n1 ), hi = max( r.n0, r.n1 );
    if ( lo >= 0 ) {
      r.xmin  = lo;
      r.nfill = max( hi - lo - 1, 0 );
    } else if ( lo == -1 && hi >= 0 ) {
      r.xmin  = -1;
      r.nfill = hi;
    }
  }
  return r;
}

template <bool EMIT, bool PLR>
__global__ void __launch_bounds__( WARPS * 32 ) k_reproject_interleaved( const ReprojArgs a ) {
  const int     lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int64_t wi   = (int64_t)blockIdx.x * WARPS + wq;
  if ( wi >= a.nWI ) { return; }
  const RbPatch p  = a.patches[a.wi_patch[wi]];
  const int     lb = a.wi_local[wi];
  const int     ub = lb % p.su0, vb = lb / p.su0;
  int           bx, by;
  canvas_block( p, ub, vb, bx, by );
  const int  f     = p.frame;
  const bool owned = a.b2p[( (size_t)f * a.Hb + by ) * a.Wb + bx] == (uint32_t)p.frame_patch + 1u;  // :650
  if ( !owned ) {
    if ( !EMIT && lane == 0 ) { a.wi_count[wi] = 0; }
    return;
  }
  const int       X0 = bx * 16, Y0 = by * 16;
  const TileMap   tm = tile_map( p.orient );
  const int       v1 = lane >> 1, ubase = 8 * ( lane & 1 );
  const uint32_t* bm = a.bitmap + (size_t)f * a.H * a.bmWords;
  const size_t    plane = (size_t)a.W * a.H;
  rb200_plr_mode  pm{};  // context.getPointLocalReconstructionMode( patch.getPointLocalReconstructionMode( u0, v0 ) ), :783-784
  if ( PLR ) { pm = a.plr_modes[a.plr_block_mode[a.plr_block_off[a.wi_patch[wi]] + lb]]; }
  // ---- per-lane counts, ordered scan (lane order == emission order v1, u1) ----
  int myCount = 0;
  for ( int j = 0; j < 8; j++ ) {
    int tx, ty;
    TILE_XY( tm, ubase + j, v1, tx, ty );
    const int x = X0 + tx, y = Y0 + ty;
    if ( !( ( bm[(size_t)y * a.bmWords + ( x >> 5 )] >> ( x & 31 ) ) & 1u ) ) { continue; }
    const IlvPixel q = PLR ? plr_pixel( a, p, f, x, y, pm ) : interleave_pixel( a, p, f, x, y );
    myCount += 1 + ( q.two && !( a.remove_dup && q.n1 == q.n0 ) ? 1 : 0 ) + ( q.two ? q.nfill : 0 );
  }
  int incl = myCount;
#pragma unroll
  for ( int d = 1; d < 32; d <<= 1 ) {
    const int t = __shfl_up_sync( 0xFFFFFFFFu, incl, d );
    if ( lane >= d ) { incl += t; }
  }
  if ( !EMIT ) {
    if ( lane == 31 ) { a.wi_count[wi] = incl; }
    return;
  }
  int64_t o    = a.frame_off[f] + ( a.wi_base[wi] - a.wi_base[a.frame_wi_off[f]] ) + ( incl - myCount );
  int     maxc = 0;
  for ( int j = 0; j < 8; j++ ) {
    int       tx, ty;
    const int u1 = ubase + j;
    TILE_XY( tm, u1, v1, tx, ty );
    const int x = X0 + tx, y = Y0 + ty;
    if ( !( ( bm[(size_t)y * a.bmWords + ( x >> 5 )] >> ( x & 31 ) ) & 1u ) ) { continue; }
    const IlvPixel q  = PLR ? plr_pixel( a, p, f, x, y, pm ) : interleave_pixel( a, p, f, x, y );
    const int      u  = ub * 16 + u1, v = vb * 16 + v1;
    const int      np = 1 + ( q.two ? 1 + q.nfill : 0 );
    for ( int i = 0; i < np; i++ ) {  // createdPoints: coded, interpolated, fills (:781-835)
      if ( i == 1 && a.remove_dup && q.n1 == q.n0 ) { continue; }  // :794-795 (a fill never equals the coded point)
      const int nn = i == 0 ? q.n0 : ( i == 1 ? q.n1 : q.xmin + ( i - 1 ) );
      int16_t   Q[3] = {0, 0, 0};
      set_axis( Q, p.normal_axis, nn );
      set_axis( Q, p.tangent_axis, u * p.lodx + p.u1 );
      set_axis( Q, p.bitangent_axis, v * p.lody + p.v1 );
      if ( p.addplane ) { inverse_rotate45( p.addplane, a.bitdepth3d, Q ); }
      // pointToPixel layer (:821-825): the coded point sits on its own checkerboard layer, the interpolated one on
      // the other, fills on g_intermediateLayerIndex (PCCCommon.h:126)
      // (:826-828) point local reconstruction: 0, g_intermediateLayerIndex, g_intermediateLayerIndex + 1
      const int layer = PLR ? ( i == 0 ? 0 : ( i == 1 ? 100 : 101 ) )
                            : ( i == 0 ? ( ( x + y ) & 1 ) : ( i == 1 ? ( ( x + y + 1 ) & 1 ) : 100 ) );
      a.pos[o]        = make_short4( Q[0], Q[1], Q[2], 0 );
      ushort4 cv      = make_ushort4( 0, 0, 0, (unsigned short)layer );
      if ( i == 0 && a.attr_count > 0 ) {  // colorPointCloud :1367-1374: only the coded point reads the attribute frame
        const uint16_t* at = a.attr + (size_t)f * a.M * 3 * plane + (size_t)y * a.W + x;
        cv                 = make_ushort4( at[0], at[plane], at[2 * plane], (unsigned short)layer );
      }
      a.col[o]  = cv;
      a.pix[o]  = (uint32_t)x | ( (uint32_t)y << 16 );
      a.part[o] = (uint32_t)p.frame_patch;
      maxc      = max( maxc, max( (int)Q[0], max( (int)Q[1], (int)Q[2] ) ) );
      o++;
    }
  }
#pragma unroll
  for ( int d = 16; d > 0; d >>= 1 ) { maxc = max( maxc, __shfl_xor_sync( 0xFFFFFFFFu, maxc, d ) ); }
  if ( lane == 0 && maxc > 0 ) { atomicMax( &a.finfo[f].max_coord, maxc ); }
}

// ---------------------------------------------------------------------------------------------------
// exclusive scan of int32 counts into int64 bases (n+1 outputs), n is O(1e5..1e6): two launches.
//   k_scan_tiles : every CTA scans one tile of 2048 counts (16-byte loads, shuffle scans) and publishes the tile sum
//   k_scan_apply : every CTA reduces the sums of the tiles before it (at most a few hundred) and adds the offset
// ---------------------------------------------------------------------------------------------------
constexpr int SCAN_TILE = 2048;  // 256 threads x 8

__device__ __forceinline__ int64_t block_exclusive_256( int64_t v, int64_t* warpSum, int64_t& total ) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int64_t   incl = v;
#pragma unroll
  for ( int d = 1; d < 32; d <<= 1 ) {
    const int64_t t = __shfl_up_sync( 0xFFFFFFFFu, incl, d );
    if ( lane >= d ) { incl += t; }
  }
  if ( lane == 31 ) { warpSum[w] = incl; }
  __syncthreads();
  int64_t off = 0, tot = 0;
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {
    const int64_t x = warpSum[k];
    if ( k < w ) { off += x; }
    tot += x;
  }
  total = tot;
  __syncthreads();
  return off + incl - v;
}

__global__ void __launch_bounds__( 256 ) k_scan_tiles( const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out,
                                                       int64_t* __restrict__ tileSum ) {
  __shared__ int64_t warpSum[8];
  const int64_t      i0 = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 8;
  int32_t            v[8];
  if ( i0 + 8 <= n ) {
    const int4 a = *reinterpret_cast<const int4*>( in + i0 ), b = *reinterpret_cast<const int4*>( in + i0 + 4 );
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  } else {
#pragma unroll
    for ( int k = 0; k < 8; k++ ) { v[k] = i0 + k < n ? in[i0 + k] : 0; }
  }
  int64_t s = 0;
#pragma unroll
  for ( int k = 0; k < 8; k++ ) { s += v[k]; }
  int64_t total;
  int64_t run = block_exclusive_256( s, warpSum, total );
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {
    if ( i0 + k < n ) { out[i0 + k] = run; }
    run += v[k];
  }
  if ( threadIdx.x == 0 ) { tileSum[blockIdx.x] = total; }
}

__global__ void __launch_bounds__( 256 ) k_scan_apply( int64_t n, int64_t* __restrict__ out, const int64_t* __restrict__ tileSum,
                                                       int nTiles ) {
  __shared__ int64_t warpSum[8];
  int64_t            s = 0;  // sum of the tiles before this one (the extra last CTA sums all: out[n])
  for ( int t = threadIdx.x; t < min( (int)blockIdx.x, nTiles ); t += 256 ) { s += tileSum[t]; }
  int64_t total;
  block_exclusive_256( s, warpSum, total );
  if ( (int)blockIdx.x == nTiles ) {
    if ( threadIdx.x == 0 ) { out[n] = total; }
    return;
  }
  if ( total == 0 ) { return; }
  const int64_t i0 = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 8;
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {
    if ( i0 + k < n ) { out[i0 + k] += total; }
  }
}

struct FrameLayout {
  int64_t off, regular, eom, raw;
};

// EOM segments: (EOM patch, member) pairs in the reference's append order (:849-882)
struct EomSeg {
  int32_t frame;
  int32_t patch;       // global patch index whose staged EOM points are copied
  int32_t eom_patch;   // index of the EOM patch inside the frame (restarts the synthetic pixel counter)
  int32_t first_in_eom_patch;
  int32_t u0, v0;      // origin of the synthetic pixel addresses (the EOM patch's; 0 with the auxiliary video, :852-853)
  int32_t cu0, cv0;    // auxiliary video: origin of the colour addresses (the EOM patch's, :1556-1557)
};

// per-patch staged EOM range = [wi_eom_base[firstWI(patch)], wi_eom_base[firstWI(patch)+nblocks))
__global__ void k_eom_segment_sizes( const EomSeg* __restrict__ segs, int nSeg, const int32_t* __restrict__ patch_first_wi,
                                     const int32_t* __restrict__ patch_nblk, const int64_t* __restrict__ wi_eom_base,
                                     int32_t* __restrict__ seg_size ) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= nSeg ) { return; }
  const int     q = segs[i].patch;
  const int64_t b = wi_eom_base[patch_first_wi[q]], e = wi_eom_base[patch_first_wi[q] + patch_nblk[q]];
  seg_size[i]     = (int32_t)( e - b );
}

// frame layout: offsets of every frame in the GOF arena = regular + eom + raw points, in that order (:841-949)
__global__ void k_frame_layout( int F, const int32_t* __restrict__ frame_wi_off, const int64_t* __restrict__ wi_base,
                                const EomSeg* __restrict__ segs, const int32_t* __restrict__ seg_size, int nSeg,
                                const int32_t* __restrict__ raw_per_frame, FrameLayout* __restrict__ out,
                                int64_t* __restrict__ frame_off, RbFrameInfo* __restrict__ finfo ) {
  // every frame's counts are fetched by its own thread (the loads are dependent chains), then one thread forms the
  // running offsets from shared memory
  __shared__ int64_t tot[256];
  if ( blockIdx.x != 0 ) { return; }
  for ( int f0 = 0; f0 < F; f0 += 256 ) {  // 256 frames per pass (blockDim.x), the offset is carried
    const int f = f0 + threadIdx.x;
    if ( f < F ) {
      FrameLayout L;
      L.off     = 0;
      L.regular = wi_base[frame_wi_off[f + 1]] - wi_base[frame_wi_off[f]];
      L.eom     = 0;
      for ( int s = 0; s < nSeg; s++ ) {
        if ( segs[s].frame == f ) { L.eom += seg_size[s]; }
      }
      L.raw  = raw_per_frame ? raw_per_frame[f] : 0;
      out[f] = L;
      tot[threadIdx.x] = L.regular + L.eom + L.raw;
      RbFrameInfo z{};
      finfo[f] = z;
    }
    __syncthreads();
    if ( threadIdx.x == 0 ) {
      int64_t run = f0 == 0 ? 0 : frame_off[f0];
      for ( int k = 0; k < min( 256, F - f0 ); k++ ) {
        frame_off[f0 + k] = run;
        out[f0 + k].off   = run;
        run += tot[k];
      }
      frame_off[min( F, f0 + 256 )] = run;
    }
    __syncthreads();
  }
}

// EOM append (:846-891): one CTA per segment; synthetic pixel addresses, occupancy marks, colours from
// attribute frame 0 at those addresses (colorPointCloud, :1365-1422 with f = 0)
__global__ void k_eom_append( const EomSeg* __restrict__ segs, const int32_t* __restrict__ seg_size,
                              const int64_t* __restrict__ seg_dst,  // offset of the segment inside its frame's EOM range
                              const int64_t* __restrict__ seg_pix,  // running totalPointCount inside the EOM patch
                              const int32_t* __restrict__ patch_first_wi, const int64_t* __restrict__ wi_eom_base,
                              const short4* __restrict__ stage, const FrameLayout* __restrict__ layout,
                              const int32_t* __restrict__ patch_count_of_frame, const uint16_t* __restrict__ attr,
                              uint32_t* __restrict__ bitmap, int W, int H, int Wb, int M, int words, int attr_count,
                              const uint16_t* __restrict__ aux_attr, int auxW, int auxH, int aux,
                              short4* __restrict__ pos, ushort4* __restrict__ col, uint32_t* __restrict__ pix,
                              uint32_t* __restrict__ part, RbFrameInfo* __restrict__ finfo ) {
  const int     s = blockIdx.x;
  const EomSeg  g = segs[s];
  const int     n = seg_size[s];
  const int64_t src = wi_eom_base[patch_first_wi[g.patch]];
  const int64_t dst = layout[g.frame].off + layout[g.frame].regular + seg_dst[s];
  int           maxc = 0;
  for ( int i = threadIdx.x; i < n; i += blockDim.x ) {
    const int64_t t   = seg_pix[s] + i;  // totalPointCount (:862-869)
    const int64_t blk = t / 256, inb = t - blk * 256;
    const int     uu  = (int)( blk % Wb ) * 16 + (int)( inb % 16 ) + g.u0;
    const int     vv  = (int)( blk / Wb ) * 16 + (int)( inb / 16 ) + g.v0;
    short4        P   = stage[src + i];
    P.w               = 0;
    pos[dst + i]      = P;
    ushort4 cv        = make_ushort4( 0, 0, 0, 0 );
    if ( aux ) {
      // generateRawPointsAttributefromVideo (:1551-1580): block of the EOM patch from the point's index k inside the patch,
      // pixel inside the block from a counter that runs over ALL EOM points of the tile (it is not reset per patch), in
      // the auxiliary attribute video; the values pass through 8-bit PCCColor3B (:1574-1576, :1437)
      if ( attr_count > 0 ) {
        const int     wbA = auxW / 16;
        const int64_t npx = ( seg_dst[s] + i ) % 256;
        const int     xx  = g.cu0 + (int)( blk % wbA ) * 16 + (int)( npx % 16 );
        const int     yy  = g.cv0 + (int)( blk / wbA ) * 16 + (int)( npx / 16 );
        // (getValue( c, xx, yy ) is channel[yy * width + xx] without a range check, PCCImage.h:205-212: a block row that
        // runs past the right edge continues in the next pixel row, as in the reference; past the plane: 0)
        const size_t plane = (size_t)auxW * auxH, o = (size_t)yy * auxW + xx, fb = (size_t)g.frame * 3 * plane;
        if ( o < plane ) {
          cv = make_ushort4( aux_attr[fb + o] & 0xFFu, aux_attr[fb + plane + o] & 0xFFu, aux_attr[fb + 2 * plane + o] & 0xFFu, 0 );
        }
      }
    } else if ( attr_count > 0 && uu < W && vv < H ) {
      const size_t plane = (size_t)W * H;
      const size_t o     = (size_t)vv * W + uu;
      const size_t fb    = ( (size_t)g.frame * M ) * 3 * plane;
      cv                 = make_ushort4( attr[fb + o], attr[fb + plane + o], attr[fb + 2 * plane + o], 0 );
    }
    col[dst + i]  = cv;
    pix[dst + i]  = (uint32_t)uu | ( (uint32_t)vv << 16 );
    part[dst + i] = (uint32_t)patch_count_of_frame[g.frame];  // partition = patches.size() (:843,877)
    if ( !aux && uu < W && vv < H ) { atomicOr( &bitmap[( (size_t)g.frame * H + vv ) * words + ( uu >> 5 )], 1u << ( uu & 31 ) ); }  // :880
    maxc = max( maxc, max( (int)P.x, max( (int)P.y, (int)P.z ) ) );
  }
  if ( maxc > 0 ) { atomicMax( &finfo[g.frame].max_coord, maxc ); }
}

// raw (missed-point) patches inside the geometry video (:894-949) + their colours (:1365-1422, f = 0)
struct RawDesc {
  int32_t frame, u0, v0, sizeU, sizeV, u1, v1, d1, n;
  int32_t pad;
  int64_t dst;  // offset inside the frame's raw range
};
// gW x gH: the frames the raw patches address — the atlas (map 0 of frame f: gM frames per atlas frame), or the auxiliary
// video (aux != 0, one frame per atlas frame; the colours pass through PCCColor3B there, PCCCodec.cpp:1541-1543, :1438)
__global__ void k_raw_points( const RawDesc* __restrict__ descs, const FrameLayout* __restrict__ layout,
                              const int32_t* __restrict__ patch_count_of_frame, const uint16_t* __restrict__ geo,
                              const uint16_t* __restrict__ attr, int gW, int gH, int gM, int aux, int attr_count,
                              short4* __restrict__ pos, ushort4* __restrict__ col, uint32_t* __restrict__ pix,
                              uint32_t* __restrict__ part, RbFrameInfo* __restrict__ finfo ) {
  const RawDesc d     = descs[blockIdx.y];
  const size_t  plane = (size_t)gW * gH;
  const uint16_t* g   = geo + ( (size_t)d.frame * gM ) * plane;
  const int64_t dst   = layout[d.frame].off + layout[d.frame].regular + layout[d.frame].eom + d.dst;
  int           maxc  = 0;
  for ( int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += gridDim.x * blockDim.x ) {
    int16_t Q[3];
    for ( int k = 0; k < 3; k++ ) {
      const int64_t t = (int64_t)k * d.n + i;  // numRawPointsAdded (:913-928): X || Y || Z runs
      const int     u = (int)( t % d.sizeU ), v = (int)( t / d.sizeU );
      const int     val = g[(size_t)( d.v0 + v ) * gW + d.u0 + u];
      Q[k]              = (int16_t)( val + ( k == 0 ? d.u1 : ( k == 1 ? d.v1 : d.d1 ) ) );
    }
    const int u = i % d.sizeU, v = i / d.sizeU;  // :930-941
    const int x = d.u0 + u, y = d.v0 + v;
    pos[dst + i] = make_short4( Q[0], Q[1], Q[2], 0 );
    ushort4 cv   = make_ushort4( 0, 0, 0, 0 );
    if ( attr_count > 0 ) {
      const size_t fb = ( (size_t)d.frame * gM ) * 3 * plane, o = (size_t)y * gW + x;
      cv              = make_ushort4( attr[fb + o], attr[fb + plane + o], attr[fb + 2 * plane + o], 0 );
      if ( aux ) { cv.x &= 0xFFu, cv.y &= 0xFFu, cv.z &= 0xFFu; }
    }
    col[dst + i]  = cv;
    pix[dst + i]  = (uint32_t)x | ( (uint32_t)y << 16 );
    part[dst + i] = (uint32_t)patch_count_of_frame[d.frame];
    maxc          = max( maxc, max( (int)Q[0], max( (int)Q[1], (int)Q[2] ) ) );
  }
  if ( maxc > 0 ) { atomicMax( &finfo[d.frame].max_coord, maxc ); }
}

// boundary classification as a separate pass (:953-973) — needed when EOM marks changed the occupancy map
// after the regular points were emitted.  Covers regular + EOM points (not raw points).
__global__ void k_classify_points( const FrameLayout* __restrict__ layout, int F, const uint32_t* __restrict__ bitmap,
                                   int W, int H, int words, const uint32_t* __restrict__ pix, short4* __restrict__ pos,
                                   uint32_t* __restrict__ blist, uint32_t* __restrict__ blist_n ) {
  const int          f = blockIdx.y;
  const FrameLayout  L = layout[f];
  const int64_t      n = L.regular + L.eom;
  const uint32_t*    bm = bitmap + (size_t)f * H * words;
  for ( int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x ) {
    const uint32_t pv = pix[L.off + i];
    const int      x = pv & 0xFFFF, y = pv >> 16;
    auto occ = [&]( int xx, int yy ) -> int {
      if ( xx < 0 || yy < 0 || xx >= W || yy >= H ) { return 1; }
      return ( bm[(size_t)yy * words + ( xx >> 5 )] >> ( xx & 31 ) ) & 1;
    };
    if ( x >= W || y >= H || !occ( x, y ) ) { continue; }
    int t = 0;
    if ( x == 0 || y == 0 || x == W - 1 || y == H - 1 ) {
      t = 1;
    } else {
      for ( int dy = -1; dy <= 1 && !t; dy++ ) {
        for ( int dx = -1; dx <= 1; dx++ ) {
          if ( !occ( x + dx, y + dy ) ) { t = 1; break; }
        }
      }
      if ( !t ) {
        for ( int dy = -2; dy <= 2 && !t; dy++ ) {
          for ( int dx = -2; dx <= 2; dx++ ) {
            if ( !occ( x + dx, y + dy ) ) { t = 1; break; }
          }
        }
      }
      if ( !t && ( x == 1 || y == 1 || x == W - 2 || y == H - 2 ) ) { t = 1; }
    }
    if ( t ) {
      pos[L.off + i].w                = 1;
      blist[atomicAdd( blist_n, 1u )] = (uint32_t)( L.off + i );
    }
  }
}

}  // namespace

// tensor map of a stack of u16 planes [planes][H][W], box 16 x 16 x boxPlanes (the tiles of one patch block)
static int encode_plane_map( rb200_ctx* c, CUtensorMap* map, void* base, int W, int H, int64_t planes, int boxPlanes ) {
  typedef CUresult ( *EncodeFn )( CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill );
  static EncodeFn encode = nullptr;
  if ( !encode ) {
    void*                           fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if ( cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q ) != cudaSuccess || !fn ) {
      return rb_fail( c, RB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver" );
    }
    encode = (EncodeFn)fn;
  }
  const cuuint64_t dims[3]    = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)std::max<int64_t>( planes, 1 )};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 2, (cuuint64_t)W * H * 2};
  const cuuint32_t box[3]     = {16, 16, (cuuint32_t)boxPlanes};
  const cuuint32_t estr[3]    = {1, 1, 1};
  const CUresult   r = encode( map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
  if ( r != CUDA_SUCCESS ) { return rb_fail( c, RB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for %dx%d x %lld planes", (int)r, W, H, (long long)planes ); }
  return RB200_OK;
}

int rb_reconstruct_impl( rb200_ctx* c ) {
  const rb200_params& P = c->P;
  const int           F = c->F;
  const bool          eom = P.enhanced_occupancy_map_code != 0;
  // pixel interleaving wins when both are set (generatePoints tests it first, :350 / :472)
  const bool          plr = P.point_local_reconstruction != 0 && !P.single_map_pixel_interleaving;
  const bool          ilv = P.single_map_pixel_interleaving != 0 || plr;  // the per-pixel variable-count path
  if ( plr && !c->have_plr ) { return rb_fail( c, RB200_ERR_STATE, "point_local_reconstruction: rb200_gof_set_plr was not called for this GOF" ); }
  const int64_t       nWI = c->nWI;
  RB_CUDA( cudaMemsetAsync( c->d_b2p.p, 0, (size_t)F * c->Wb * c->Hb * 4, c->stream ) );
  const bool pbf = P.pbf_enable != 0;
  {
    // occupancy synthesis: generateOccupancyMap is skipped (PCCDecoder.cpp:362-366), so block-to-patch sees the raw
    // video samples != 0 (:1757) — the same test as the EOM symbol
    const int64_t total = (int64_t)F * c->H * c->bmWords;
    RB_LAUNCH( "occupancy_bitmap", k_occupancy_bitmap, rb_div_up( total, 256 ), 256, 0, c->d_occ_video.as<uint8_t>(),
               c->d_bitmap.as<uint32_t>(), F, c->W, c->H, c->oW, c->oH, c->prec, c->bmWords, P.threshold_lossy_om,
               ( eom || pbf ) ? 1 : 0 );
  }
  if ( nWI > 0 ) {
    RB_LAUNCH( "block_to_patch", k_block_to_patch, rb_div_up( nWI, 256 ), 256, 0, c->d_patches.as<RbPatch>(),
               c->d_wi_patch.as<int32_t>(), c->d_wi_local.as<int32_t>(), nWI, c->d_bitmap.as<uint32_t>(),
               c->d_b2p.as<uint32_t>(), c->H, c->Wb, c->Hb, c->bmWords );
    if ( P.enable_size_quantization ) {
      RB_LAUNCH( "size_quantization", k_size_quantization, (unsigned)nWI, 256, 0, c->d_patches.as<RbPatch>(),
                 c->d_wi_patch.as<int32_t>(), c->d_wi_local.as<int32_t>(), nWI, c->d_bitmap.as<uint32_t>(),
                 c->d_b2p.as<uint32_t>(), c->H, c->Wb, c->Hb, c->bmWords, P.log2_quantizer_x, P.log2_quantizer_y );
    }
  }
  if ( pbf ) {  // the synthesised occupancy replaces the bitmap; the points' boundary types come with it
    int r = rb_pbf_impl( c );
    if ( r ) { return r; }
  }
  // geometry smoothing is only signalled through flagGeometrySmoothing_ (:953); occupancy synthesis types every point (:809)
  const bool classify = P.flag_geometry_smoothing != 0 || pbf;

  ReprojArgs a{};
  a.patches      = c->d_patches.as<RbPatch>();
  a.wi_patch     = c->d_wi_patch.as<int32_t>();
  a.wi_local     = c->d_wi_local.as<int32_t>();
  a.nWI          = nWI;
  a.occ_video    = c->d_occ_video.as<uint8_t>();
  a.geo          = c->d_geometry.as<uint16_t>();
  a.attr         = c->d_attribute.as<uint16_t>();
  a.bitmap       = c->d_bitmap.as<uint32_t>();
  a.bnd_bitmap   = pbf ? c->d_bnd_bitmap.as<uint32_t>() : nullptr;
  a.b2p          = c->d_b2p.as<uint32_t>();
  a.W            = c->W;
  a.H            = c->H;
  a.oW           = c->oW;
  a.oH           = c->oH;
  a.Wb           = c->Wb;
  a.Hb           = c->Hb;
  a.M            = c->M;
  a.prec         = c->prec;
  a.bmWords      = c->bmWords;
  a.absolute_d1  = P.absolute_d1;
  a.remove_dup   = P.remove_duplicate_points;
  a.eom_fix_bits = P.eom_fix_bit_count;
  a.classify     = ( classify && !eom && !ilv ) ? 1 : 0;  // with EOM the marks (:880) come first, classification after
  a.surface_thickness = P.surface_thickness;
  a.plr_modes         = c->d_plr_modes.as<rb200_plr_mode>();
  a.plr_block_mode    = c->d_plr_block_mode.as<uint8_t>();
  a.plr_block_off     = c->d_plr_block_off.as<int64_t>();
  a.attr_count   = P.attribute_count;
  // image0.getDeprecatedColorFormat() == 0 ? 8 : 16 (:1388): format 0 is the 4:4:4 (RGB) video, PCCVideoDecoder.cpp:132
  a.t1_bits      = ( P.multiple_streams && P.relative_t1 ) ? ( P.attribute_rgb444 ? 8 : 16 ) : 0;
  a.bitdepth3d   = P.geometry_bitdepth_3d;
  a.wi_count     = c->d_wi_count.as<int32_t>();
  a.wi_cnt4      = nullptr;
  if ( !eom && !ilv && nWI > 0 ) {  // 128 bytes per patch block: the emitting pass skips its own counting of the pixels
    RB_CUDA( c->d_wi_eom_slot.ensure( (size_t)nWI * 128 ) );
    a.wi_cnt4 = c->d_wi_eom_slot.as<uint32_t>();
  }
  a.frame_wi_off = c->d_frame_wi_off.as<int32_t>();
  a.frame_off    = c->d_frame_off.as<int64_t>();
  a.finfo        = c->d_frame_info.as<RbFrameInfo>();
  if ( eom ) {
    RB_CUDA( c->d_wi_eom_count.ensure( ( nWI + 1 ) * 4 ) );
    RB_CUDA( c->d_wi_eom_base.ensure( ( nWI + 1 ) * 8 ) );
    a.wi_eom_count = c->d_wi_eom_count.as<int32_t>();
  }
  const int G = rb_div_up( nWI, WARPS );
  // the planes as TMA tensors: geometry [F * M][H][W], attribute [F * M * 3][H][W]; a box is the tile stack of one block
  alignas( 64 ) CUtensorMap tmGeo{}, tmAttr{};
  if ( nWI > 0 && !ilv ) {
    int r = encode_plane_map( c, &tmGeo, c->d_geometry.p, c->W, c->H, (int64_t)F * c->M, c->M );
    if ( r ) { return r; }
    if ( P.attribute_count > 0 ) {
      r = encode_plane_map( c, &tmAttr, c->d_attribute.p, c->W, c->H, (int64_t)F * c->M * 3, c->M * 3 );
      if ( r ) { return r; }
    } else {
      tmAttr = tmGeo;
    }
  }
  if ( nWI > 0 ) {
    const bool std_cfg = !eom && !pbf && c->M == 2 && P.absolute_d1 && P.attribute_count > 0 && classify && a.t1_bits == 0;
    if ( ilv ) {
      auto kCount = plr ? k_reproject_interleaved<false, true> : k_reproject_interleaved<false, false>;
      RB_LAUNCH( "reproject_count", kCount, G, WARPS * 32, 0, a );
    } else {
      auto kCount = eom ? k_reproject<false, true, false> : ( std_cfg ? k_reproject<false, false, true> : k_reproject<false, false, false> );
      RB_LAUNCH( "reproject_count", kCount, G, WARPS * 32, 0, a, tmGeo, tmAttr );
    }
  }
  {
    const int nTiles = rb_div_up( nWI, SCAN_TILE );
    RB_CUDA( c->d_scratch[7].ensure( (size_t)( nTiles + 1 ) * 8 ) );
    int64_t* tileSum = c->d_scratch[7].as<int64_t>();
    RB_LAUNCH( "scan_counts", k_scan_tiles, std::max( nTiles, 1 ), 256, 0, c->d_wi_count.as<int32_t>(), nWI, c->d_wi_base.as<int64_t>(), tileSum );
    RB_LAUNCH( "scan_counts", k_scan_apply, nTiles + 1, 256, 0, nWI, c->d_wi_base.as<int64_t>(), tileSum, nTiles );
    if ( eom ) {
      RB_LAUNCH( "scan_counts", k_scan_tiles, std::max( nTiles, 1 ), 256, 0, c->d_wi_eom_count.as<int32_t>(), nWI, c->d_wi_eom_base.as<int64_t>(), tileSum );
      RB_LAUNCH( "scan_counts", k_scan_apply, nTiles + 1, 256, 0, nWI, c->d_wi_eom_base.as<int64_t>(), tileSum, nTiles );
    }
  }
  a.wi_base     = c->d_wi_base.as<int64_t>();
  a.wi_eom_base = c->d_wi_eom_base.as<int64_t>();

  // ---- host-side tables for EOM segments / raw patches / per-frame patch counts ----
  std::vector<EomSeg>  segs;
  std::vector<int32_t> patchFirstWI( c->h_patches.size() + 1, 0 ), patchNblk( c->h_patches.size() + 1, 0 );
  std::vector<int32_t> patchCountOfFrame( F, 0 ), rawPerFrame( F, 0 );
  std::vector<RawDesc> raws;
  {
    // first work item of every patch (work items are laid out per frame in emission order)
    int64_t wi = 0;
    for ( int f = 0; f < F; f++ ) {
      const int b = c->h_patch_off[f], e = c->h_patch_off[f + 1];
      patchCountOfFrame[f] = e - b;
      for ( int k = 0; k < e - b; k++ ) {
        const int i     = P.patch_precedence_reverse ? ( e - 1 - k ) : ( b + k );
        patchFirstWI[i] = (int32_t)wi;
        patchNblk[i]    = c->h_patches[i].size_u0 * c->h_patches[i].size_v0;
        wi += patchNblk[i];
      }
    }
  }
  if ( eom ) {
    for ( int f = 0; f < F; f++ ) {
      const int pb = c->h_patch_off[f], np = c->h_patch_off[f + 1] - pb;
      for ( int j = c->h_eom_off[f]; j < c->h_eom_off[f + 1]; j++ ) {
        const rb200_eom_patch& e = c->h_eom[j];
        for ( int k = 0; k < e.member_count; k++ ) {
          int m = c->h_eom_members[e.member_begin + k];
          if ( P.patch_precedence_reverse ) { m = np - m - 1; }  // :857-859
          if ( m < 0 || m >= np ) { return rb_fail( c, RB200_ERR_INVALID, "EOM member patch %d out of range", m ); }
          EomSeg s{};
          s.frame              = f;
          s.patch              = pb + m;
          s.eom_patch          = j - c->h_eom_off[f];
          s.first_in_eom_patch = ( k == 0 );
          const bool auxEom    = P.use_aux_separate_video != 0;
          s.u0                 = auxEom ? 0 : e.u0 * c->R;
          s.v0                 = auxEom ? 0 : e.v0 * c->R;
          s.cu0                = e.u0 * c->R;
          s.cv0                = e.v0 * c->R;
          segs.push_back( s );
        }
      }
    }
  }
  if ( P.use_additional_points_patch ) {
    for ( int f = 0; f < F; f++ ) {
      int64_t run = 0;
      for ( int j = c->h_raw_off[f]; j < c->h_raw_off[f + 1]; j++ ) {
        const rb200_raw_patch& r = c->h_raw[j];
        RawDesc                d{};
        d.frame = f;
        d.u0    = r.u0 * c->R;
        d.v0    = r.v0 * c->R;
        d.sizeU = r.size_u0 * c->R;
        d.sizeV = r.size_v0 * c->R;
        d.u1    = r.u1;
        d.v1    = r.v1;
        d.d1    = r.d1;
        d.n     = r.num_points;
        d.dst   = run;
        run += r.num_points;
        raws.push_back( d );
      }
      rawPerFrame[f] = (int32_t)run;
    }
  }
  const int nSeg = (int)segs.size();
  // scratch layout: [segs][segSize][patchFirstWI][patchNblk][patchCountOfFrame][rawPerFrame][layout][raws]
  const size_t szSegs = nSeg * sizeof( EomSeg ), szSegSize = ( nSeg + 1 ) * 4, szPf = patchFirstWI.size() * 4,
               szPc = F * 4, szLayout = F * sizeof( FrameLayout ), szRaw = raws.size() * sizeof( RawDesc );
  auto   al   = []( size_t x ) { return ( x + 255 ) & ~size_t( 255 ); };
  size_t o    = 0;
  const size_t oSegs = o; o += al( szSegs );
  const size_t oSegSize = o; o += al( szSegSize );
  const size_t oPfw = o; o += al( szPf );
  const size_t oPnb = o; o += al( szPf );
  const size_t oPc = o; o += al( szPc );
  const size_t oRpf = o; o += al( szPc );
  const size_t oLayout = o; o += al( szLayout );
  const size_t oRaw = o; o += al( szRaw );
  const size_t oSegDst = o; o += al( ( nSeg + 1 ) * 8 );
  const size_t oSegPix = o; o += al( ( nSeg + 1 ) * 8 );
  RB_CUDA( c->d_scratch[0].ensure( o + 256 ) );
  char* dS = c->d_scratch[0].as<char>();
  {
    std::vector<char> hs( o, 0 );
    if ( nSeg ) { memcpy( hs.data() + oSegs, segs.data(), szSegs ); }
    memcpy( hs.data() + oPfw, patchFirstWI.data(), szPf );
    memcpy( hs.data() + oPnb, patchNblk.data(), szPf );
    memcpy( hs.data() + oPc, patchCountOfFrame.data(), szPc );
    memcpy( hs.data() + oRpf, rawPerFrame.data(), szPc );
    if ( !raws.empty() ) { memcpy( hs.data() + oRaw, raws.data(), szRaw ); }
    RB_CUDA( cudaMemcpyAsync( dS, hs.data(), o, cudaMemcpyHostToDevice, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );  // hs is a stack-lifetime pageable buffer
    c->stats.h2d_bytes += (int64_t)o;
  }
  if ( nSeg ) {
    RB_LAUNCH( "eom_segment_sizes", k_eom_segment_sizes, rb_div_up( nSeg, 128 ), 128, 0, (const EomSeg*)( dS + oSegs ), nSeg,
               (const int32_t*)( dS + oPfw ), (const int32_t*)( dS + oPnb ), c->d_wi_eom_base.as<int64_t>(),
               (int32_t*)( dS + oSegSize ) );
  }
  RB_LAUNCH( "frame_layout", k_frame_layout, 1, 256, 0, F, c->d_frame_wi_off.as<int32_t>(), c->d_wi_base.as<int64_t>(),
             (const EomSeg*)( dS + oSegs ), (const int32_t*)( dS + oSegSize ), nSeg, (const int32_t*)( dS + oRpf ),
             (FrameLayout*)( dS + oLayout ), c->d_frame_off.as<int64_t>(), c->d_frame_info.as<RbFrameInfo>() );

  // ---- read the layout back (the caller needs the counts; the arena may need to grow) ----
  const size_t rbBytes = szLayout + ( nSeg + 1 ) * 4;
  char*        hp      = (char*)rb_pinned( c, rbBytes + 64 );
  if ( !hp ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  RB_CUDA( cudaMemcpyAsync( hp, dS + oLayout, szLayout, cudaMemcpyDeviceToHost, c->stream ) );
  if ( nSeg ) { RB_CUDA( cudaMemcpyAsync( hp + szLayout, dS + oSegSize, nSeg * 4, cudaMemcpyDeviceToHost, c->stream ) ); }
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  c->stats.d2h_bytes += (int64_t)rbBytes;
  const FrameLayout* L = (const FrameLayout*)hp;
  int64_t            total = 0;
  for ( int f = 0; f < F; f++ ) {
    c->h_frame_off[f]      = L[f].off;
    c->h_counts[f]         = rb200_frame_counts{};
    c->h_counts[f].regular = L[f].regular;
    c->h_counts[f].eom     = 0;  // filled below: TotalNumberOfEOMPoints is the sum of eomCount_ (:854,886)
    c->h_counts[f].raw     = L[f].raw;
    c->h_counts[f].total   = L[f].regular + L[f].eom + L[f].raw;
    total                  = L[f].off + c->h_counts[f].total;
  }
  c->h_frame_off[F] = total;
  if ( eom ) {
    for ( int f = 0; f < F; f++ ) {
      int64_t s = 0;
      for ( int j = c->h_eom_off[f]; j < c->h_eom_off[f + 1]; j++ ) { s += c->h_eom[j].eom_count; }
      c->h_counts[f].eom = s;
    }
  }
  const size_t cap = (size_t)std::max<int64_t>( total, 1 );
  RB_CUDA( c->d_pos.ensure( cap * 8 ) );
  RB_CUDA( c->d_col.ensure( cap * 8 ) );
  RB_CUDA( c->d_pix.ensure( cap * 4 ) );
  RB_CUDA( c->d_part.ensure( cap * 4 ) );
  RB_CUDA( c->d_rgb.ensure( cap * 4 ) );
  a.pos  = c->d_pos.as<short4>();
  a.col  = c->d_col.as<ushort4>();
  a.pix  = c->d_pix.as<uint32_t>();
  a.part = c->d_part.as<uint32_t>();
  RB_CUDA( c->d_blist.ensure( cap * 4 ) );
  RB_CUDA( c->d_blist_n.ensure( 64 ) );
  RB_CUDA( cudaMemsetAsync( c->d_blist_n.p, 0, 4, c->stream ) );
  RB_CUDA( c->d_moved_bits.ensure( ( cap / 32 + 2 ) * 4 ) );  // no point of a fresh reconstruction is of type 3
  RB_CUDA( cudaMemsetAsync( c->d_moved_bits.p, 0, ( cap / 32 + 2 ) * 4, c->stream ) );
  a.blist   = c->d_blist.as<uint32_t>();
  a.blist_n = c->d_blist_n.as<uint32_t>();
  if ( eom ) {
    // staged EOM extras (per patch, emission order)
    int64_t* hb = (int64_t*)rb_pinned( c, 64 );
    RB_CUDA( cudaMemcpyAsync( hb, c->d_wi_eom_base.as<int64_t>() + nWI, 8, cudaMemcpyDeviceToHost, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
    const int64_t nStage = *hb;
    RB_CUDA( c->d_scratch[1].ensure( (size_t)std::max<int64_t>( nStage, 1 ) * 8 ) );
    a.eom_stage = c->d_scratch[1].as<short4>();
    // re-read: the pinned block was reused, restore the segment sizes view
    RB_CUDA( cudaMemcpyAsync( hp, dS + oLayout, szLayout, cudaMemcpyDeviceToHost, c->stream ) );
    if ( nSeg ) { RB_CUDA( cudaMemcpyAsync( hp + szLayout, dS + oSegSize, nSeg * 4, cudaMemcpyDeviceToHost, c->stream ) ); }
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
  }
  if ( nWI > 0 ) {
    const bool std_cfg = !eom && !pbf && c->M == 2 && P.absolute_d1 && P.attribute_count > 0 && classify && a.t1_bits == 0;
    if ( ilv ) {
      auto kEmit = plr ? k_reproject_interleaved<true, true> : k_reproject_interleaved<true, false>;
      RB_LAUNCH( "reproject_emit", kEmit, G, WARPS * 32, 0, a );
    } else {
      auto kEmit = eom ? k_reproject<true, true, false> : ( std_cfg ? k_reproject<true, false, true> : k_reproject<true, false, false> );
      RB_LAUNCH( "reproject_emit", kEmit, G, WARPS * 32, 0, a, tmGeo, tmAttr );
    }
  }
  if ( eom && nSeg ) {
    // per-segment destination offset inside the frame's EOM range and synthetic pixel counter (resets per EOM patch)
    const int32_t*       segSize = (const int32_t*)( hp + szLayout );
    std::vector<int64_t> segDst( nSeg + 1, 0 ), segPix( nSeg + 1, 0 );
    int64_t              runF = 0, runP = 0;
    for ( int s = 0; s < nSeg; s++ ) {
      if ( s == 0 || segs[s].frame != segs[s - 1].frame ) { runF = 0; }
      if ( segs[s].first_in_eom_patch ) { runP = 0; }
      segDst[s] = runF;
      segPix[s] = runP;
      runF += segSize[s];
      runP += segSize[s];
    }
    if ( P.use_aux_separate_video ) {
      // the colours are addressed by the signalled eomCount_ of the patches (:1561): it must be what the patches produce
      for ( int s = 0; s < nSeg; ) {
        int64_t   got = 0;
        const int f = segs[s].frame, j = segs[s].eom_patch;
        for ( ; s < nSeg && segs[s].frame == f && segs[s].eom_patch == j; s++ ) { got += segSize[s]; }
        if ( got != c->h_eom[c->h_eom_off[f] + j].eom_count ) {
          return rb_fail( c, RB200_ERR_INVALID, "EOM patch %d of frame %d signals %d points but its member patches produce %lld", j, f,
                          c->h_eom[c->h_eom_off[f] + j].eom_count, (long long)got );
        }
      }
    }
    RB_CUDA( cudaMemcpyAsync( dS + oSegDst, segDst.data(), nSeg * 8, cudaMemcpyHostToDevice, c->stream ) );
    RB_CUDA( cudaMemcpyAsync( dS + oSegPix, segPix.data(), nSeg * 8, cudaMemcpyHostToDevice, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
    RB_LAUNCH( "eom_append", k_eom_append, nSeg, 256, 0, (const EomSeg*)( dS + oSegs ), (const int32_t*)( dS + oSegSize ),
               (const int64_t*)( dS + oSegDst ), (const int64_t*)( dS + oSegPix ), (const int32_t*)( dS + oPfw ),
               c->d_wi_eom_base.as<int64_t>(), c->d_scratch[1].as<short4>(), (const FrameLayout*)( dS + oLayout ),
               (const int32_t*)( dS + oPc ), c->d_attribute.as<uint16_t>(), c->d_bitmap.as<uint32_t>(), c->W, c->H, c->Wb,
               c->M, c->bmWords, P.attribute_count, c->d_aux_attr.as<uint16_t>(), P.aux_width, P.aux_height,
               P.use_aux_separate_video ? 1 : 0, a.pos, a.col, a.pix, a.part, c->d_frame_info.as<RbFrameInfo>() );
  }
  if ( !raws.empty() ) {
    const bool aux = P.use_aux_separate_video != 0;
    RB_LAUNCH( "raw_points", k_raw_points, dim3( 64, (unsigned)raws.size() ), 256, 0, (const RawDesc*)( dS + oRaw ),
               (const FrameLayout*)( dS + oLayout ), (const int32_t*)( dS + oPc ),
               aux ? c->d_aux_geo.as<uint16_t>() : c->d_geometry.as<uint16_t>(),
               aux ? c->d_aux_attr.as<uint16_t>() : c->d_attribute.as<uint16_t>(), aux ? P.aux_width : c->W,
               aux ? P.aux_height : c->H, aux ? 1 : c->M, aux ? 1 : 0, P.attribute_count, a.pos, a.col, a.pix, a.part,
               c->d_frame_info.as<RbFrameInfo>() );
  }
  if ( classify && ( eom || ilv ) ) {
    RB_LAUNCH( "classify_points", k_classify_points, dim3( 256, F ), 256, 0, (const FrameLayout*)( dS + oLayout ), F,
               c->d_bitmap.as<uint32_t>(), c->W, c->H, c->bmWords, a.pix, a.pos, a.blist, a.blist_n );
  }
  {  // length of the boundary list (launch size of the smoothing filters)
    uint32_t* hb = (uint32_t*)rb_pinned( c, 64 );
    RB_CUDA( cudaMemcpyAsync( hb, a.blist_n, 4, cudaMemcpyDeviceToHost, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
    c->blist_cap = hb[0];
  }
  // pixel interleaving: the interpolated and the fill points take their colour from the coded ones (:1367-1434)
  if ( ilv && P.attribute_count > 0 ) { return rb_interleave_colors_impl( c ); }
  return RB200_OK;
}
