// rb_smooth.cu — grid-based geometry smoothing, colour smoothing and YUV16->RGB8 for a whole GOF (sm_100a).
//
// Restates
//   PCCCodec::smoothPointCloudPostprocess + addGridCentroid      PccLibCommon/source/PCCCodec.cpp:52-147, :980-998
//   PCCCodec::smoothPointCloudGrid + gridFiltering               :1065-1104, :1000-1063
//   PCCCodec::colorSmoothing + addGridColorCentroid              :149-236, :1159-1180
//   PCCCodec::gridFilteringColor + smoothPointCloudColorLC       :1182-1306   (median/mean: PCCCodec.h:271-285)
//   PCCPointSet3::convertYUV16ToRGB8 / copyRGB16ToRGB8           PccLibCommon/include/PCCPointSet.h:121-166
//
// The reference walks a dense int grid (w^3 ints, memset per frame) and appends to std::vectors.  Here:
//  * a dense int32 index grid per frame stays resident and *clean* between GOFs: the mark pass records every
//    cell it claims, and a cleanup pass resets exactly those cells and their accumulators, so no per-GOF memset
//    of w^3 cells (8 MB .. 537 MB per frame) is ever paid;
//  * accumulators are integer atomics.  The reference sums small integers in float in emission order; below 2^24
//    every partial sum is exact, so the order-free integer sum converts to the identical float;
//    a per-frame flag reports any cell that leaves that range;
//  * "doSmooth" (a cell holds two different partitions) is kept order-free as max(p+1) and max(~(p+1));
//  * the filter passes repeat the reference's double arithmetic operation by operation (compiled with
//    -fmad=false), including the integer-truncating abs() of the colour gates.
#include <algorithm>

#include "rb_common.cuh"

namespace {

struct Cell {  // compact accumulator of one marked cell; all-zero == empty
  uint32_t cnt;
  uint32_t s0, s1, s2;   // coordinate sums (geometry) / colour sums (colour)
  uint32_t pmax, pinv;   // max(partition+1), max(0xFFFFFFFF - (partition+1))
  uint32_t lum_off;      // colour: start of the cell's luma list;   geometry: unused
  uint32_t aux;          // colour: scatter cursor, then the mean/median gate flag
};

struct GridArgs {
  int            F;
  int            g;          // cell size
  int            wmax;       // dense grid stride (cells per axis)
  int            by_bbox;    // geometry: th = g * ceil(maxCoord / g); colour: th = 2^bitdepth
  int            pcmax;      // 2^geometryBitDepth3D
  int32_t*       grid;       // [F][wmax^3]
  Cell*          cells;
  uint64_t*      cell_addr;  // grid address of every claimed cell (for cleanup)
  int64_t        cell_cap;
  int32_t*       counters;   // [0] cells claimed, [1] overflow flag, [2] luma cursor
  const int64_t* frame_off;
  RbFrameInfo*   finfo;
  short4*        pos;
  ushort4*       col;
  const uint32_t* part;
};

__device__ __forceinline__ int frame_of( const int64_t* __restrict__ off, int F, int64_t i ) {
  int lo = 0, hi = F - 1;
  while ( lo < hi ) {
    const int mid = ( lo + hi + 1 ) >> 1;
    if ( off[mid] <= i ) {
      lo = mid;
    } else {
      hi = mid - 1;
    }
  }
  return lo;
}

__device__ __forceinline__ int grid_th( const GridArgs& a, int f ) {
  if ( !a.by_bbox ) { return a.pcmax; }
  const int w = ( a.finfo[f].max_coord + a.g - 1 ) / a.g;  // :77-79
  return a.g * min( w, a.wmax );                             // :88 (coordinates >= 2^bitdepth are invalid input)
}

__device__ __forceinline__ bool inside( int x, int y, int z, int disth, int th ) {  // :92-95, :175-178
  return !( x < disth || y < disth || z < disth || th <= x + disth || th <= y + disth || th <= z + disth );
}

// ---- pass 1: type-1 points inside the margin claim their 2x2x2 cell neighbourhood (:89-112, :171-195) ----
__global__ void k_mark_cells( const GridArgs a, int64_t n ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  const short4 p = a.pos[i];
  if ( p.w != 1 ) { return; }
  const int f     = frame_of( a.frame_off, a.F, i );
  const int disth = max( a.g / 2, 1 ), th = grid_th( a, f );
  if ( !inside( p.x, p.y, p.z, disth, th ) ) { return; }
  const int     g  = a.g, hg = g / 2;
  const int     qx = p.x / g + ( ( p.x % g < hg ) ? -1 : 0 );
  const int     qy = p.y / g + ( ( p.y % g < hg ) ? -1 : 0 );
  const int     qz = p.z / g + ( ( p.z % g < hg ) ? -1 : 0 );
  const int64_t w  = a.wmax;
  int32_t*      G  = a.grid + (size_t)f * w * w * w;
  for ( int k = 0; k < 8; k++ ) {
    const int64_t cid = ( qx + ( k & 1 ) ) + ( qy + ( ( k >> 1 ) & 1 ) ) * w + ( qz + ( k >> 2 ) ) * w * w;
    if ( G[cid] != -1 ) { continue; }
    if ( atomicCAS( &G[cid], -1, -2 ) == -1 ) {
      const int idx = atomicAdd( &a.counters[0], 1 );
      if ( idx < a.cell_cap ) {
        a.cell_addr[idx] = (uint64_t)( (size_t)f * w * w * w + cid );
        atomicExch( &G[cid], idx );
      } else {
        a.counters[1] = 1;          // overflow: host grows the tables and repeats the pass
        atomicExch( &G[cid], -1 );  // leave the grid clean
      }
    }
  }
}

// ---- pass 2 (geometry): every inside point whose own cell is claimed accumulates (:120-134, :980-998) ----
__global__ void k_accumulate_geo( const GridArgs a, int64_t n ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  const short4 p     = a.pos[i];
  const int    f     = frame_of( a.frame_off, a.F, i );
  const int    disth = max( a.g / 2, 1 ), th = grid_th( a, f );
  if ( !inside( p.x, p.y, p.z, disth, th ) ) { return; }
  const int64_t w   = a.wmax;
  const int64_t cid = ( p.x / a.g ) + ( p.y / a.g ) * w + ( p.z / a.g ) * w * w;
  const int     idx = a.grid[(size_t)f * w * w * w + cid];
  if ( idx < 0 ) { return; }
  Cell*          c  = a.cells + idx;
  const uint32_t pp = a.part[i] + 1u;
  atomicAdd( &c->cnt, 1u );
  atomicAdd( &c->s0, (uint32_t)p.x );
  atomicAdd( &c->s1, (uint32_t)p.y );
  atomicAdd( &c->s2, (uint32_t)p.z );
  atomicMax( &c->pmax, pp );
  atomicMax( &c->pinv, 0xFFFFFFFFu - pp );
}

// ---- pass 2 (colour): every point (no margin test, :208-224) whose own cell is claimed accumulates ----
__global__ void k_accumulate_col( const GridArgs a, int64_t n ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  const short4  p = a.pos[i];
  const int     f = frame_of( a.frame_off, a.F, i );
  const int64_t w = a.wmax;
  if ( p.x < 0 || p.y < 0 || p.z < 0 ) { return; }
  const int64_t cid = ( p.x / a.g ) + ( p.y / a.g ) * w + ( p.z / a.g ) * w * w;
  if ( cid >= w * w * w || p.x / a.g >= w || p.y / a.g >= w ) { return; }  // :212 guard
  const int idx = a.grid[(size_t)f * w * w * w + cid];
  if ( idx < 0 ) { return; }
  Cell*          c  = a.cells + idx;
  const ushort4  cv = a.col[i];
  const uint32_t pp = a.part[i] + 1u;
  atomicAdd( &c->cnt, 1u );
  atomicAdd( &c->s0, (uint32_t)cv.x );
  atomicAdd( &c->s1, (uint32_t)cv.y );
  atomicAdd( &c->s2, (uint32_t)cv.z );
  atomicMax( &c->pmax, pp );
  atomicMax( &c->pinv, 0xFFFFFFFFu - pp );
}

// exactness guard of App. A.3 + luma list allocation (colour)
__global__ void k_finalize_cells( const GridArgs a, int nCells, int colour, int64_t wcube ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= nCells ) { return; }
  Cell* c = a.cells + i;
  if ( c->s0 >= ( 1u << 24 ) || c->s1 >= ( 1u << 24 ) || c->s2 >= ( 1u << 24 ) || c->cnt > 65535u ) {
    const int f = (int)( a.cell_addr[i] / (uint64_t)wcube );
    a.finfo[f].sum_overflow = 1;
  }
  if ( colour ) {
    c->lum_off = (uint32_t)atomicAdd( &a.counters[2], (int)c->cnt );
    c->aux     = 0;
  }
}

// ---- geometry filter: smoothPointCloudGrid + gridFiltering (:1000-1104) ----
__global__ void k_filter_geo( const GridArgs a, int64_t n, double threshold ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  const short4 p = a.pos[i];
  if ( p.w != 1 ) { return; }  // :1087
  const int f     = frame_of( a.frame_off, a.F, i );
  const int g     = a.g, hg = g / 2;
  const int disth = max( hg, 1 ), th = grid_th( a, f );
  if ( !inside( p.x, p.y, p.z, disth, th ) ) { return; }  // :1078-1081
  const int      P[3] = {p.x, p.y, p.z};
  int            S[3];
  for ( int k = 0; k < 3; k++ ) { S[k] = P[k] / g + ( ( P[k] - ( P[k] / g ) * g < hg ) ? -1 : 0 ); }  // :1014-1017
  const int64_t  w = a.wmax;
  const int32_t* G = a.grid + (size_t)f * w * w * w;
  int            idx[8];
  bool           other = false;
  uint32_t       cnt[8];
  for ( int k = 0; k < 8; k++ ) {  // k = dz*4 + dy*2 + dx, the reference's loop order (:1019-1027)
    const int dx = k & 1, dy = ( k >> 1 ) & 1, dz = k >> 2;
    idx[k]       = G[( S[0] + dx ) + ( S[1] + dy ) * w + ( S[2] + dz ) * w * w];
    const Cell* c = a.cells + idx[k];
    cnt[k]        = c->cnt;
    if ( cnt[k] != 0 && c->pmax != 0xFFFFFFFFu - c->pinv ) { other = true; }  // doSmooth && count (:1024)
  }
  if ( !other ) { return; }  // :1028
  const int    g2 = 2 * g;
  int          Wt[3], Q[3];
  for ( int k = 0; k < 3; k++ ) {
    Wt[k] = ( P[k] - S[k] * g - hg ) * 2 + 1;  // :1034
    Q[k]  = g2 - Wt[k];                         // :1047
  }
  double c4[3] = {0.0, 0.0, 0.0};
  int    count = 0;
  for ( int k = 0; k < 8; k++ ) {  // :1050-1058
    const int    dx = k & 1, dy = ( k >> 1 ) & 1, dz = k >> 2;
    const int    wgt = ( dx ? Wt[0] : Q[0] ) * ( dy ? Wt[1] : Q[1] ) * ( dz ? Wt[2] : Q[2] );
    double       v[3];
    if ( cnt[k] > 0 ) {  // :1040: centre = float sum / float count (one IEEE float division, :135-137)
      const Cell* c = a.cells + idx[k];
      const float fc = (float)cnt[k];
      v[0]           = (double)__fdiv_rn( (float)c->s0, fc );
      v[1]           = (double)__fdiv_rn( (float)c->s1, fc );
      v[2]           = (double)__fdiv_rn( (float)c->s2, fc );
    } else {
      v[0] = (double)P[0];
      v[1] = (double)P[1];
      v[2] = (double)P[2];
    }
    const double dw = (double)wgt;
    c4[0] = __dadd_rn( c4[0], __dmul_rn( v[0], dw ) );
    c4[1] = __dadd_rn( c4[1], __dmul_rn( v[1], dw ) );
    c4[2] = __dadd_rn( c4[2], __dmul_rn( v[2], dw ) );
    count += wgt * (int)cnt[k];
  }
  const double den = (double)( g2 * g2 * g2 );
  c4[0]            = c4[0] / den;  // :1059
  c4[1]            = c4[1] / den;
  c4[2]            = c4[2] / den;
  count /= g2 * g2 * g2;  // :1060 integer division; 0 is common (App. A.4)
  if ( count == 0 ) { return; }  // 0.0/0.0 = NaN, NaN >= x is false: the reference leaves the point alone
  const double dc = (double)count;
  double       cen[3], d2 = 0.0;
  {
    const double e0 = __dsub_rn( __dmul_rn( (double)P[0], dc ), ( cen[0] = __dmul_rn( c4[0], dc ) ) );
    const double e1 = __dsub_rn( __dmul_rn( (double)P[1], dc ), ( cen[1] = __dmul_rn( c4[1], dc ) ) );
    const double e2 = __dsub_rn( __dmul_rn( (double)P[2], dc ), ( cen[2] = __dmul_rn( c4[2], dc ) ) );
    d2 = __dadd_rn( __dadd_rn( __dmul_rn( e0, e0 ), __dmul_rn( e1, e1 ) ), __dmul_rn( e2, e2 ) );  // getNorm2
  }
  const double dist2 = __dadd_rn( d2 / dc, 0.5 );  // :1093
  const int    lim   = max( (int)threshold, count ) * 2;
  if ( dist2 >= (double)lim ) {  // :1094
    short4 q = p;
    q.x      = (short)(long long)( __dadd_rn( cen[0] / dc, 0.5 ) );  // :1095-1097
    q.y      = (short)(long long)( __dadd_rn( cen[1] / dc, 0.5 ) );
    q.z      = (short)(long long)( __dadd_rn( cen[2] / dc, 0.5 ) );
    q.w      = 3;  // :1099
    a.pos[i] = q;
    atomicAdd( &a.finfo[f].smoothed, 1 );
  }
}

// ---- colour: scatter lumas into the per-cell lists (colorSmoothingLum_, :1179) ----
__global__ void k_scatter_lum( const GridArgs a, int64_t n, uint16_t* __restrict__ lum ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  const short4  p = a.pos[i];
  const int     f = frame_of( a.frame_off, a.F, i );
  const int64_t w = a.wmax;
  if ( p.x < 0 || p.y < 0 || p.z < 0 ) { return; }
  const int64_t cid = ( p.x / a.g ) + ( p.y / a.g ) * w + ( p.z / a.g ) * w * w;
  if ( cid >= w * w * w || p.x / a.g >= w || p.y / a.g >= w ) { return; }
  const int idx = a.grid[(size_t)f * w * w * w + cid];
  if ( idx < 0 ) { return; }
  Cell* c = a.cells + idx;
  if ( c->cnt < 2 ) { return; }  // the median is only consulted for count > 1 (:1228, :1239)
  const uint32_t slot       = atomicAdd( &c->aux, 1u );
  lum[c->lum_off + slot]    = a.col[i].x;
}

// ---- colour: per-cell mean/median gate (:1228-1236, :1239-1243); one warp per cell, rank selection ----
__global__ void k_cell_median_gate( const GridArgs a, int nCells, const uint16_t* __restrict__ lum, double mmThresh ) {
  const int cell = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
  const int lane = threadIdx.x & 31;
  if ( cell >= nCells ) { return; }
  Cell*     c = a.cells + cell;
  const int n = (int)c->cnt;
  if ( n < 2 ) {
    if ( lane == 0 ) { c->aux = 0; }
    return;
  }
  const uint16_t* L  = lum + c->lum_off;
  const int       hi = n / 2, lo = n / 2 - 1;
  int             vhi = -1, vlo = -1;
  for ( int i = lane; i < n; i += 32 ) {
    const int v    = L[i];
    int       rank = 0;
    for ( int j = 0; j < n; j++ ) {
      const int u = L[j];
      rank += ( u < v ) || ( u == v && j < i );
    }
    if ( rank == hi ) { vhi = v; }
    if ( rank == lo ) { vlo = v; }
  }
#pragma unroll
  for ( int d = 16; d > 0; d >>= 1 ) {
    vhi = max( vhi, __shfl_xor_sync( 0xFFFFFFFFu, vhi, d ) );
    vlo = max( vlo, __shfl_xor_sync( 0xFFFFFFFFu, vlo, d ) );
  }
  if ( lane == 0 ) {
    // median (PCCCodec.h:271-278) and mean (:280-285); abs() on the double difference is int abs(int) (App. A.9)
    const double med  = ( n % 2 == 0 ) ? ( (double)vhi + (double)vlo ) / 2.0 : (double)vhi;
    const double mean = (double)c->s0 / (double)n;
    const int    diff = (int)( mean - med );
    c->aux            = ( (double)( diff < 0 ? -diff : diff ) > mmThresh ) ? 1u : 0u;
  }
}

// ---- colour filter: smoothPointCloudColorLC + gridFilteringColor (:1182-1306) ----
__global__ void k_filter_col( const GridArgs a, int64_t n, double thrSmoothing, double yThresh ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  const short4 p = a.pos[i];
  if ( p.w != 1 ) { return; }  // :1288
  const int g = a.g, hg = g / 2, disth = max( hg, 1 );
  if ( !inside( p.x, p.y, p.z, disth, a.pcmax ) ) { return; }  // :1280-1283
  const int      f = frame_of( a.frame_off, a.F, i );
  const int      P[3] = {p.x, p.y, p.z};
  int            S[3];
  for ( int k = 0; k < 3; k++ ) { S[k] = P[k] / g + ( ( ( P[k] % g ) < hg ) ? -1 : 0 ); }  // :1197-1199
  const int64_t  w = a.wmax;
  const int32_t* G = a.grid + (size_t)f * w * w * w;
  int            idx[8];
  bool           other = false;
  for ( int k = 0; k < 8; k++ ) {
    const int dx = k & 1, dy = ( k >> 1 ) & 1, dz = k >> 2;
    idx[k]       = G[( S[0] + dx ) + ( S[1] + dy ) * w + ( S[2] + dz ) * w * w];
    const Cell* c = a.cells + idx[k];
    if ( c->cnt != 0 && c->pmax != 0xFFFFFFFFu - c->pinv ) { other = true; }  // :1204
  }
  if ( !other ) { return; }  // :1210
  const ushort4 cv     = a.col[i];
  const double  cur[3] = {(double)cv.x, (double)cv.y, (double)cv.z};
  int           Wt[3], Q[3];
  const int     g2 = 2 * g;
  for ( int k = 0; k < 3; k++ ) {
    Wt[k] = ( P[k] - S[k] * g - hg ) * 2 + 1;  // :1212
    Q[k]  = g2 - Wt[k];                         // :1252
  }
  double c3[8][3];
  double Y0       = 0.0;
  bool   keep_own = false;
  for ( int k = 0; k < 8; k++ ) {  // :1218-1251, loop order dz, dy, dx
    const Cell* c = a.cells + idx[k];
    double*     d = c3[k];
    if ( c->cnt > 0 ) {
      const double dn = (double)c->cnt;
      d[0]            = (double)(float)c->s0 / dn;  // :1225 (float accumulator read back as double)
      d[1]            = (double)(float)c->s1 / dn;
      d[2]            = (double)(float)c->s2 / dn;
      if ( k == 0 ) {
        if ( c->cnt > 1 && c->aux ) {  // :1228-1235: result = own colour
          keep_own = true;
          break;
        }
      } else {
        const int dy0 = (int)( Y0 - d[0] );  // abs() truncates (App. A.9), :1238
        bool      own = (double)( dy0 < 0 ? -dy0 : dy0 ) > yThresh;
        if ( c->cnt > 1 && c->aux ) { own = true; }  // :1239-1243
        if ( own ) {
          d[0] = cur[0];
          d[1] = cur[1];
          d[2] = cur[2];
        }
      }
    } else {
      d[0] = cur[0];
      d[1] = cur[1];
      d[2] = cur[2];
    }
    if ( k == 0 ) { Y0 = d[0]; }  // :1248
  }
  if ( keep_own ) { return; }  // centroid = own colour -> |dY| = 0 < threshold unless threshold <= 0
  double c4[3] = {0.0, 0.0, 0.0};
  for ( int k = 0; k < 8; k++ ) {  // :1254-1261
    const int    dx = k & 1, dy = ( k >> 1 ) & 1, dz = k >> 2;
    const double dw = (double)( ( dx ? Wt[0] : Q[0] ) * ( dy ? Wt[1] : Q[1] ) * ( dz ? Wt[2] : Q[2] ) );
    c4[0]           = __dadd_rn( c4[0], __dmul_rn( c3[k][0], dw ) );
    c4[1]           = __dadd_rn( c4[1], __dmul_rn( c3[k][1], dw ) );
    c4[2]           = __dadd_rn( c4[2], __dmul_rn( c3[k][2], dw ) );
  }
  const double den = (double)( g2 * g2 * g2 );
  double       out[3];
  for ( int k = 0; k < 3; k++ ) {
    const double v = __dadd_rn( c4[k] / den, 0.5 ) / 1.0;  // :1262, :1294 with colorCount = 1
    out[k]         = (double)(long long)v;                  // :1295
  }
  const int    dY   = (int)( out[0] - cur[0] );                                // :1299 abs() truncates
  const double dist = (double)( dY < 0 ? -dY : dY ) * 10. / 256.;
  if ( dist >= thrSmoothing ) {                                                 // :1300-1303
    ushort4 q = cv;
    q.x       = (unsigned short)out[0];
    q.y       = (unsigned short)out[1];
    q.z       = (unsigned short)out[2];
    if ( q.x != cv.x || q.y != cv.y || q.z != cv.z ) { atomicAdd( &a.finfo[f].recolored, 1 ); }
    a.col[i] = q;
  }
}

// ---- cleanup: reset exactly the claimed grid cells and accumulators ----
__global__ void k_cleanup_cells( const GridArgs a, int nCells ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= nCells ) { return; }
  a.grid[a.cell_addr[i]] = -1;
  Cell z{};
  a.cells[i] = z;
}

// ---- convertYUV16ToRGB8 (PCCPointSet.h:133-166) / copyRGB16ToRGB8 (:121-127) ----
__global__ void k_to_rgb8( const ushort4* __restrict__ col, uchar4* __restrict__ rgb, int64_t n, int rgb444,
                           int attr_count ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  if ( attr_count == 0 ) {
    rgb[i] = make_uchar4( 127, 127, 127, 0 );  // PCCCodec.cpp:1327-1330
    return;
  }
  const ushort4 c = col[i];
  if ( rgb444 ) {
    rgb[i] = make_uchar4( (unsigned char)c.x, (unsigned char)c.y, (unsigned char)c.z, 0 );
    return;
  }
  const double offset = 32768.0, scale = 65535.0, weight = 1.0 / scale;
  double       y1 = __dmul_rn( weight, (double)c.x );
  double       u1 = __dmul_rn( weight, __dsub_rn( (double)c.y, offset ) );
  double       v1 = __dmul_rn( weight, __dsub_rn( (double)c.z, offset ) );
  y1              = fmin( fmax( y1, 0.0 ), 1.0 );
  u1              = fmin( fmax( u1, -0.5 ), 0.5 );
  v1              = fmin( fmax( v1, -0.5 ), 0.5 );
  double r = __dadd_rn( y1, __dmul_rn( 1.57480, v1 ) );
  double g = __dsub_rn( __dsub_rn( y1, __dmul_rn( 0.18733, u1 ) ), __dmul_rn( 0.46813, v1 ) );
  double b = __dadd_rn( y1, __dmul_rn( 1.85563, u1 ) );
  r        = fmin( fmax( round( __dmul_rn( r, 255.0 ) ), 0.0 ), 255.0 );
  g        = fmin( fmax( round( __dmul_rn( g, 255.0 ) ), 0.0 ), 255.0 );
  b        = fmin( fmax( round( __dmul_rn( b, 255.0 ) ), 0.0 ), 255.0 );
  rgb[i]   = make_uchar4( (unsigned char)r, (unsigned char)g, (unsigned char)b, 0 );
}

__global__ void k_fill_i32( int32_t* p, int64_t n, int32_t v ) {
  for ( int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x ) { p[i] = v; }
}

// shared driver of the mark pass (with table growth)
int run_mark( rb200_ctx* c, GridArgs& a, RbBuf& cellsBuf, RbBuf& addrBuf, int64_t& cellCap, int64_t n, int* nCellsOut,
              const char* name ) {
  for ( int attempt = 0; attempt < 3; attempt++ ) {
    a.cells     = cellsBuf.as<Cell>();
    a.cell_addr = addrBuf.as<uint64_t>();
    a.cell_cap  = cellCap;
    RB_CUDA( cudaMemsetAsync( a.counters, 0, 16, c->stream ) );
    RB_LAUNCH( "mark_cells", k_mark_cells, rb_div_up( n, 256 ), 256, 0, a, n );
    int32_t* h = (int32_t*)rb_pinned( c, 64 );
    if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
    RB_CUDA( cudaMemcpyAsync( h, a.counters, 16, cudaMemcpyDeviceToHost, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
    c->stats.d2h_bytes += 16;
    const int claimed = h[0], overflow = h[1];
    if ( !overflow ) {
      *nCellsOut = claimed;
      return RB200_OK;
    }
    // undo the cells that were claimed, grow, repeat
    const int kept = (int)std::min<int64_t>( claimed, cellCap );
    if ( kept > 0 ) { RB_LAUNCH( "cleanup_cells", k_cleanup_cells, rb_div_up( kept, 256 ), 256, 0, a, kept ); }
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
    const int64_t want = (int64_t)claimed + claimed / 4 + 1024;
    RB_CUDA( cellsBuf.ensure( (size_t)want * sizeof( Cell ) ) );
    RB_CUDA( addrBuf.ensure( (size_t)want * 8 ) );
    RB_CUDA( cudaMemsetAsync( cellsBuf.p, 0, (size_t)want * sizeof( Cell ), c->stream ) );
    cellCap = want;
    (void)name;
  }
  return rb_fail( c, RB200_ERR_NOMEM, "cell table did not converge" );
}

int ensure_grid( rb200_ctx* c, RbBuf& gridBuf, int& curW, int& curF, int w, int F ) {
  const size_t cells = (size_t)F * w * w * w;
  if ( gridBuf.p && curW == w && curF >= F ) { return RB200_OK; }
  cudaError_t e = gridBuf.ensure( cells * 4 );
  if ( e != cudaSuccess ) {
    cudaGetLastError();
    return rb_fail( c, RB200_ERR_NOMEM, "cannot allocate the %d^3 x %d-frame cell grid (%zu MB)", w, F, cells * 4 >> 20 );
  }
  RB_LAUNCH( "grid_init", k_fill_i32, 148 * 8, 256, 0, gridBuf.as<int32_t>(), (int64_t)( gridBuf.cap / 4 ), -1 );
  curW = w;
  curF = F;
  return RB200_OK;
}

}  // namespace

int rb_smooth_geometry_impl( rb200_ctx* c ) {
  const rb200_params& P = c->P;
  const int64_t       n = c->h_frame_off[c->F];
  if ( n == 0 || !P.flag_geometry_smoothing || !P.grid_smoothing ) { return RB200_OK; }  // :64-66
  const int g = P.grid_size;
  if ( g < 1 || g > 64 ) { return rb_fail( c, RB200_ERR_INVALID, "grid_size %d out of range", g ); }
  const int pcmax = 1 << P.geometry_bitdepth_3d;
  const int wmax  = ( pcmax + g - 1 ) / g + 1;
  if ( P.attribute_count > 0 && P.attr_transfer_filter_type == 1 ) {
    // tempFrameBuffer = reconstruct (PCCDecoder.cpp:435): the colour transfer needs the pre-smoothing cloud
    RB_CUDA( c->d_pos_pre.ensure( (size_t)n * 8 ) );
    RB_CUDA( cudaMemcpyAsync( c->d_pos_pre.p, c->d_pos.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream ) );
  }
  int r = ensure_grid( c, c->d_geo_grid, c->geo_grid_w, c->geo_grid_frames, wmax, c->F );
  if ( r ) { return r; }
  RB_CUDA( c->d_scratch[2].ensure( 64 ) );
  if ( c->geo_cell_cap == 0 ) {
    c->geo_cell_cap = 1 << 16;
    RB_CUDA( c->d_geo_cells.ensure( (size_t)c->geo_cell_cap * sizeof( Cell ) ) );
    RB_CUDA( c->d_geo_cell_ids.ensure( (size_t)c->geo_cell_cap * 8 ) );
    RB_CUDA( cudaMemsetAsync( c->d_geo_cells.p, 0, (size_t)c->geo_cell_cap * sizeof( Cell ), c->stream ) );
  }
  GridArgs a{};
  a.F         = c->F;
  a.g         = g;
  a.wmax      = wmax;
  a.by_bbox   = 1;
  a.pcmax     = pcmax;
  a.grid      = c->d_geo_grid.as<int32_t>();
  a.counters  = c->d_scratch[2].as<int32_t>();
  a.frame_off = c->d_frame_off.as<int64_t>();
  a.finfo     = c->d_frame_info.as<RbFrameInfo>();
  a.pos       = c->d_pos.as<short4>();
  a.col       = c->d_col.as<ushort4>();
  a.part      = c->d_part.as<uint32_t>();
  int nCells  = 0;
  r           = run_mark( c, a, c->d_geo_cells, c->d_geo_cell_ids, c->geo_cell_cap, n, &nCells, "geo" );
  if ( r ) { return r; }
  if ( nCells == 0 ) { return RB200_OK; }
  const int G = rb_div_up( n, 256 );
  RB_LAUNCH( "geo_accumulate", k_accumulate_geo, G, 256, 0, a, n );
  RB_LAUNCH( "geo_finalize", k_finalize_cells, rb_div_up( nCells, 256 ), 256, 0, a, nCells, 0, (int64_t)wmax * wmax * wmax );
  RB_LAUNCH( "geo_filter", k_filter_geo, G, 256, 0, a, n, P.threshold_smoothing );
  RB_LAUNCH( "geo_cleanup", k_cleanup_cells, rb_div_up( nCells, 256 ), 256, 0, a, nCells );
  return RB200_OK;
}

int rb_smooth_color_impl( rb200_ctx* c ) {
  const rb200_params& P = c->P;
  const int64_t       n = c->h_frame_off[c->F];
  if ( n == 0 || P.attribute_count == 0 ) { return RB200_OK; }
  const int g     = P.occupancy_precision;  // the colour grid uses occupancyPrecision, not cgridSize (:152)
  const int pcmax = 1 << P.geometry_bitdepth_3d;
  const int wmax  = pcmax / g;  // :154
  if ( wmax < 2 ) { return rb_fail( c, RB200_ERR_INVALID, "colour grid degenerate" ); }
  int r = ensure_grid( c, c->d_col_grid, c->col_grid_w, c->col_grid_frames, wmax, c->F );
  if ( r ) { return r; }
  RB_CUDA( c->d_scratch[3].ensure( 64 ) );
  if ( c->col_cell_cap == 0 ) {
    c->col_cell_cap = 1 << 16;
    RB_CUDA( c->d_col_cells.ensure( (size_t)c->col_cell_cap * sizeof( Cell ) ) );
    RB_CUDA( c->d_col_cell_ids.ensure( (size_t)c->col_cell_cap * 8 ) );
    RB_CUDA( cudaMemsetAsync( c->d_col_cells.p, 0, (size_t)c->col_cell_cap * sizeof( Cell ), c->stream ) );
  }
  GridArgs a{};
  a.F         = c->F;
  a.g         = g;
  a.wmax      = wmax;
  a.by_bbox   = 0;
  a.pcmax     = pcmax;
  a.grid      = c->d_col_grid.as<int32_t>();
  a.counters  = c->d_scratch[3].as<int32_t>();
  a.frame_off = c->d_frame_off.as<int64_t>();
  a.finfo     = c->d_frame_info.as<RbFrameInfo>();
  a.pos       = c->d_pos.as<short4>();
  a.col       = c->d_col.as<ushort4>();
  a.part      = c->d_part.as<uint32_t>();
  int nCells  = 0;
  r           = run_mark( c, a, c->d_col_cells, c->d_col_cell_ids, c->col_cell_cap, n, &nCells, "col" );
  if ( r ) { return r; }
  if ( nCells == 0 ) { return RB200_OK; }
  const int G = rb_div_up( n, 256 );
  RB_CUDA( c->d_col_lum.ensure( (size_t)n * 2 + 64 ) );
  RB_LAUNCH( "col_accumulate", k_accumulate_col, G, 256, 0, a, n );
  RB_LAUNCH( "col_finalize", k_finalize_cells, rb_div_up( nCells, 256 ), 256, 0, a, nCells, 1, (int64_t)wmax * wmax * wmax );
  RB_LAUNCH( "col_scatter_lum", k_scatter_lum, G, 256, 0, a, n, c->d_col_lum.as<uint16_t>() );
  RB_LAUNCH( "col_median_gate", k_cell_median_gate, rb_div_up( (int64_t)nCells * 32, 256 ), 256, 0, a, nCells,
             c->d_col_lum.as<uint16_t>(), P.threshold_color_variation * 256.0 );
  RB_LAUNCH( "col_filter", k_filter_col, G, 256, 0, a, n, P.threshold_color_smoothing, P.threshold_color_difference * 256.0 );
  RB_LAUNCH( "col_cleanup", k_cleanup_cells, rb_div_up( nCells, 256 ), 256, 0, a, nCells );
  return RB200_OK;
}

int rb_convert_rgb8_impl( rb200_ctx* c ) {
  const int64_t n = c->h_frame_off[c->F];
  if ( n == 0 ) { return RB200_OK; }
  RB_LAUNCH( "to_rgb8", k_to_rgb8, rb_div_up( n, 256 ), 256, 0, c->d_col.as<ushort4>(), c->d_rgb.as<uchar4>(), n,
             c->P.attribute_rgb444, c->P.attribute_count );
  return RB200_OK;
}
