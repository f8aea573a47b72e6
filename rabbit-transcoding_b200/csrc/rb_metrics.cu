// rb_metrics.cu — PccLibMetrics on the GPU: duplicate removal, exact nearest-neighbour tie sets, D1 (point-to-point),
// D2 (point-to-plane) and colour PSNR for a batch of (source, reconstruction) pairs (sm_100a).
//
// Restates
//   PCCPointSet3::removeDuplicate( out, dropDuplicates )   PccLibCommon/source/PCCPointSet.cpp:169-218
//   PCCPointSet3::copyNormals / scaleNormals               :2282-2380
//   QualityMetrics::compute / getPSNR / convertRGBtoYUVBT709 / operator+   PccLibMetrics/source/PCCMetrics.cpp:44-231,299-332
//   PCCMetrics::compute( sources, reconstructs, normals )  :334-385
//   PCCKdTree::search (nanoflann knnSearch)                PccLibCommon/source/PCCKdTree.cpp:61-66
//
// Design.  The reference builds a nanoflann kd-tree per cloud and asks for k = 5, 10, ... 30 neighbours until the
// nearest-distance shell is complete; only that shell (the "tie set") is used.  On an integer lattice the shell is
// found exactly with a 2-D column hash instead of a tree:
//  * every cloud becomes a CSR over its (x, y) columns: a dense Dim x Dim table of segment starts (4 MB at vox10)
//    and, per column, the sorted list of z values.  Column order + sorted z IS the lexicographic (x, y, z) order the
//    reference's nested std::map produces, so duplicate removal, the output order of removeDuplicate() and the
//    point indices the reference sorts tie sets by (PCCMetrics.cpp:110) all fall out of the same structure — no
//    radix sort, no tree;
//  * a query walks Chebyshev rings of columns (1, 8, 16, ... columns) and stops as soon as the best squared
//    distance is below (R+1)^2; ties at the best distance are collected on the way (up to the reference's 30).
//    One thread per query handles rings 0..2 (>99 % of the points of a decoded cloud); the rest is queued for a
//    warp-per-query kernel whose lanes split the columns of each ring (per-warp candidate reduction);
//  * d^2 sums are exact uint64 atomics; the double sums (point-to-plane, colour) are reduced per CTA in a fixed
//    order and summed by one CTA per direction, so results do not depend on scheduling (the rare far queries add
//    through double atomics);
//  * MSE -> PSNR is done on the host with the same float libm calls as the reference (log10f), SURVEY App. A.10.
#include <math.h>

#include <algorithm>

#include "rb_common.cuh"
#include "rb_kdtree.cuh"
#include "rb_kdtree_build.cuh"

namespace {

constexpr int MAX_TIES  = 30;  // num_results_max, PCCMetrics.cpp:88
constexpr int NEAR_RING = 2;   // rings handled by the thread-per-query kernel
constexpr int TPB       = 256;

// ------------------------------------------------------------------------------------------------
// batch description
// ------------------------------------------------------------------------------------------------
struct Batch {
  int             nClouds;
  const int64_t*  off;     // [nClouds + 1] first point of every cloud in the concatenated arrays
  int             dim;     // table is dim x dim columns per cloud (+1 lead entry): stride = dim * dim + 1
  int             ox, oy;  // coordinate origin of the table
  int             oz;      // smallest z of the batch (origin of the kd forest of neighborsProc 0)
  int64_t         stride;
  uint32_t*       tab;     // [nClouds * stride] CSR starts (see build)
  const short4*   in_pos;  // [N] x, y, z
  const uchar4*   in_col;  // [N]
  uint64_t*       key_a;   // [N] scattered (z, index) keys
  uint64_t*       key_b;   // [N] keys sorted inside their column
  uint32_t*       first;   // [N + 1] 1 where the sorted slot starts a new position; scanned in place
  short4*         u_pos;   // unique points, cloud c at [off[c], off[c] + ucount[c])
  int16_t*        u_z;
  uchar4*         u_col;
  float4*         u_yuv;   // convertRGBtoYUVBT709 of u_col (PCCMetrics.cpp:50-55), once per unique point
  uint32_t*       u_orig;  // index (inside the cloud) of the first original point at that position
  uint32_t*       ucount;  // [nClouds]
  int             drop;    // dropDuplicates_: 0 keep every point, 1 first colour, 2 mean colour
};

__device__ __forceinline__ void rgb_to_yuv709( int r, int g, int b, float yuv[3] ) {  // PCCMetrics.cpp:50-55
  yuv[0] = (float)( ( 0.2126 * r + 0.7152 * g + 0.0722 * b ) / 255.0 );
  yuv[1] = (float)( ( -0.1146 * r - 0.3854 * g + 0.5000 * b ) / 255.0 + 0.5000 );
  yuv[2] = (float)( ( 0.5000 * r - 0.4542 * g - 0.0458 * b ) / 255.0 + 0.5000 );
}

__device__ __forceinline__ int cloud_of( const int64_t* __restrict__ off, int n, int64_t i ) {
  int lo = 0, hi = n - 1;
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

// import of all clouds of a batch in one launch (grid.y = cloud): host clouds arrive as packed int16 x 3 / uint8 x 3
// rows, resident reconstructions as the context's short4 / uchar4 arrays
struct ImportDesc {
  const void* pos;  // int16 [n][3]  or  short4 [n] (resident)
  const void* col;  // uint8 [n][3]  or  uchar4 [n] (resident); may be null for a host cloud
  int64_t     n, dst;
  int32_t     resident, pad;
};
__global__ void __launch_bounds__( 256 ) k_import_clouds( const ImportDesc* __restrict__ descs, short4* __restrict__ out_pos,
                                                          uchar4* __restrict__ out_col ) {
  const ImportDesc d = descs[blockIdx.y];
  short4*          op = out_pos + d.dst;
  uchar4*          oc = out_col + d.dst;
  if ( d.resident ) {
    const short4* pos = (const short4*)d.pos;
    const uchar4* rgb = (const uchar4*)d.col;
    for ( int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x ) {
      short4 p = pos[i];
      p.w      = 0;
      op[i]    = p;
      oc[i]    = rgb[i];
    }
    return;
  }
  // four points per thread: 24 bytes of positions as three 8-byte loads, 12 bytes of colours as three 4-byte loads
  const int64_t  groups = d.n / 4;
  const uint2*   p8     = (const uint2*)d.pos;
  const uint32_t* c4    = (const uint32_t*)d.col;
  for ( int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x ) {
    const uint2 a = p8[3 * g], b = p8[3 * g + 1], c = p8[3 * g + 2];  // x0 y0 | z0 x1 ; y1 z1 | x2 y2 ; z2 x3 | y3 z3
    uint2*      o = reinterpret_cast<uint2*>( op + 4 * g );  // (the clouds start at arbitrary points: 8-byte stores)
    o[0] = make_uint2( a.x, a.y & 0xFFFFu );
    o[1] = make_uint2( ( a.y >> 16 ) | ( b.x << 16 ), b.x >> 16 );
    o[2] = make_uint2( b.y, c.x & 0xFFFFu );
    o[3] = make_uint2( ( c.x >> 16 ) | ( c.y << 16 ), c.y >> 16 );
    uint4 q = make_uint4( 0, 0, 0, 0 );
    if ( c4 ) {
      const uint32_t u = c4[3 * g], v = c4[3 * g + 1], w = c4[3 * g + 2];  // r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
      q = make_uint4( u & 0xFFFFFFu, ( u >> 24 ) | ( ( v & 0xFFFFu ) << 8 ), ( v >> 16 ) | ( ( w & 0xFFu ) << 16 ), w >> 8 );
    }
    uint32_t* oq = reinterpret_cast<uint32_t*>( oc + 4 * g );
    oq[0] = q.x, oq[1] = q.y, oq[2] = q.z, oq[3] = q.w;
  }
  if ( blockIdx.x == 0 && threadIdx.x < d.n - 4 * groups ) {  // up to three trailing points
    const int64_t  i   = 4 * groups + threadIdx.x;
    const int16_t* pos = (const int16_t*)d.pos;
    const uint8_t* col = (const uint8_t*)d.col;
    op[i] = make_short4( pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2], 0 );
    oc[i] = col ? make_uchar4( col[i * 3], col[i * 3 + 1], col[i * 3 + 2], 0 ) : make_uchar4( 0, 0, 0, 0 );
  }
}

// bounding box of every coordinate of every cloud: bb = { minx, miny, minz, maxx, maxy, maxz }
__global__ void k_bbox( const short4* __restrict__ pos, int64_t n, int* __restrict__ bb ) {
  int mn[3] = {32767, 32767, 32767}, mx[3] = {-32768, -32768, -32768};
  for ( int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x ) {
    const short4 p = pos[i];
    mn[0] = min( mn[0], (int)p.x ), mn[1] = min( mn[1], (int)p.y ), mn[2] = min( mn[2], (int)p.z );
    mx[0] = max( mx[0], (int)p.x ), mx[1] = max( mx[1], (int)p.y ), mx[2] = max( mx[2], (int)p.z );
  }
#pragma unroll
  for ( int k = 0; k < 3; k++ ) {
#pragma unroll
    for ( int d = 16; d > 0; d >>= 1 ) {
      mn[k] = min( mn[k], __shfl_xor_sync( 0xFFFFFFFFu, mn[k], d ) );
      mx[k] = max( mx[k], __shfl_xor_sync( 0xFFFFFFFFu, mx[k], d ) );
    }
    if ( ( threadIdx.x & 31 ) == 0 ) {
      atomicMin( &bb[k], mn[k] );
      atomicMax( &bb[3 + k], mx[k] );
    }
  }
}

__device__ __forceinline__ int64_t column_of( const Batch& b, int c, int x, int y ) {
  return (int64_t)c * b.stride + 1 + (int64_t)( x - b.ox ) * b.dim + ( y - b.oy );
}

// CSR build 1/4: points per column
// (the CSR kernels take the first point / slot they work on: with cached source clouds only the reconstructions'
// part of the batch — the points, slots and tables behind the sources' — is rebuilt)
__global__ void k_column_count( const Batch b, int64_t i0, int64_t n ) {
  const int64_t i = i0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  const int    c = cloud_of( b.off, b.nClouds, i );
  const short4 p = b.in_pos[i];
  atomicAdd( &b.tab[column_of( b, c, p.x, p.y )], 1u );
}
// CSR build 2/4 (after the exclusive scan): scatter (z, index) keys; tab[col] ends up as the END of the column, so
// column col of cloud c is [tab[col - 1], tab[col]) with the per-cloud lead entry closing the first column
__global__ void k_column_scatter( const Batch b, int64_t i0, int64_t n ) {
  const int64_t i = i0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  const int      c    = cloud_of( b.off, b.nClouds, i );
  const short4   p    = b.in_pos[i];
  const uint32_t slot = atomicAdd( &b.tab[column_of( b, c, p.x, p.y )], 1u );
  b.key_a[slot]       = ( (uint64_t)(uint16_t)( (int)p.z + 32768 ) << 32 ) | (uint32_t)( i - b.off[c] );
}
// CSR build 3/4: order every column by (z, original index) by rank counting (columns are short; a long column costs
// O(L) per element but stays parallel over its elements) and flag the first point of every distinct position
__global__ void k_column_rank( const Batch b, int64_t i0, int64_t n ) {
  const int64_t s = i0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( s >= n ) { return; }
  const int      c   = cloud_of( b.off, b.nClouds, s );
  const uint64_t key = b.key_a[s];
  const short4   p   = b.in_pos[b.off[c] + (uint32_t)key];
  const int64_t  col = column_of( b, c, p.x, p.y );
  const uint32_t beg = b.tab[col - 1], end = b.tab[col];
  uint32_t       rank = 0;
  bool           dup  = false;
  for ( uint32_t t = beg; t < end; t++ ) {
    const uint64_t k = b.key_a[t];
    rank += k < key;
    dup |= ( k < key ) && ( ( k >> 32 ) == ( key >> 32 ) );
  }
  b.key_b[beg + rank] = key;
  b.first[beg + rank] = ( b.drop == 0 || !dup ) ? 1u : 0u;
}
// CSR build 4/4 (after the scan of `first`): write the unique points (position, merged colour, first original index)
__global__ void k_column_compact( const Batch b, int64_t i0, int64_t n ) {
  const int64_t s = i0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( s >= n ) { return; }
  const uint32_t u0 = b.first[s], u1 = b.first[s + 1];
  if ( u1 == u0 ) { return; }  // not the first of its position
  const int      c    = cloud_of( b.off, b.nClouds, s );
  const int64_t  base = b.off[c];
  const uint64_t key  = b.key_b[s];
  const int64_t  g0   = base + (uint32_t)key;
  const short4   p    = b.in_pos[g0];
  const int64_t  U    = base + ( u0 - b.first[base] );
  uchar4         cv   = b.in_col[g0];
  if ( b.drop == 2 ) {  // integer mean of the duplicates' colours, PCCPointSet.cpp:188-201
    const uint32_t end = b.tab[column_of( b, c, p.x, p.y )];
    uint32_t       r = cv.x, g = cv.y, bl = cv.z, m = 1;
    for ( int64_t t = s + 1; t < end && ( b.key_b[t] >> 32 ) == ( key >> 32 ); t++ ) {
      const uchar4 q = b.in_col[base + (uint32_t)b.key_b[t]];
      r += q.x, g += q.y, bl += q.z, m++;
    }
    if ( m > 1 ) { cv = make_uchar4( (unsigned char)( r / m ), (unsigned char)( g / m ), (unsigned char)( bl / m ), 0 ); }
  }
  b.u_pos[U]  = p;
  b.u_z[U]    = p.z;
  b.u_col[U]  = cv;
  {
    float y[3];
    rgb_to_yuv709( cv.x, cv.y, cv.z, y );
    b.u_yuv[U] = make_float4( y[0], y[1], y[2], 0.f );
  }
  b.u_orig[U] = (uint32_t)key;
}
// the CSR of the sorted slots becomes the CSR of the unique points; per-cloud unique counts
__global__ void k_table_unique( const Batch b, int firstCloud ) {
  const int64_t total = (int64_t)b.nClouds * b.stride;
  for ( int64_t i = (int64_t)firstCloud * b.stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
        i += (int64_t)gridDim.x * blockDim.x ) {
    const int     c    = (int)( i / b.stride );
    const int64_t base = b.off[c];
    b.tab[i]           = (uint32_t)( base + ( b.first[b.tab[i]] - b.first[base] ) );
    if ( i == (int64_t)c * b.stride ) { b.ucount[c] = b.first[b.off[c + 1]] - b.first[base]; }
  }
}

// ------------------------------------------------------------------------------------------------
// exact nearest-neighbour tie sets on the column CSR
// ------------------------------------------------------------------------------------------------
struct TieSet {
  uint32_t best;  // squared distance
  int      n;
  bool     overflow;
  uint32_t idx[MAX_TIES];  // global unique indices
};

__device__ __forceinline__ void tie_add( TieSet& t, uint32_t d2, uint32_t k ) {
  if ( d2 < t.best ) {
    t.best = d2;
    t.n    = 0;
  }
  if ( t.n < MAX_TIES ) {
    t.idx[t.n++] = k;
  } else {
    t.overflow = true;
  }
}

// all points of one column of cloud cB at squared distance <= best
__device__ __forceinline__ void probe_column( const Batch& b, int cB, int cx, int cy, int z, uint32_t r2, TieSet& t ) {
  if ( (unsigned)( cx - b.ox ) >= (unsigned)b.dim || (unsigned)( cy - b.oy ) >= (unsigned)b.dim ) { return; }
  const int64_t  col = column_of( b, cB, cx, cy );
  const uint32_t beg = b.tab[col - 1], end = b.tab[col];
  if ( end - beg <= 8 ) {
    for ( uint32_t k = beg; k < end; k++ ) {
      const int      dz = z - (int)b.u_z[k];
      const uint32_t d2 = r2 + (uint32_t)( dz * dz );
      if ( d2 <= t.best ) { tie_add( t, d2, k ); }
    }
    return;
  }
  uint32_t lo = beg, hi = end;  // first entry with z' >= z
  while ( lo < hi ) {
    const uint32_t mid = ( lo + hi ) >> 1;
    if ( (int)b.u_z[mid] < z ) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  for ( uint32_t k = lo; k < end; k++ ) {
    const int      dz = (int)b.u_z[k] - z;
    const uint32_t d2 = r2 + (uint32_t)( dz * dz );
    if ( d2 > t.best ) { break; }
    tie_add( t, d2, k );
  }
  for ( uint32_t k = lo; k > beg; k-- ) {
    const int      dz = z - (int)b.u_z[k - 1];
    const uint32_t d2 = r2 + (uint32_t)( dz * dz );
    if ( d2 > t.best ) { break; }
    tie_add( t, d2, k - 1 );
  }
}

// column (dx, dy) number j of Chebyshev ring R (8R columns; ring 0 has one)
__device__ __forceinline__ void ring_offset( int R, int j, int& dx, int& dy ) {
  if ( R == 0 ) {
    dx = dy = 0;
    return;
  }
  const int side = 2 * R;
  const int s = j / side, o = j % side;
  switch ( s ) {
    case 0: dx = -R + o; dy = -R; break;
    case 1: dx = R; dy = -R + o; break;
    case 2: dx = R - o; dy = R; break;
    default: dx = -R; dy = R - o; break;
  }
}

// rings 0..NEAR_RING; returns true when the tie set is complete
__device__ __forceinline__ bool search_near( const Batch& b, int cB, int x, int y, int z, TieSet& t ) {
  t.best     = 0xFFFFFFFFu;
  t.n        = 0;
  t.overflow = false;
  for ( int R = 0; R <= NEAR_RING; R++ ) {
    const int cnt = R == 0 ? 1 : 8 * R;
    for ( int j = 0; j < cnt; j++ ) {
      int dx, dy;
      ring_offset( R, j, dx, dy );
      const uint32_t r2 = (uint32_t)( dx * dx + dy * dy );
      if ( r2 <= t.best ) { probe_column( b, cB, x + dx, y + dy, z, r2, t ); }
    }
    if ( t.best < (uint32_t)( ( R + 1 ) * ( R + 1 ) ) ) { return true; }
  }
  return false;
}

// ring 0 alone (the query's own column): complete exactly when the point itself is in B (distance 0 < 1)
__device__ __forceinline__ bool search_ring0( const Batch& b, int cB, int x, int y, int z, TieSet& t ) {
  t.best     = 0xFFFFFFFFu;
  t.n        = 0;
  t.overflow = false;
  probe_column( b, cB, x, y, z, 0u, t );
  return t.best == 0u;
}

__device__ __forceinline__ void sort_ties( TieSet& t ) {  // ascending index = the reference's std::sort, PCCMetrics.cpp:110
  for ( int i = 1; i < t.n; i++ ) {
    const uint32_t v = t.idx[i];
    int            j = i - 1;
    while ( j >= 0 && t.idx[j] > v ) {
      t.idx[j + 1] = t.idx[j];
      j--;
    }
    t.idx[j + 1] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// what a completed tie set is used for
// ------------------------------------------------------------------------------------------------
enum Mode { MODE_SCALE = 0, MODE_FILL = 1, MODE_METRIC = 2 };

struct Direction {  // one A -> B pass
  int32_t cloudA, cloudB;
  int32_t nrmA, nrmB;   // index of the normal arrays of A / B (-1: none)
  int32_t acc;          // accumulator slot
  int32_t block_begin;  // first CTA of this direction in the near kernel
};

struct Acc {
  unsigned long long sse_c2c, max_c2c;  // exact
  unsigned long long max_c2p_bits;      // non-negative doubles order like their bit patterns
  double             far_c2p, far_col[3];
  double             sum_c2p, sum_col[3];  // written by k_reduce
  unsigned int       tie_overflow, far_queries;
};

struct NNArgs {
  Batch            b;
  const Direction* dirs;
  int              nDirs;
  double*          nrm;       // [N][3] per-unique-point normals (copyNormals / scaleNormals results)
  uint32_t*        nrm_cnt;   // [N] scaleNormals' count[]
  Acc*             acc;
  double*          partial;   // [blocks][4]
  uint32_t*        far_list;  // [capacity][2] (direction, unique index)
  uint32_t*        far_count;
  // queries the ring-0 pass could not finish, compacted IN ORDER (so the double sums do not depend on scheduling):
  uint32_t*        pend_mask;  // [warps of the ring-0 grid] ballot of the unfinished lanes
  uint32_t*        pend_off;   // [warps + 1] their counts, then the exclusive prefix
  uint32_t*        pend_list;  // [pending] global thread id of the query in the ring-0 grid
  double*          contrib;    // [pending][4] MODE_METRIC: the query's c2p / colour terms (added by k_reduce_partials)
  uint32_t         nPending;
  int              compute_c2p, compute_color, neighbors_proc;
  const uint32_t*  first_idx;  // neighborsProc 0: per unique point of A, the nearest point of B nanoflann returns first
  int              nPairs;     // clouds [0, nPairs) are the sources, [nPairs, 2 nPairs) the reconstructions
};


struct Contribution {
  double             c2p, col[3];
  unsigned long long d2;        // squared distance to the nearest neighbour (exact)
  double             c2p_max;   // same as c2p (kept separately so the reduction can take a maximum)
  unsigned int       overflow;  // tie set larger than 30
};

template <int MODE>
__device__ __forceinline__ void consume( const NNArgs& a, const Direction& d, int64_t uA, TieSet& t, Contribution& out ) {
  const Batch& b = a.b;
  if ( MODE == MODE_SCALE ) {  // scaleNormals first loop, PCCPointSet.cpp:2337-2355
    const double* nA = a.nrm + 3 * uA;
    for ( int j = 0; j < t.n; j++ ) {
      atomicAdd( &a.nrm[3 * (int64_t)t.idx[j] + 0], nA[0] );
      atomicAdd( &a.nrm[3 * (int64_t)t.idx[j] + 1], nA[1] );
      atomicAdd( &a.nrm[3 * (int64_t)t.idx[j] + 2], nA[2] );
      atomicAdd( &a.nrm_cnt[t.idx[j]], 1u );
    }
    return;
  }
  if ( MODE == MODE_FILL ) {  // scaleNormals second loop (count == 0), :2362-2378
    sort_ties( t );
    double s[3] = {0.0, 0.0, 0.0};
    for ( int j = 0; j < t.n; j++ ) {
      const double* nB = a.nrm + 3 * (int64_t)t.idx[j];
      s[0] += nB[0], s[1] += nB[1], s[2] += nB[2];
    }
    a.nrm[3 * uA + 0] = s[0] / (double)t.n;
    a.nrm[3 * uA + 1] = s[1] / (double)t.n;
    a.nrm[3 * uA + 2] = s[2] / (double)t.n;
    a.nrm_cnt[uA]     = 1u;  // a finished normal: "sum / 1"
    return;
  }
  // QualityMetrics::compute body for one point of A, PCCMetrics.cpp:92-191
  out.d2       = t.best;
  out.overflow = t.overflow ? 1u : 0u;
  sort_ties( t );
  const short4 pA = b.u_pos[uA];
  if ( a.compute_c2p && d.nrmA >= 0 && d.nrmB >= 0 ) {  // :113-124
    double sum = 0.0;
    for ( int j = 0; j < t.n; j++ ) {
      const short4  pB = b.u_pos[t.idx[j]];
      const double* sB = a.nrm + 3 * (int64_t)t.idx[j];
      double        nB[3] = {sB[0], sB[1], sB[2]};
      if ( d.cloudB >= a.nPairs ) {  // B is a reconstruction: scaleNormals' sum / count (PCCPointSet.cpp:2357-2361)
        const double cnt = (double)a.nrm_cnt[t.idx[j]];
        nB[0] /= cnt, nB[1] /= cnt, nB[2] /= cnt;
      }
      const double  e0 = (double)( (int)pA.x - (int)pB.x ), e1 = (double)( (int)pA.y - (int)pB.y ),
                   e2 = (double)( (int)pA.z - (int)pB.z );
      const double dot = __dadd_rn( __dadd_rn( __dmul_rn( e0, nB[0] ), __dmul_rn( e1, nB[1] ) ), __dmul_rn( e2, nB[2] ) );
      sum              = __dadd_rn( sum, __dmul_rn( dot, dot ) );
    }
    const double v = sum / (double)t.n;
    out.c2p        = v;
    out.c2p_max    = v;
  }
  if ( a.compute_color ) {  // :126-178
    const float4 fA = b.u_yuv[uA];  // the conversion of the point's own colour, done once when the cloud was built
    float        yA[3] = {fA.x, fA.y, fA.z}, yB[3];
    int          rB = 0, gB = 0, bB = 0;
    bool         haveB = false;
    if ( a.neighbors_proc == 0 ) {  // the colour of result.indices( 0 ), :126, :177: the first nearest point in traversal order
      const float4 fB = b.u_yuv[t.n == 1 ? t.idx[0] : a.first_idx[uA]];
      yB[0] = fB.x, yB[1] = fB.y, yB[2] = fB.z;
      haveB = true;
    } else if ( ( a.neighbors_proc == 1 || a.neighbors_proc == 2 ) && t.n == 1 ) {  // mean of one colour = that colour
      const float4 fB = b.u_yuv[t.idx[0]];
      yB[0] = fB.x, yB[1] = fB.y, yB[2] = fB.z;
      haveB = true;
    } else if ( a.neighbors_proc == 1 || a.neighbors_proc == 2 ) {  // mean colour of the tie set, :137-156
      unsigned int r = 0, g = 0, bl = 0;
      for ( int j = 0; j < t.n; j++ ) {
        const uchar4 q = b.u_col[t.idx[j]];
        r += q.x, g += q.y, bl += q.z;
      }
      rB = (unsigned char)round( (double)r / t.n );
      gB = (unsigned char)round( (double)g / t.n );
      bB = (unsigned char)round( (double)bl / t.n );
    } else if ( a.neighbors_proc == 4 ) {  // the tie-set member with the largest colour distance, :157-176
      float    distBest = 0.f;
      uint32_t best     = (uint32_t)b.off[d.cloudB];  // indexBest starts at point 0 of B
      for ( int j = 0; j < t.n; j++ ) {
        const uchar4 q = b.u_col[t.idx[j]];
        rgb_to_yuv709( q.x, q.y, q.z, yB );
        const float d0 = __fsub_rn( yA[0], yB[0] ), d1 = __fsub_rn( yA[1], yB[1] ), d2 = __fsub_rn( yA[2], yB[2] );
        const float dist = __fadd_rn( __fadd_rn( __fmul_rn( d0, d0 ), __fmul_rn( d1, d1 ) ), __fmul_rn( d2, d2 ) );
        if ( dist > distBest ) {
          distBest = dist;
          best     = t.idx[j];
        }
      }
      const uchar4 q = b.u_col[best];
      rB = q.x, gB = q.y, bB = q.z;
    } else {  // neighborsProc 3 ("min"): dist < distBest (= 0) never holds, so the reference ends up with point 0 of B
      const uchar4 q = b.u_col[b.off[d.cloudB]];
      rB = q.x, gB = q.y, bB = q.z;
    }
    if ( !haveB ) { rgb_to_yuv709( rB, gB, bB, yB ); }
#pragma unroll
    for ( int k = 0; k < 3; k++ ) {
      const float df = __fsub_rn( yA[k], yB[k] );
      out.col[k]     = (double)__fmul_rn( df, df );  // pow( float, 2.F ) is powf: the correctly rounded square
    }
  }
}

__device__ __forceinline__ const Direction& direction_of_block( const NNArgs& a, int blk ) {
  int lo = 0, hi = a.nDirs - 1;
  while ( lo < hi ) {
    const int mid = ( lo + hi + 1 ) >> 1;
    if ( a.dirs[mid].block_begin <= blk ) {
      lo = mid;
    } else {
      hi = mid - 1;
    }
  }
  return a.dirs[lo];
}

// the fixed-order reduction of one warp's METRIC contributions: exact integer sums / maxima through one atomic per warp,
// the double terms as per-warp partials (k_reduce_partials adds them in warp order)
__device__ __forceinline__ void warp_reduce_metric( const NNArgs& a, const Direction& d, const Contribution& ct, int64_t partialSlot ) {
  double    v[4] = {ct.c2p, ct.col[0], ct.col[1], ct.col[2]};
  const int lane = threadIdx.x & 31;
  unsigned long long s2 = ct.d2, m2 = ct.d2, mp = (unsigned long long)__double_as_longlong( ct.c2p_max );
  unsigned int       ov = ct.overflow;
#pragma unroll
  for ( int s = 16; s > 0; s >>= 1 ) {
    s2 += __shfl_down_sync( 0xFFFFFFFFu, s2, s );
    m2 = max( m2, __shfl_down_sync( 0xFFFFFFFFu, m2, s ) );
    mp = max( mp, __shfl_down_sync( 0xFFFFFFFFu, mp, s ) );
    ov += __shfl_down_sync( 0xFFFFFFFFu, ov, s );
  }
  if ( lane == 0 ) {
    Acc* acc = a.acc + d.acc;
    if ( s2 ) { atomicAdd( &acc->sse_c2c, s2 ); }
    if ( m2 ) { atomicMax( &acc->max_c2c, m2 ); }
    if ( mp ) { atomicMax( &acc->max_c2p_bits, mp ); }
    if ( ov ) { atomicAdd( &acc->tie_overflow, ov ); }
  }
#pragma unroll
  for ( int k = 0; k < 4; k++ ) {
#pragma unroll
    for ( int s = 16; s > 0; s >>= 1 ) { v[k] += __shfl_down_sync( 0xFFFFFFFFu, v[k], s ); }
    // per-warp partials go to global memory: no barrier (a CTA's warps finish at very different times), and
    // k_reduce_partials adds them in the same fixed order (warp 0 .. 7 of CTA 0, then CTA 1, ...)
    if ( lane == 0 ) { a.partial[partialSlot * 4 + k] = v[k]; }
  }
}

// pass 1, one thread per unique point of A: the query's own column.  Most points of a decoded cloud ARE in the other
// cloud (distance 0) and are finished here; the others are only marked — walking the 8 + 16 columns of rings 1 and 2
// for a few lanes of every warp is what kept two thirds of the lanes idle.
template <int MODE>
__global__ void __launch_bounds__( TPB ) k_nn_near( const NNArgs a ) {
  const Direction&  d  = direction_of_block( a, blockIdx.x );
  const Batch&      b  = a.b;
  const int64_t     iu = (int64_t)( blockIdx.x - d.block_begin ) * TPB + threadIdx.x;
  Contribution      ct{};
  const bool        active  = iu < (int64_t)b.ucount[d.cloudA];
  bool              pending = false;
  if ( active ) {
    const int64_t uA   = b.off[d.cloudA] + iu;
    bool          need = true;
    if ( MODE == MODE_FILL ) {
      // points that received normals are only normalised (:2357-2361): the sums stay as they are and the METRIC pass
      // divides by the count where it reads them (the same double division, without a pass over 24 bytes per point)
      if ( a.nrm_cnt[uA] > 0 ) { need = false; }
    }
    if ( need && MODE == MODE_FILL ) {
      // a point no source point chose has no source point at its own position (that one would have chosen it): its
      // search starts in the dense pass, where every lane has one
      pending = true;
    } else if ( need ) {
      const short4 p = b.u_pos[uA];
      TieSet       t;
      if ( search_ring0( b, d.cloudB, p.x, p.y, p.z, t ) ) {
        consume<MODE>( a, d, uA, t, ct );
      } else {
        pending = true;
      }
    }
  }
  const int      lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t pm   = __ballot_sync( 0xFFFFFFFFu, pending );
  if ( lane == 0 ) {
    const int64_t gw = (int64_t)blockIdx.x * ( TPB / 32 ) + w;
    a.pend_mask[gw]  = pm;
    a.pend_off[gw]   = __popc( pm );
  }
  if ( MODE == MODE_METRIC ) { warp_reduce_metric( a, d, ct, (int64_t)blockIdx.x * ( TPB / 32 ) + w ); }
}

// neighborsProc 0 (PCCMetrics.cpp:126, :177): the colour comes from result.indices( 0 ), the nearest point nanoflann's
// traversal meets first.  The unique points of every cloud are in the order removeDuplicate leaves them (x, y, z
// ascending), i.e. the order PCCKdTree indexes them in: the emulated trees of rb_kdtree.cu are built over them and a
// 1-NN search (the first point at the smallest distance is the same for every k) gives that index.
__global__ void k_gather_unique( const Batch b, const int64_t* __restrict__ coff, short4* __restrict__ out ) {
  const int     cloud = blockIdx.y;
  const int64_t n     = b.ucount[cloud];
  for ( int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x ) {
    short4 p = b.u_pos[b.off[cloud] + i];
    p.w      = 0;
    out[coff[cloud] + i] = p;
  }
}
__global__ void __launch_bounds__( 128 ) k_nn_first( const NNArgs a, const KdForest forest, uint32_t* __restrict__ first_idx ) {
  // (CTAs of 128 threads, two per CTA of the ring-0 grid the directions are laid out for)
  const int        blk = blockIdx.x / ( TPB / 128 );
  const Direction& d   = direction_of_block( a, blk );
  const Batch&     b   = a.b;
  const int64_t    iu  = (int64_t)( blk - d.block_begin ) * TPB + ( blockIdx.x % ( TPB / 128 ) ) * 128 + threadIdx.x;
  if ( iu >= (int64_t)b.ucount[d.cloudA] ) { return; }
  const int64_t uA   = b.off[d.cloudA] + iu;
  const short4  p    = b.u_pos[uA];
  const int     q[3] = {p.x - forest.ox, p.y - forest.oy, p.z - forest.oz};
  KdResult<1>   res;
  kd_search<1>( forest, (uint32_t)d.cloudB + 1u, q, res );
  first_idx[uA] = (uint32_t)( b.off[d.cloudB] + res.idx[0] );
}

// the unfinished queries as an ordered list (pend_off is the exclusive prefix of the per-warp counts by now): one thread
// per warp of the ring-0 grid
__global__ void __launch_bounds__( TPB ) k_nn_list_pending( const NNArgs a, int64_t nWarps ) {
  const int64_t gw = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if ( gw >= nWarps ) { return; }
  uint32_t o = a.pend_off[gw];
  for ( uint32_t pm = a.pend_mask[gw]; pm; pm &= pm - 1 ) { a.pend_list[o++] = (uint32_t)( gw * 32 + ( __ffs( pm ) - 1 ) ); }
}

// pass 2, one thread per unfinished query: rings 0..2 with every lane busy; what is still open goes to the warp-per-query
// kernel.  The double terms of a METRIC query are stored per list slot and added in slot order by k_reduce_partials.
template <int MODE>
__global__ void __launch_bounds__( TPB ) k_nn_pending( const NNArgs a ) {
  const Batch&   b = a.b;
  const uint32_t s = blockIdx.x * TPB + threadIdx.x;
  Contribution   ct{};
  int            accIdx = -1;
  if ( s < a.nPending ) {
    const uint32_t   gt  = a.pend_list[s];
    const int        blk = (int)( gt / TPB );
    const Direction& d   = direction_of_block( a, blk );
    const int64_t    iu  = (int64_t)( blk - d.block_begin ) * TPB + ( gt % TPB );
    const int64_t    uA  = b.off[d.cloudA] + iu;
    const short4     p   = b.u_pos[uA];
    TieSet           t;
    if ( search_near( b, d.cloudB, p.x, p.y, p.z, t ) ) {
      consume<MODE>( a, d, uA, t, ct );
      accIdx = d.acc;
    } else {
      const uint32_t slot      = atomicAdd( a.far_count, 1u );
      a.far_list[2 * slot]     = (uint32_t)( &d - a.dirs );
      a.far_list[2 * slot + 1] = (uint32_t)iu;
    }
    if ( MODE == MODE_METRIC ) {
      double* o = a.contrib + (size_t)s * 4;
      o[0] = ct.c2p, o[1] = ct.col[0], o[2] = ct.col[1], o[3] = ct.col[2];
    }
  }
  if ( MODE == MODE_METRIC ) {  // exact integer parts: one atomic per group of lanes with the same accumulator
    const uint32_t live = __ballot_sync( 0xFFFFFFFFu, accIdx >= 0 );
    if ( accIdx >= 0 ) {
      const uint32_t peers  = __match_any_sync( live, accIdx );
      const int      leader = __ffs( peers ) - 1;
      unsigned long long s2 = 0, m2 = 0, mp = 0;
      unsigned int       ov = 0;
      for ( uint32_t m = peers; m; m &= m - 1 ) {
        const int src = __ffs( m ) - 1;
        const unsigned long long q2 = __shfl_sync( peers, ct.d2, src );
        const unsigned long long qp = __shfl_sync( peers, (unsigned long long)__double_as_longlong( ct.c2p_max ), src );
        s2 += q2, m2 = max( m2, q2 ), mp = max( mp, qp ), ov += __shfl_sync( peers, ct.overflow, src );
      }
      if ( ( threadIdx.x & 31 ) == leader ) {
        Acc* acc = a.acc + accIdx;
        if ( s2 ) { atomicAdd( &acc->sse_c2c, s2 ); }
        if ( m2 ) { atomicMax( &acc->max_c2c, m2 ); }
        if ( mp ) { atomicMax( &acc->max_c2p_bits, mp ); }
        if ( ov ) { atomicAdd( &acc->tie_overflow, ov ); }
      }
    }
  }
}

// warp per far query: lanes split the columns of each ring; the nearest distance first, then the tie set
template <int MODE>
__global__ void __launch_bounds__( TPB ) k_nn_far( const NNArgs a ) {
  __shared__ uint32_t tieBuf[TPB / 32][MAX_TIES];
  __shared__ int      tieCnt[TPB / 32];
  const int           lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t      nFar = *a.far_count;
  const Batch&        b    = a.b;
  for ( uint32_t q = blockIdx.x * ( TPB / 32 ) + w; q < nFar; q += gridDim.x * ( TPB / 32 ) ) {
    const Direction& d  = a.dirs[a.far_list[2 * q]];
    const int64_t    uA = b.off[d.cloudA] + a.far_list[2 * q + 1];
    const short4     p  = b.u_pos[uA];
    // phase 1: nearest squared distance
    uint32_t best = 0xFFFFFFFFu;
    for ( int R = 0; R < 2 * b.dim; R++ ) {
      const int cnt = R == 0 ? 1 : 8 * R;
      for ( int j = lane; j < cnt; j += 32 ) {
        int dx, dy;
        ring_offset( R, j, dx, dy );
        const uint32_t r2 = (uint32_t)( dx * dx + dy * dy );
        const int      cx = p.x + dx, cy = p.y + dy;
        if ( r2 > best || (unsigned)( cx - b.ox ) >= (unsigned)b.dim || (unsigned)( cy - b.oy ) >= (unsigned)b.dim ) { continue; }
        const int64_t  col = column_of( b, d.cloudB, cx, cy );
        const uint32_t beg = b.tab[col - 1], end = b.tab[col];
        if ( beg == end ) { continue; }
        uint32_t lo = beg, hi = end;
        while ( lo < hi ) {
          const uint32_t mid = ( lo + hi ) >> 1;
          if ( (int)b.u_z[mid] < (int)p.z ) {
            lo = mid + 1;
          } else {
            hi = mid;
          }
        }
        if ( lo < end ) {
          const int dz = (int)b.u_z[lo] - p.z;
          best         = min( best, r2 + (uint32_t)( dz * dz ) );
        }
        if ( lo > beg ) {
          const int dz = p.z - (int)b.u_z[lo - 1];
          best         = min( best, r2 + (uint32_t)( dz * dz ) );
        }
      }
#pragma unroll
      for ( int s = 16; s > 0; s >>= 1 ) { best = min( best, __shfl_xor_sync( 0xFFFFFFFFu, best, s ) ); }
      if ( best < (uint32_t)( ( R + 1 ) * ( R + 1 ) ) ) { break; }
    }
    // phase 2: every point at exactly that distance
    if ( lane == 0 ) { tieCnt[w] = 0; }
    __syncwarp();
    const int Rmax = (int)sqrtf( (float)best ) + 1;
    for ( int R = 0; R <= Rmax && R < 2 * b.dim; R++ ) {
      const int cnt = R == 0 ? 1 : 8 * R;
      for ( int j = lane; j < cnt; j += 32 ) {
        int dx, dy;
        ring_offset( R, j, dx, dy );
        const uint32_t r2 = (uint32_t)( dx * dx + dy * dy );
        const int      cx = p.x + dx, cy = p.y + dy;
        if ( r2 > best || (unsigned)( cx - b.ox ) >= (unsigned)b.dim || (unsigned)( cy - b.oy ) >= (unsigned)b.dim ) { continue; }
        const int64_t  col = column_of( b, d.cloudB, cx, cy );
        const uint32_t beg = b.tab[col - 1], end = b.tab[col];
        for ( uint32_t k = beg; k < end; k++ ) {  // far columns are short; exact test
          const int dz = p.z - (int)b.u_z[k];
          if ( r2 + (uint32_t)( dz * dz ) == best ) {
            const int slot = atomicAdd( &tieCnt[w], 1 );
            if ( slot < MAX_TIES ) { tieBuf[w][slot] = k; }
          }
        }
      }
    }
    __syncwarp();
    if ( lane == 0 ) {
      TieSet t;
      t.best     = best;
      t.overflow = tieCnt[w] > MAX_TIES;
      t.n        = min( tieCnt[w], MAX_TIES );
      for ( int j = 0; j < t.n; j++ ) { t.idx[j] = tieBuf[w][j]; }
      sort_ties( t );
      Contribution ct{};
      consume<MODE>( a, d, uA, t, ct );
      if ( MODE == MODE_METRIC ) {
        Acc* acc = a.acc + d.acc;
        atomicAdd( &acc->sse_c2c, ct.d2 );
        atomicMax( &acc->max_c2c, ct.d2 );
        atomicMax( &acc->max_c2p_bits, (unsigned long long)__double_as_longlong( ct.c2p_max ) );
        if ( ct.overflow ) { atomicAdd( &acc->tie_overflow, 1u ); }
        atomicAdd( &acc->far_c2p, ct.c2p );
        atomicAdd( &acc->far_col[0], ct.col[0] );
        atomicAdd( &acc->far_col[1], ct.col[1] );
        atomicAdd( &acc->far_col[2], ct.col[2] );
        atomicAdd( &acc->far_queries, 1u );
      }
    }
    __syncwarp();
  }
}

// the double terms of a direction are added in a fixed order: RED_SEG CTAs per direction each sum a fixed slice of the
// direction's warp partials and pass-2 terms (k_reduce_segments), one CTA per direction adds the slices in slice order
constexpr int RED_SEG = 16;
__device__ __forceinline__ void cta_sum4( double s[4], double ( *red )[4] ) {
#pragma unroll
  for ( int k = 0; k < 4; k++ ) { red[threadIdx.x][k] = s[k]; }
  __syncthreads();
  for ( int st = TPB / 2; st > 0; st >>= 1 ) {
    if ( threadIdx.x < st ) {
#pragma unroll
      for ( int k = 0; k < 4; k++ ) { red[threadIdx.x][k] += red[threadIdx.x + st][k]; }
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__( TPB ) k_reduce_segments( const NNArgs a, const int32_t* __restrict__ block_end, double* __restrict__ seg ) {
  __shared__ double red[TPB][4];
  const int         dir = blockIdx.x / RED_SEG, sg = blockIdx.x % RED_SEG;
  const Direction&  d   = a.dirs[dir];
  double            s[4] = {0, 0, 0, 0};
  {
    const int64_t nb = block_end[dir] - d.block_begin, per = ( nb + RED_SEG - 1 ) / RED_SEG;
    const int64_t b0 = d.block_begin + sg * per, b1 = min( (int64_t)block_end[dir], b0 + per );
    for ( int64_t blk = b0 + threadIdx.x; blk < b1; blk += TPB ) {
#pragma unroll
      for ( int k = 0; k < 4; k++ ) {
        double bs = 0.0;  // the CTA's sum, warp by warp
        for ( int w = 0; w < TPB / 32; w++ ) { bs += a.partial[( blk * ( TPB / 32 ) + w ) * 4 + k]; }
        s[k] += bs;
      }
    }
    const int64_t pb = a.pend_off[(int64_t)d.block_begin * ( TPB / 32 )], pe = a.pend_off[(int64_t)block_end[dir] * ( TPB / 32 )];
    const int64_t pper = ( pe - pb + RED_SEG - 1 ) / RED_SEG, q0 = pb + sg * pper, q1 = min( pe, q0 + pper );
    for ( int64_t q = q0 + threadIdx.x; q < q1; q += TPB ) {
      const double2 u = *reinterpret_cast<const double2*>( a.contrib + (size_t)q * 4 ), v = *reinterpret_cast<const double2*>( a.contrib + (size_t)q * 4 + 2 );
      s[0] += u.x, s[1] += u.y, s[2] += v.x, s[3] += v.y;
    }
  }
  cta_sum4( s, red );
  if ( threadIdx.x < 4 ) { seg[(size_t)blockIdx.x * 4 + threadIdx.x] = red[0][threadIdx.x]; }
}
__global__ void k_reduce_partials( const NNArgs a, const double* __restrict__ seg ) {
  const int dir = blockIdx.x * blockDim.x + threadIdx.x;
  if ( dir >= a.nDirs ) { return; }
  double s[4] = {0, 0, 0, 0};
  for ( int g = 0; g < RED_SEG; g++ ) {
    for ( int k = 0; k < 4; k++ ) { s[k] += seg[( (size_t)dir * RED_SEG + g ) * 4 + k]; }
  }
  Acc* acc     = a.acc + a.dirs[dir].acc;
  acc->sum_c2p = s[0] + acc->far_c2p;
  for ( int k = 0; k < 3; k++ ) { acc->sum_col[k] = s[1 + k] + acc->far_col[k]; }
}

// copyNormals (PCCPointSet.cpp:2282-2320): exact position lookup of every point of the normal cloud in the source
// cloud's CSR; the LAST normal-cloud index at a position wins (map[x][y][z] = i)
struct NormalDesc {  // one source cloud with normals: its cloud index in the batch, its normals, its point count
  const float* nrm_in;
  int64_t      n;
  int32_t      cloudS, pad;
};
// all source clouds in one launch (grid.y = cloud); the normal cloud's positions are the imported positions of the
// source cloud itself (in_pos, file order)
__global__ void __launch_bounds__( 256 ) k_normal_lookup( const Batch b, const NormalDesc* __restrict__ descs,
                                                          uint32_t* __restrict__ last_idx ) {
  const NormalDesc d   = descs[blockIdx.y];
  const short4*    pos = b.in_pos + b.off[d.cloudS];
  for ( int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x ) {
    const short4 p = pos[i];
    const int    x = p.x, y = p.y, z = p.z;
    if ( (unsigned)( x - b.ox ) < (unsigned)b.dim && (unsigned)( y - b.oy ) < (unsigned)b.dim ) {
      const int64_t  col = column_of( b, d.cloudS, x, y );
      const uint32_t beg = b.tab[col - 1], end = b.tab[col];
      for ( uint32_t k = beg; k < end; k++ ) {
        if ( (int)b.u_z[k] == z ) {
          atomicMax( &last_idx[k], (uint32_t)i + 1u );
          break;
        }
      }
    }
    // a normal-cloud point that is not in the source is harmless to the reference unless a source point stays uncovered
  }
}
__global__ void __launch_bounds__( 256 ) k_normal_gather( const Batch b, const NormalDesc* __restrict__ descs,
                                                          const uint32_t* __restrict__ last_idx, double* __restrict__ nrm,
                                                          uint32_t* __restrict__ err ) {
  const NormalDesc d  = descs[blockIdx.y];
  const int64_t    nU = b.ucount[d.cloudS];
  if ( blockIdx.x == 0 && threadIdx.x == 0 && nU != d.n ) { atomicOr( err, 1u ); }  // "must have the same number of points", :2287-2293
  for ( int64_t iu = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; iu < nU; iu += (int64_t)gridDim.x * blockDim.x ) {
    const int64_t  U  = b.off[d.cloudS] + iu;
    const uint32_t li = last_idx[U];
    if ( li == 0 ) {
      atomicOr( err, 2u );  // "point i of the current points cloud is not present in the normal point cloud", :2311-2318
      continue;
    }
    nrm[3 * U + 0] = (double)d.nrm_in[3 * (int64_t)( li - 1 ) + 0];
    nrm[3 * U + 1] = (double)d.nrm_in[3 * (int64_t)( li - 1 ) + 1];
    nrm[3 * U + 2] = (double)d.nrm_in[3 * (int64_t)( li - 1 ) + 2];
  }
}

__global__ void k_pack_unique( const Batch b, int cloud, int16_t* __restrict__ pos, uint8_t* __restrict__ col ) {
  const int64_t iu = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( iu >= (int64_t)b.ucount[cloud] ) { return; }
  const int64_t U = b.off[cloud] + iu;
  const short4  p = b.u_pos[U];
  if ( pos ) { pos[3 * iu] = p.x, pos[3 * iu + 1] = p.y, pos[3 * iu + 2] = p.z; }
  if ( col ) {
    const uchar4 q = b.u_col[U];
    col[3 * iu] = q.x, col[3 * iu + 1] = q.y, col[3 * iu + 2] = q.z;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct MetricsScratch {  // lives in the context's scratch buffers (grow-only)
  RbBuf u_yuv, descs, ndescs;
  RbBuf in_pos, in_col, raw, key_a, key_b, first, u_pos, u_z, u_col, u_orig, tab, sums, small, nrm, nrm_cnt, last_idx,
      nrm_raw, partial, far_list, pend_mask, pend_off, pend_list, contrib, seg;
  // Host clouds (positions, colours, normals) come in on a copy stream, one chunk of pairs ahead of the kernels: two
  // sets of import buffers, `cur` = the set the running chunk reads.
  RbBuf                rawSet[2], nrmSet[2];
  // source clouds kept across calls (rb200_metrics_cache_sources): their part of the batch — imported points, unique
  // points, column tables, gathered normals — stays valid while the same device clouds come in with the same settings
  struct SrcKey {
    const void *pos, *col, *nrm;
    int64_t     n;
    bool        operator==( const SrcKey& o ) const { return pos == o.pos && col == o.col && nrm == o.nrm && n == o.n; }
  };
  bool                 cache_on = false, cache_valid = false;
  bool                 skip_sources = false;  // this call's prefetch left the (cached) sources where they are
  std::vector<SrcKey>  cache_keys;
  int                  cache_drop = 0, cache_c2p = 0, cache_ox = 0, cache_oy = 0, cache_oz = 0, cache_dim = 0;
  RbKdBuild            kd;  // neighborsProc 0: the forest over the unique clouds
  RbBuf                kd_pos, kd_off, first_idx;
  int                  cur = 0;
  bool                 prefetched = false;  // the running chunk's clouds are already in rawSet[cur] / nrmSet[cur]
  cudaStream_t         copy_stream = nullptr;
  cudaEvent_t          ev_free[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  std::vector<int64_t> raw_off;  // per cloud: offset of its raw positions inside the import buffer (-1: resident frame)
};

}  // namespace
void rb_metrics_release( rb200_ctx* c ) {
  MetricsScratch* s = static_cast<MetricsScratch*>( c->metrics_scratch );
  if ( !s ) { return; }
  RbBuf* bufs[] = {&s->in_pos, &s->in_col, &s->raw, &s->key_a, &s->key_b, &s->first, &s->u_pos, &s->u_z, &s->u_col, &s->u_orig,
                   &s->tab, &s->sums, &s->small, &s->nrm, &s->nrm_cnt, &s->last_idx, &s->nrm_raw, &s->partial, &s->far_list,
                   &s->pend_mask, &s->pend_off, &s->pend_list, &s->contrib, &s->seg};
  for ( auto* b : bufs ) { b->release(); }
  s->u_yuv.release();
  s->kd.release();
  s->kd_pos.release();
  s->kd_off.release();
  s->first_idx.release();
  s->descs.release();
  s->ndescs.release();
  for ( int k = 0; k < 2; k++ ) {
    s->rawSet[k].release();
    s->nrmSet[k].release();
    if ( s->ev_free[k] ) { cudaEventDestroy( s->ev_free[k] ); }
    if ( s->ev_copied[k] ) { cudaEventDestroy( s->ev_copied[k] ); }
  }
  if ( s->copy_stream ) { cudaStreamDestroy( s->copy_stream ); }
  delete s;
  c->metrics_scratch = nullptr;
}
namespace {
MetricsScratch* scratch_of( rb200_ctx* c ) {  // one context is driven by one host thread at a time
  if ( !c->metrics_scratch ) { c->metrics_scratch = new MetricsScratch; }
  return static_cast<MetricsScratch*>( c->metrics_scratch );
}

float get_psnr( float dist, float p, float factor = 1.0 ) {  // PCCMetrics.cpp:44-48 (log10 on a float is log10f)
  float max_energy = p * p;
  float psnr       = 10 * log10f( ( factor * max_energy ) / dist );
  return psnr;
}

void finish_quality( rb200_quality& q, const Acc& a, int64_t num, const rb200_metrics_params& mp, bool haveNormals ) {
  memset( &q, 0, sizeof( q ) );
  q.num     = num;
  q.sse_c2c = (double)a.sse_c2c;
  q.max_c2c = std::max( 2.2250738585072014e-308, (double)a.max_c2c );  // starts at numeric_limits<double>::min(), :76
  double maxC2p;
  memcpy( &maxC2p, &a.max_c2p_bits, 8 );
  q.max_c2p = std::max( 2.2250738585072014e-308, maxC2p );
  q.sse_c2p = ( mp.compute_c2p && haveNormals ) ? a.sum_c2p : 0.0;
  for ( int k = 0; k < 3; k++ ) { q.sse_color[k] = mp.compute_color ? a.sum_col[k] : 0.0; }
  const float p = mp.resolution;
  if ( mp.compute_c2c ) {  // :204-211
    q.c2c_mse  = float( q.sse_c2c / num );
    q.c2c_psnr = get_psnr( q.c2c_mse, p, 3 );
    if ( mp.compute_hausdorff ) {
      q.c2c_hausdorff      = float( q.max_c2c );
      q.c2c_hausdorff_psnr = get_psnr( q.c2c_hausdorff, p, 3 );
    }
  }
  if ( mp.compute_c2p ) {  // :213-220
    q.c2p_mse  = float( q.sse_c2p / num );
    q.c2p_psnr = get_psnr( q.c2p_mse, p, 3 );
    if ( mp.compute_hausdorff ) {
      q.c2p_hausdorff      = float( q.max_c2p );
      q.c2p_hausdorff_psnr = get_psnr( q.c2p_hausdorff, p, 3 );
    }
  }
  if ( mp.compute_color ) {  // :221-226
    for ( int k = 0; k < 3; k++ ) {
      q.color_mse[k]  = float( q.sse_color[k] / num );
      q.color_psnr[k] = get_psnr( q.color_mse[k], 1.0 );
    }
  }
}

void combine_quality( rb200_quality& f, const rb200_quality& a, const rb200_quality& b, const rb200_metrics_params& mp ) {
  memset( &f, 0, sizeof( f ) );  // QualityMetrics::operator+, :299-332
  if ( mp.compute_c2c ) {
    f.c2c_mse  = std::max( a.c2c_mse, b.c2c_mse );
    f.c2c_psnr = std::min( a.c2c_psnr, b.c2c_psnr );
  }
  if ( mp.compute_c2p ) {
    f.c2p_mse  = std::max( a.c2p_mse, b.c2p_mse );
    f.c2p_psnr = std::min( a.c2p_psnr, b.c2p_psnr );
  }
  if ( mp.compute_hausdorff ) {
    if ( mp.compute_c2c ) {
      f.c2c_hausdorff      = std::max( a.c2c_hausdorff, b.c2c_hausdorff );
      f.c2c_hausdorff_psnr = std::min( a.c2c_hausdorff_psnr, b.c2c_hausdorff_psnr );
    }
    if ( mp.compute_c2p ) {
      f.c2p_hausdorff      = std::max( a.c2p_hausdorff, b.c2p_hausdorff );
      f.c2p_hausdorff_psnr = std::min( a.c2p_hausdorff_psnr, b.c2p_hausdorff_psnr );
    }
  }
  if ( mp.compute_color ) {
    for ( int k = 0; k < 3; k++ ) {
      f.color_mse[k]  = std::max( a.color_mse[k], b.color_mse[k] );
      f.color_psnr[k] = std::min( a.color_psnr[k], b.color_psnr[k] );
    }
  }
}

struct CloudIn {
  const rb200_cloud_view* view;      // external cloud (host or device pointers), or
  int                     resident;  // frame index of the GOF resident in the context (view == nullptr)
  int64_t                 n;
};

constexpr int RB_REBUILD = -1000;  // build_batch: the cached source part cannot be reused after all (caller retries)

// import the clouds and build the column CSR of every one of them.  `B` receives the device view.  With `nKeep` > 0 the
// first nKeep clouds (the cached sources) are already in place: only the clouds behind them are imported and indexed.
int build_batch( rb200_ctx* c, MetricsScratch* S, const std::vector<CloudIn>& clouds, int drop, Batch& B,
                 std::vector<int64_t>& hOff, int nKeep = 0 ) {
  const int nC = (int)clouds.size();
  hOff.assign( nC + 1, 0 );
  for ( int i = 0; i < nC; i++ ) { hOff[i + 1] = hOff[i] + clouds[i].n; }
  const int64_t N = hOff[nC], N0 = hOff[nKeep];  // N0: first point that is (re)built
  if ( N >= ( 1ll << 31 ) ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "metrics batch larger than 2^31 points" ); }
  if ( nKeep > 0 ) {  // every array must still hold the kept part: growing one would lose it
    const size_t need8 = (size_t)( N + 1 ) * 8, need4 = (size_t)( N + 8 ) * 4;
    if ( S->in_pos.cap < need8 || S->in_col.cap < need4 || S->key_a.cap < need8 || S->key_b.cap < need8 || S->first.cap < need4 ||
         S->u_pos.cap < need8 || S->u_z.cap < (size_t)( N + 1 ) * 2 || S->u_col.cap < need4 || S->u_yuv.cap < (size_t)( N + 1 ) * 16 ||
         S->u_orig.cap < need4 ) {
      return RB_REBUILD;
    }
  }
  RB_CUDA( S->in_pos.ensure( (size_t)( N + 1 ) * 8 ) );
  RB_CUDA( S->in_col.ensure( (size_t)( N + 1 ) * 4 ) );
  RB_CUDA( S->small.ensure( 4096 + (size_t)( nC + 1 ) * 16 ) );
  int64_t maxRaw = 0;
  for ( auto& cl : clouds ) {
    if ( cl.view ) { maxRaw = std::max( maxRaw, cl.n ); }
  }
  // ---- import ----
  short4* inPos = S->in_pos.as<short4>();
  uchar4* inCol = S->in_col.as<uchar4>();
  RbBuf& rawBuf = S->rawSet[S->cur];
  if ( maxRaw > 0 && !S->prefetched ) { RB_CUDA( rawBuf.ensure( (size_t)N * 9 + (size_t)nC * 32 + 64 ) ); }
  if ( S->prefetched ) { RB_CUDA( cudaStreamWaitEvent( c->stream, S->ev_copied[S->cur], 0 ) ); }
  int64_t rawOff = 0;
  S->raw_off.assign( nC, -1 );
  std::vector<ImportDesc> hd;
  int64_t                 maxN = 0;
  for ( int i = 0; i < nC; i++ ) {
    const CloudIn& cl = clouds[i];
    if ( i < nKeep ) {  // already imported by an earlier call; its bytes still lead the prefetched set
      if ( cl.view && cl.n > 0 ) { rawOff += ( ( cl.n * 6 + 15 ) & ~15ll ) + ( ( cl.n * 3 + 15 ) & ~15ll ); }
      continue;
    }
    ImportDesc     d  = ImportDesc{nullptr, nullptr, cl.n, hOff[i], 0, 0};
    maxN              = std::max( maxN, cl.n );
    if ( cl.n == 0 ) {
      hd.push_back( d );
      continue;
    }
    if ( cl.view ) {
      S->raw_off[i] = rawOff;
      char*   rp = rawBuf.as<char>() + rawOff;
      char*   rc = rp + ( ( cl.n * 6 + 15 ) & ~15ll );
      if ( !S->prefetched ) {
        RB_CUDA( cudaMemcpyAsync( rp, cl.view->positions, cl.n * 6, cudaMemcpyDefault, c->stream ) );
        if ( cl.view->colors ) { RB_CUDA( cudaMemcpyAsync( rc, cl.view->colors, cl.n * 3, cudaMemcpyDefault, c->stream ) ); }
        c->stats.h2d_bytes += cl.n * ( cl.view->colors ? 9 : 6 );
      }
      d.pos = rp;
      d.col = cl.view->colors ? rc : nullptr;
      rawOff += ( ( cl.n * 6 + 15 ) & ~15ll ) + ( ( cl.n * 3 + 15 ) & ~15ll );
    } else {
      const int64_t b = c->h_frame_off[cl.resident];
      d.pos      = c->d_pos.as<short4>() + b;
      d.col      = c->d_rgb.as<uchar4>() + b;
      d.resident = 1;
    }
    hd.push_back( d );
  }
  if ( maxN > 0 ) {
    ImportDesc* hp = (ImportDesc*)rb_pinned_ring( c, hd.size() * sizeof( ImportDesc ) + 64 );
    if ( !hp ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
    memcpy( hp, hd.data(), hd.size() * sizeof( ImportDesc ) );
    RB_CUDA( S->descs.ensure( hd.size() * sizeof( ImportDesc ) + 64 ) );
    RB_CUDA( cudaMemcpyAsync( S->descs.p, hp, hd.size() * sizeof( ImportDesc ), cudaMemcpyHostToDevice, c->stream ) );
    const dim3 grid( (unsigned)std::min<int64_t>( rb_div_up( maxN, 4 * 256 ), 2048 ), (unsigned)hd.size() );
    RB_LAUNCH( "met_import", k_import_clouds, grid, 256, 0, S->descs.as<ImportDesc>(), inPos, inCol );
  }
  // ---- bounding box -> table geometry (one small read-back) ----
  int* dSmall = S->small.as<int>();
  {
    int* h  = (int*)rb_pinned( c, 64 );
    int* hi = (int*)rb_pinned_ring( c, 64 );
    if ( !h || !hi ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
    hi[0] = hi[1] = hi[2] = 32767;
    hi[3] = hi[4] = hi[5] = -32768;
    RB_CUDA( cudaMemcpyAsync( dSmall, hi, 24, cudaMemcpyHostToDevice, c->stream ) );
    if ( N > N0 ) { RB_LAUNCH( "met_bbox", k_bbox, 148 * 4, TPB, 0, inPos + N0, N - N0, dSmall ); }
    RB_CUDA( cudaMemcpyAsync( h, dSmall, 24, cudaMemcpyDeviceToHost, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
    c->stats.d2h_bytes += 24;
    if ( N == N0 ) { h[0] = h[1] = h[2] = h[3] = h[4] = h[5] = 0; }
    if ( nKeep > 0 ) {
      // the tables of the kept clouds fix the geometry: the new clouds must fit into it
      if ( N > N0 && ( h[0] < S->cache_ox || h[1] < S->cache_oy || h[3] >= S->cache_ox + S->cache_dim || h[4] >= S->cache_oy + S->cache_dim ) ) {
        return RB_REBUILD;
      }
      B.ox = S->cache_ox, B.oy = S->cache_oy, B.dim = S->cache_dim;
      B.oz = N > N0 ? std::min( S->cache_oz, h[2] ) : S->cache_oz;
    } else {
      // (with the source cache on, a margin lets the reconstructions of later calls fit the same tables)
      const int m = S->cache_on ? 16 : 0;
      B.ox  = h[0] - m;
      B.oy  = h[1] - m;
      B.oz  = N == 0 ? 0 : h[2];
      B.dim = std::max( h[3] - h[0], h[4] - h[1] ) + 1 + 2 * m;
    }
  }
  B.nClouds = nC;
  B.stride  = (int64_t)B.dim * B.dim + 1;
  B.drop    = drop;
  const int64_t tabN = (int64_t)nC * B.stride, tab0 = (int64_t)nKeep * B.stride;
  if ( tabN >= ( 1ll << 32 ) ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "metrics column table too large" ); }
  if ( nKeep > 0 && S->tab.cap < (size_t)( tabN + 4 ) * 4 ) { return RB_REBUILD; }
  RB_CUDA( S->tab.ensure( (size_t)( tabN + 4 ) * 4 ) );
  RB_CUDA( S->sums.ensure( rb_scan_scratch_bytes( std::max( tabN, N + 1 ) ) ) );
  RB_CUDA( S->key_a.ensure( (size_t)( N + 1 ) * 8 ) );
  RB_CUDA( S->key_b.ensure( (size_t)( N + 1 ) * 8 ) );
  RB_CUDA( S->first.ensure( (size_t)( N + 8 ) * 4 ) );
  RB_CUDA( S->u_pos.ensure( (size_t)( N + 1 ) * 8 ) );
  RB_CUDA( S->u_z.ensure( (size_t)( N + 1 ) * 2 ) );
  RB_CUDA( S->u_col.ensure( (size_t)( N + 1 ) * 4 ) );
  RB_CUDA( S->u_yuv.ensure( (size_t)( N + 1 ) * 16 ) );
  RB_CUDA( S->u_orig.ensure( (size_t)( N + 1 ) * 4 ) );
  // offsets + unique counts live in the small block: [64 B bbox][off: (nC+1) x 8][ucount: nC x 4]
  int64_t*  dOff    = (int64_t*)( S->small.as<char>() + 64 );
  uint32_t* dUcount = (uint32_t*)( S->small.as<char>() + 64 + ( nC + 1 ) * 8 );
  const uint32_t* hSeed = nullptr;  // (pinned) the slot the first rebuilt column starts at
  {
    int64_t* h = (int64_t*)rb_pinned_ring( c, ( nC + 1 ) * 8 + 64 );
    if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
    memcpy( h, hOff.data(), ( nC + 1 ) * 8 );
    *(uint32_t*)( h + nC + 1 ) = (uint32_t)N0;
    hSeed                      = (const uint32_t*)( h + nC + 1 );
    RB_CUDA( cudaMemcpyAsync( dOff, h, ( nC + 1 ) * 8, cudaMemcpyHostToDevice, c->stream ) );
  }
  B.off    = dOff;
  B.tab    = S->tab.as<uint32_t>();
  B.in_pos = inPos;
  B.in_col = inCol;
  B.key_a  = S->key_a.as<uint64_t>();
  B.key_b  = S->key_b.as<uint64_t>();
  B.first  = S->first.as<uint32_t>();
  B.u_pos  = S->u_pos.as<short4>();
  B.u_z    = S->u_z.as<int16_t>();
  B.u_col  = S->u_col.as<uchar4>();
  B.u_yuv  = S->u_yuv.as<float4>();
  B.u_orig = S->u_orig.as<uint32_t>();
  B.ucount = dUcount;
  RB_CUDA( cudaMemsetAsync( B.tab + tab0, 0, (size_t)( tabN - tab0 ) * 4, c->stream ) );
  if ( N > N0 ) {
    const int G = rb_div_up( N - N0, TPB );
    RB_LAUNCH( "met_column_count", k_column_count, G, TPB, 0, B, N0, N );
    // the rebuilt tables continue behind the kept slots: the (empty) lead entry of their first cloud carries N0 into the
    // exclusive scan and is set to N0 itself afterwards
    if ( N0 > 0 ) { RB_CUDA( cudaMemcpyAsync( B.tab + tab0, hSeed, 4, cudaMemcpyHostToDevice, c->stream ) ); }
    int r = rb_scan_u32( c, B.tab + tab0, B.tab + tab0, tabN - tab0, S->sums.as<uint32_t>() );
    if ( r ) { return r; }
    if ( N0 > 0 ) { RB_CUDA( cudaMemcpyAsync( B.tab + tab0, hSeed, 4, cudaMemcpyHostToDevice, c->stream ) ); }
    RB_LAUNCH( "met_column_scatter", k_column_scatter, G, TPB, 0, B, N0, N );
    RB_LAUNCH( "met_column_rank", k_column_rank, G, TPB, 0, B, N0, N );
    RB_CUDA( cudaMemsetAsync( B.first + N, 0, 4, c->stream ) );
    r = rb_scan_u32( c, B.first + N0, B.first + N0, N - N0 + 1, S->sums.as<uint32_t>() );
    if ( r ) { return r; }
    RB_LAUNCH( "met_column_compact", k_column_compact, G, TPB, 0, B, N0, N );
  } else if ( N0 == 0 ) {
    RB_CUDA( cudaMemsetAsync( B.first, 0, 8, c->stream ) );
  }
  RB_LAUNCH( "met_table_unique", k_table_unique, 148 * 8, TPB, 0, B, nKeep );
  return RB200_OK;
}

template <int MODE>
int run_nn( rb200_ctx* c, MetricsScratch* S, NNArgs& a, const std::vector<Direction>& dirs, int totalBlocks, int64_t farCap ) {
  if ( dirs.empty() || totalBlocks == 0 ) { return RB200_OK; }
  RB_CUDA( cudaMemsetAsync( a.far_count, 0, 4, c->stream ) );
  const char* nNear = MODE == MODE_SCALE ? "met_nn_scale" : ( MODE == MODE_FILL ? "met_nn_fill" : "met_nn_metric" );
  const char* nPend = MODE == MODE_SCALE ? "met_nn_scale_rings" : ( MODE == MODE_FILL ? "met_nn_fill_rings" : "met_nn_metric_rings" );
  const char* nFar  = MODE == MODE_SCALE ? "met_nn_scale_far" : ( MODE == MODE_FILL ? "met_nn_fill_far" : "met_nn_metric_far" );
  const int64_t nWarps = (int64_t)totalBlocks * ( TPB / 32 );
  RB_CUDA( cudaMemsetAsync( a.pend_off + nWarps, 0, 4, c->stream ) );
  RB_LAUNCH( nNear, k_nn_near<MODE>, totalBlocks, TPB, 0, a );
  int r = rb_scan_u32( c, a.pend_off, a.pend_off, nWarps + 1, S->sums.as<uint32_t>() );
  if ( r ) { return r; }
  uint32_t* h = (uint32_t*)rb_pinned( c, 64 );
  if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  RB_CUDA( cudaMemcpyAsync( h, a.pend_off + nWarps, 4, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  a.nPending = h[0];
  if ( a.nPending ) {
    RB_CUDA( S->pend_list.ensure( (size_t)a.nPending * 4 ) );
    a.pend_list = S->pend_list.as<uint32_t>();
    if ( MODE == MODE_METRIC ) {
      RB_CUDA( S->contrib.ensure( (size_t)a.nPending * 32 ) );
      a.contrib = S->contrib.as<double>();
    }
    RB_LAUNCH( "met_nn_list", k_nn_list_pending, rb_div_up( nWarps, TPB ), TPB, 0, a, nWarps );
    RB_LAUNCH( nPend, k_nn_pending<MODE>, rb_div_up( a.nPending, TPB ), TPB, 0, a );
  }
  RB_LAUNCH( nFar, k_nn_far<MODE>, 148 * 2, TPB, 0, a );
  (void)farCap;
  return RB200_OK;
}

}  // namespace

extern "C" {

// pairs [first, first + nPairs) of one rb200_metrics call; their host clouds are already in the import buffers of
// set S->cur (prefetch_pairs)
static int metrics_range( rb200_ctx* c, const rb200_metrics_params* mp, int first, int nPairs, const rb200_cloud_view* sources,
                          const rb200_cloud_view* recs, rb200_metrics_result* results, bool cacheable ) {
  MetricsScratch* S = scratch_of( c );
  // clouds: the sources first, then the reconstructions (pair i = clouds i and nPairs + i)
  std::vector<CloudIn> clouds( 2 * nPairs );
  bool                 anyNormals = false;
  for ( int i = 0; i < nPairs; i++ ) {
    if ( !sources[i].positions || sources[i].count <= 0 ) { return rb_fail( c, RB200_ERR_INVALID, "metrics: empty source cloud %d", i ); }
    clouds[i] = CloudIn{&sources[i], -1, sources[i].count};
    if ( recs[i].positions ) {
      clouds[nPairs + i] = CloudIn{&recs[i], -1, recs[i].count};
    } else {
      const int fr = first + i;
      if ( !c->reconstructed || !c->rgb_done || fr >= c->F ) {
        return rb_fail( c, RB200_ERR_STATE, "metrics: pair %d asks for the resident frame but no decoded GOF is resident", fr );
      }
      clouds[nPairs + i] = CloudIn{nullptr, fr, c->h_frame_off[fr + 1] - c->h_frame_off[fr]};
    }
    if ( clouds[nPairs + i].n <= 0 ) { return rb_fail( c, RB200_ERR_INVALID, "metrics: empty reconstruction %d", i ); }
    if ( sources[i].normals ) { anyNormals = true; }
  }
  // ---- normals of the sources: uploaded by prefetch_pairs at these offsets of nrmSet[cur] ----
  std::vector<int64_t> nrmOff( nPairs, -1 );
  if ( mp->compute_c2p != 0 && anyNormals ) {
    int64_t tot = 0;
    for ( int i = 0; i < nPairs; i++ ) {
      if ( !sources[i].normals ) { continue; }
      nrmOff[i] = tot;
      tot += ( sources[i].count * 12 + 255 ) & ~255ll;
    }
  }
  // ---- cached sources: the same device clouds with the same settings as the previous call keep their part ----
  const bool wantC2p = mp->compute_c2p != 0 && anyNormals;
  std::vector<MetricsScratch::SrcKey> keys;
  if ( cacheable && S->cache_on ) {
    for ( int i = 0; i < nPairs; i++ ) { keys.push_back( {sources[i].positions, sources[i].colors, sources[i].normals, sources[i].count} ); }
  }
  bool reuse = !keys.empty() && S->cache_valid && keys == S->cache_keys && S->cache_drop == mp->drop_duplicates &&
               S->cache_c2p == ( wantC2p ? 1 : 0 );
  if ( reuse && wantC2p ) {  // the gathered normals of the sources must survive as well
    int64_t n = 0;
    for ( auto& cl : clouds ) { n += cl.n; }
    if ( S->nrm.cap < (size_t)( n + 1 ) * 24 || S->nrm_cnt.cap < (size_t)( n + 1 ) * 4 ) { reuse = false; }
  }
  S->cache_valid = false;
  if ( !reuse && S->skip_sources ) { return RB_REBUILD; }  // the sources were not copied in: the caller starts over
  Batch                B{};
  std::vector<int64_t> hOff;
  int                  r = build_batch( c, S, clouds, mp->drop_duplicates, B, hOff, reuse ? nPairs : 0 );
  if ( r == RB_REBUILD ) {
    if ( S->skip_sources ) { return RB_REBUILD; }
    reuse = false;
    r     = build_batch( c, S, clouds, mp->drop_duplicates, B, hOff, 0 );
  }
  if ( r ) { return r; }
  const int     nC = 2 * nPairs;
  const int64_t N = hOff[nC], N0 = reuse ? hOff[nPairs] : 0;

  // ---- directions and CTA ranges ----
  auto makeDirs = [&]( std::vector<Direction>& dirs, std::vector<int32_t>& blockEnd, auto pick ) {
    int blocks = 0;
    for ( int i = 0; i < nPairs; i++ ) {
      for ( int k = 0; k < 2; k++ ) {
        Direction d{};
        if ( !pick( i, k, d ) ) { continue; }
        d.block_begin = blocks;
        blocks += rb_div_up( clouds[d.cloudA].n, TPB );
        dirs.push_back( d );
        blockEnd.push_back( blocks );
      }
    }
    return blocks;
  };
  std::vector<Direction> dMetric, dScale, dFill;
  std::vector<int32_t>   eMetric, eScale, eFill;
  const int bMetric = makeDirs( dMetric, eMetric, [&]( int i, int k, Direction& d ) {
    const bool hn = sources[i].normals != nullptr;
    d.cloudA      = k == 0 ? i : nPairs + i;
    d.cloudB      = k == 0 ? nPairs + i : i;
    d.nrmA = d.nrmB = hn ? 0 : -1;
    d.acc           = 2 * i + k;
    return true;
  } );
  int bScale = 0, bFill = 0;
  if ( wantC2p ) {
    bScale = makeDirs( dScale, eScale, [&]( int i, int k, Direction& d ) {  // source (with normals) -> reconstruction
      if ( k != 0 || !sources[i].normals ) { return false; }
      d.cloudA = i, d.cloudB = nPairs + i, d.nrmA = d.nrmB = 0, d.acc = 0;
      return true;
    } );
    bFill = makeDirs( dFill, eFill, [&]( int i, int k, Direction& d ) {  // reconstruction -> source for count == 0
      if ( k != 1 || !sources[i].normals ) { return false; }
      d.cloudA = nPairs + i, d.cloudB = i, d.nrmA = d.nrmB = 0, d.acc = 0;
      return true;
    } );
  }
  const int    maxBlocks = std::max( bMetric, std::max( bScale, bFill ) );
  const size_t szDirs    = ( dMetric.size() + dScale.size() + dFill.size() ) * sizeof( Direction );
  const size_t szEnds    = ( eMetric.size() + eScale.size() + eFill.size() ) * 4;
  const size_t szAcc     = (size_t)nC * sizeof( Acc );
  auto         al        = []( size_t x ) { return ( x + 255 ) & ~size_t( 255 ); };
  RB_CUDA( c->d_scratch[4].ensure( al( szDirs ) + al( szEnds ) + al( szAcc ) + 1024 ) );
  char*      dS    = c->d_scratch[4].as<char>();
  Direction* dDirs = (Direction*)dS;
  int32_t*   dEnds = (int32_t*)( dS + al( szDirs ) );
  Acc*       dAcc  = (Acc*)( dS + al( szDirs ) + al( szEnds ) );
  uint32_t*  dFarCount = (uint32_t*)( dS + al( szDirs ) + al( szEnds ) + al( szAcc ) );
  uint32_t*  dErr      = dFarCount + 16;
  {
    char* h = (char*)rb_pinned_ring( c, al( szDirs ) + al( szEnds ) + 64 );
    if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
    Direction* hd = (Direction*)h;
    int32_t*   he = (int32_t*)( h + al( szDirs ) );
    size_t     k  = 0;
    for ( auto* v : {&dMetric, &dScale, &dFill} ) {
      for ( auto& d : *v ) { hd[k++] = d; }
    }
    k = 0;
    for ( auto* v : {&eMetric, &eScale, &eFill} ) {
      for ( auto e : *v ) { he[k++] = e; }
    }
    RB_CUDA( cudaMemcpyAsync( dS, h, al( szDirs ) + al( szEnds ), cudaMemcpyHostToDevice, c->stream ) );
    RB_CUDA( cudaMemsetAsync( dAcc, 0, al( szAcc ) + 1024, c->stream ) );
    c->stats.h2d_bytes += (int64_t)( szDirs + szEnds );
  }
  RB_CUDA( S->partial.ensure( (size_t)std::max( maxBlocks, 1 ) * 32 * ( TPB / 32 ) ) );
  RB_CUDA( S->far_list.ensure( (size_t)( N + 1 ) * 8 ) );
  RB_CUDA( S->pend_mask.ensure( (size_t)( std::max( maxBlocks, 1 ) * ( TPB / 32 ) + 8 ) * 4 ) );
  RB_CUDA( S->pend_off.ensure( (size_t)( std::max( maxBlocks, 1 ) * ( TPB / 32 ) + 8 ) * 4 ) );
  RB_CUDA( S->sums.ensure( rb_scan_scratch_bytes( (int64_t)std::max( maxBlocks, 1 ) * ( TPB / 32 ) + 8 ) ) );
  NNArgs a{};
  a.b              = B;
  a.acc            = dAcc;
  a.partial        = S->partial.as<double>();
  a.far_list       = S->far_list.as<uint32_t>();
  a.far_count      = dFarCount;
  a.pend_mask      = S->pend_mask.as<uint32_t>();
  a.pend_off       = S->pend_off.as<uint32_t>();
  a.compute_c2p    = wantC2p ? 1 : 0;
  a.compute_color  = mp->compute_color;
  a.neighbors_proc = mp->neighbors_proc;
  a.nPairs         = nPairs;

  // ---- normals: copyNormals on the sources, scaleNormals on the reconstructions (PCCMetrics.cpp:371-375) ----
  if ( wantC2p ) {
    RB_CUDA( S->nrm.ensure( (size_t)( N + 1 ) * 24 ) );
    RB_CUDA( S->nrm_cnt.ensure( (size_t)( N + 1 ) * 4 ) );
    RB_CUDA( S->last_idx.ensure( (size_t)( N + 1 ) * 4 ) );
    // (the sources' gathered normals are final after k_normal_gather: a reused call only clears the reconstructions')
    RB_CUDA( cudaMemsetAsync( S->nrm.as<char>() + (size_t)N0 * 24, 0, (size_t)( N - N0 ) * 24, c->stream ) );
    RB_CUDA( cudaMemsetAsync( S->nrm_cnt.as<char>() + (size_t)N0 * 4, 0, (size_t)( N - N0 ) * 4, c->stream ) );
    if ( !reuse ) { RB_CUDA( cudaMemsetAsync( S->last_idx.p, 0, (size_t)N * 4, c->stream ) ); }
    a.nrm     = S->nrm.as<double>();
    a.nrm_cnt = S->nrm_cnt.as<uint32_t>();
    std::vector<NormalDesc> hn;
    int64_t                 maxN = 0;
    for ( int i = 0; i < nPairs && !reuse; i++ ) {
      if ( !sources[i].normals ) { continue; }
      // the normal cloud of pair i is the source view itself (positions + normals in file order): its positions were
      // imported by build_batch, its normals came in on the copy stream
      hn.push_back( NormalDesc{(const float*)( S->nrmSet[S->cur].as<char>() + nrmOff[i] ), sources[i].count, i, 0} );
      maxN = std::max<int64_t>( maxN, sources[i].count );
    }
    if ( !hn.empty() ) {
      NormalDesc* hp = (NormalDesc*)rb_pinned_ring( c, hn.size() * sizeof( NormalDesc ) + 64 );
      if ( !hp ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
      memcpy( hp, hn.data(), hn.size() * sizeof( NormalDesc ) );
      RB_CUDA( S->ndescs.ensure( hn.size() * sizeof( NormalDesc ) + 64 ) );
      RB_CUDA( cudaMemcpyAsync( S->ndescs.p, hp, hn.size() * sizeof( NormalDesc ), cudaMemcpyHostToDevice, c->stream ) );
      const dim3 grid( (unsigned)std::min<int64_t>( rb_div_up( maxN, 256 ), 1024 ), (unsigned)hn.size() );
      RB_LAUNCH( "met_normal_lookup", k_normal_lookup, grid, 256, 0, B, S->ndescs.as<NormalDesc>(), S->last_idx.as<uint32_t>() );
      RB_LAUNCH( "met_normal_gather", k_normal_gather, grid, 256, 0, B, S->ndescs.as<NormalDesc>(), S->last_idx.as<uint32_t>(), a.nrm, dErr );
    }
    a.dirs  = dDirs + dMetric.size();
    a.nDirs = (int)dScale.size();
    r       = run_nn<MODE_SCALE>( c, S, a, dScale, bScale, N );
    if ( r ) { return r; }
    a.dirs  = dDirs + dMetric.size() + dScale.size();
    a.nDirs = (int)dFill.size();
    r       = run_nn<MODE_FILL>( c, S, a, dFill, bFill, N );
    if ( r ) { return r; }
  }
  // ---- the two directions of every pair ----
  a.dirs  = dDirs;
  a.nDirs = (int)dMetric.size();
  if ( mp->compute_color && mp->neighbors_proc == 0 && bMetric > 0 ) {
    // the forest over the unique clouds (their sizes are only known on the device)
    uint32_t* hu = (uint32_t*)rb_pinned( c, (size_t)nC * 4 + 64 );
    if ( !hu ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
    RB_CUDA( cudaMemcpyAsync( hu, B.ucount, (size_t)nC * 4, cudaMemcpyDeviceToHost, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
    std::vector<int64_t> coff( nC + 1, 0 );
    for ( int i = 0; i < nC; i++ ) { coff[i + 1] = coff[i] + hu[i]; }
    RB_CUDA( S->kd_pos.ensure( (size_t)( coff[nC] + 1 ) * 8 ) );
    RB_CUDA( S->kd_off.ensure( (size_t)( nC + 1 ) * 8 ) );
    RB_CUDA( S->first_idx.ensure( (size_t)( N + 1 ) * 4 ) );
    int64_t* ho = (int64_t*)rb_pinned_ring( c, (size_t)( nC + 1 ) * 8 );
    if ( !ho ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
    memcpy( ho, coff.data(), (size_t)( nC + 1 ) * 8 );
    RB_CUDA( cudaMemcpyAsync( S->kd_off.p, ho, (size_t)( nC + 1 ) * 8, cudaMemcpyHostToDevice, c->stream ) );
    int64_t maxU = 1;
    for ( int i = 0; i < nC; i++ ) { maxU = std::max<int64_t>( maxU, hu[i] ); }
    RB_LAUNCH( "met_gather_unique", k_gather_unique, dim3( (unsigned)std::min<int64_t>( rb_div_up( maxU, 256 ), 1024 ), (unsigned)nC ), 256, 0,
               B, S->kd_off.as<int64_t>(), S->kd_pos.as<short4>() );
    r = rb_kd_build( c, S->kd, S->kd_pos.as<short4>(), S->kd_off.as<int64_t>(), coff, B.ox, B.oy, B.oz );
    if ( r ) { return r; }
    RB_LAUNCH( "met_nn_first", k_nn_first, bMetric * ( TPB / 128 ), 128, 0, a, S->kd.forest, S->first_idx.as<uint32_t>() );
    a.first_idx = S->first_idx.as<uint32_t>();
  }
  r       = run_nn<MODE_METRIC>( c, S, a, dMetric, bMetric, N );
  if ( r ) { return r; }
  RB_CUDA( S->seg.ensure( dMetric.size() * RED_SEG * 32 + 64 ) );
  RB_LAUNCH( "met_reduce", k_reduce_segments, (unsigned)dMetric.size() * RED_SEG, TPB, 0, a, dEnds, S->seg.as<double>() );
  RB_LAUNCH( "met_reduce", k_reduce_partials, rb_div_up( (int64_t)dMetric.size(), 64 ), 64, 0, a, S->seg.as<double>() );

  // ---- read back ----
  const size_t rbBytes = szAcc + nC * 4 + 64;
  char*        h       = (char*)rb_pinned( c, rbBytes + 64 );
  if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  RB_CUDA( cudaMemcpyAsync( h, dAcc, szAcc, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaMemcpyAsync( h + szAcc, B.ucount, nC * 4, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaMemcpyAsync( h + szAcc + nC * 4, dErr, 4, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  c->stats.d2h_bytes += (int64_t)rbBytes;
  const Acc*      hAcc = (const Acc*)h;
  const uint32_t* hUc  = (const uint32_t*)( h + szAcc );
  const uint32_t  err  = *(const uint32_t*)( h + szAcc + nC * 4 );
  if ( err ) {
    return rb_fail( c, RB200_ERR_INVALID,
                    err & 1 ? "metrics: normal object and source must have the same number of points (PCCPointSet.cpp:2287)"
                            : "metrics: a source point is not present in the normal point cloud (PCCPointSet.cpp:2311)" );
  }
  if ( !keys.empty() ) {  // what the next call may keep
    S->cache_keys  = keys;
    S->cache_drop  = mp->drop_duplicates;
    S->cache_c2p   = wantC2p ? 1 : 0;
    S->cache_ox = B.ox, S->cache_oy = B.oy, S->cache_oz = B.oz, S->cache_dim = B.dim;
    S->cache_valid = true;
  }
  int status = RB200_OK;
  for ( int i = 0; i < nPairs; i++ ) {
    rb200_metrics_result& R = results[i];
    memset( &R, 0, sizeof( R ) );
    const bool hn = sources[i].normals != nullptr && mp->compute_c2p;
    finish_quality( R.q1, hAcc[2 * i], hUc[i], *mp, hn );
    finish_quality( R.q2, hAcc[2 * i + 1], hUc[nPairs + i], *mp, hn );
    combine_quality( R.qf, R.q1, R.q2, *mp );
    R.source_points      = clouds[i].n;
    R.rec_points         = clouds[nPairs + i].n;
    R.source_after_dedup = mp->drop_duplicates ? hUc[i] : 0;  // sourceDuplicates_, PCCMetrics.cpp:356,363
    R.rec_after_dedup    = mp->drop_duplicates ? hUc[nPairs + i] : 0;
    R.tie_overflow       = (int32_t)( hAcc[2 * i].tie_overflow + hAcc[2 * i + 1].tie_overflow );
    if ( R.tie_overflow ) { status = RB200_ERR_TIE_OVERFLOW; }
  }
  if ( status == RB200_ERR_TIE_OVERFLOW ) {
    rb_fail( c, status, "metrics: a nearest-distance shell holds more than 30 points; the reference's result depends on its "
                        "kd-tree traversal order there (results are filled in with the first 30 by index)" );
  }
  return status;
}

// host clouds of pairs [first, first + nPairs) -> import buffers of `set`, on the copy stream (same offsets as
// build_batch / metrics_range compute)
static int prefetch_pairs( rb200_ctx* c, MetricsScratch* S, const rb200_metrics_params* mp, int nPairs, const rb200_cloud_view* sources,
                           const rb200_cloud_view* recs, int set ) {
  if ( !S->copy_stream ) {
    RB_CUDA( cudaStreamCreateWithFlags( &S->copy_stream, cudaStreamNonBlocking ) );
    for ( int k = 0; k < 2; k++ ) {
      RB_CUDA( cudaEventCreateWithFlags( &S->ev_free[k], cudaEventDisableTiming ) );
      RB_CUDA( cudaEventCreateWithFlags( &S->ev_copied[k], cudaEventDisableTiming ) );
      RB_CUDA( cudaEventRecord( S->ev_free[k], c->stream ) );
    }
  }
  int64_t rawBytes = 0, nrmBytes = 0;
  for ( int i = 0; i < nPairs; i++ ) {
    const rb200_cloud_view* v[2] = {&sources[i], recs[i].positions ? &recs[i] : nullptr};
    for ( int k = 0; k < 2; k++ ) {
      if ( v[k] && v[k]->count > 0 ) { rawBytes += ( ( v[k]->count * 6 + 15 ) & ~15ll ) + ( ( v[k]->count * 3 + 15 ) & ~15ll ); }
    }
    if ( mp->compute_c2p && sources[i].normals ) { nrmBytes += ( sources[i].count * 12 + 255 ) & ~255ll; }
  }
  RB_CUDA( S->rawSet[set].ensure( (size_t)rawBytes + (size_t)nPairs * 64 + 64 ) );
  RB_CUDA( S->nrmSet[set].ensure( (size_t)nrmBytes + 64 ) );
  RB_CUDA( cudaStreamWaitEvent( S->copy_stream, S->ev_free[set], 0 ) );  // the chunk that read this set has finished
  int64_t rawOff = 0, nOff = 0;
  // the order of the batch: every source, then every reconstruction that is not the resident one (build_batch walks the
  // clouds in the same order)
  for ( int k = 0; k < 2; k++ ) {
    for ( int i = 0; i < nPairs; i++ ) {
      const rb200_cloud_view* v = k == 0 ? &sources[i] : ( recs[i].positions ? &recs[i] : nullptr );
      if ( !v || v->count <= 0 ) { continue; }
      const int64_t n  = v->count;
      char*         rp = S->rawSet[set].as<char>() + rawOff;
      char*         rc = rp + ( ( n * 6 + 15 ) & ~15ll );
      if ( !( k == 0 && S->skip_sources ) ) {  // (cached sources stay where the earlier call imported them)
        RB_CUDA( cudaMemcpyAsync( rp, v->positions, n * 6, cudaMemcpyDefault, S->copy_stream ) );
        if ( v->colors ) { RB_CUDA( cudaMemcpyAsync( rc, v->colors, n * 3, cudaMemcpyDefault, S->copy_stream ) ); }
        c->stats.h2d_bytes += n * ( v->colors ? 9 : 6 );
      }
      rawOff += ( ( n * 6 + 15 ) & ~15ll ) + ( ( n * 3 + 15 ) & ~15ll );
    }
  }
  for ( int i = 0; i < nPairs; i++ ) {
    if ( mp->compute_c2p && sources[i].normals ) {
      if ( !S->skip_sources ) {
        RB_CUDA( cudaMemcpyAsync( S->nrmSet[set].as<char>() + nOff, sources[i].normals, sources[i].count * 12, cudaMemcpyDefault, S->copy_stream ) );
        c->stats.h2d_bytes += sources[i].count * 12;
      }
      nOff += ( sources[i].count * 12 + 255 ) & ~255ll;
    }
  }
  RB_CUDA( cudaEventRecord( S->ev_copied[set], S->copy_stream ) );
  return RB200_OK;
}

int rb200_metrics( rb200_ctx* c, const rb200_metrics_params* mp, int nPairs, const rb200_cloud_view* sources,
                   const rb200_cloud_view* recs, rb200_metrics_result* results ) {
  if ( !c || !mp || nPairs <= 0 || !sources || !recs || !results ) {
    return rb_fail( c, RB200_ERR_INVALID, "metrics: bad arguments" );
  }
  if ( mp->drop_duplicates < 0 || mp->drop_duplicates > 2 ) {
    return rb_fail( c, RB200_ERR_INVALID, "metrics: drop_duplicates must be 0, 1 or 2" );
  }
  if ( mp->compute_color && ( mp->neighbors_proc < 0 || mp->neighbors_proc > 4 ) ) {
    return rb_fail( c, RB200_ERR_INVALID, "metrics: neighbors_proc %d (0..4)", mp->neighbors_proc );
  }
  for ( int i = 0; i < nPairs; i++ ) {
    if ( !sources[i].positions || sources[i].count <= 0 ) { return rb_fail( c, RB200_ERR_INVALID, "metrics: empty source cloud %d", i ); }
  }
  cudaSetDevice( c->device );
  MetricsScratch* S = scratch_of( c );
  // The pairs are independent (PCCMetrics::compute loops over the frames, PCCMetrics.cpp:348-384).  They are processed
  // in chunks so that the host clouds of chunk k+1 cross PCIe on the copy stream while the kernels of chunk k run.
  // Clouds that are already in device memory (sources cached across calls, resident reconstructions) have nothing to
  // overlap: one chunk, a quarter of the launches and host round trips.
  bool onDevice = true;
  for ( int i = 0; i < nPairs && onDevice; i++ ) {
    const void* ptrs[2] = {sources[i].positions, recs[i].positions};
    for ( const void* q : ptrs ) {
      if ( !q ) { continue; }
      cudaPointerAttributes attr{};
      if ( cudaPointerGetAttributes( &attr, q ) != cudaSuccess || attr.type != cudaMemoryTypeDevice ) {
        cudaGetLastError();
        onDevice = false;
      }
    }
  }
  int64_t total = 0;
  for ( int i = 0; i < nPairs; i++ ) { total += 2 * sources[i].count; }
  const int nDev    = (int)std::min<int64_t>( nPairs, ( total + ( 64ll << 20 ) - 1 ) / ( 64ll << 20 ) );  // scratch is ~100 B per point
  const int nChunks = onDevice ? std::max( nDev, 1 ) : ( nPairs >= 8 ? 4 : 1 ), per = ( nPairs + nChunks - 1 ) / nChunks;
  int       status = RB200_OK;
  // cached sources (rb200_metrics_cache_sources) are not even copied when the key of the previous call matches
  S->skip_sources = false;
  if ( onDevice && nChunks == 1 && S->cache_on && S->cache_valid && (int)S->cache_keys.size() == nPairs &&
       S->cache_drop == mp->drop_duplicates ) {
    S->skip_sources = true;
    for ( int i = 0; i < nPairs && S->skip_sources; i++ ) {
      S->skip_sources = S->cache_keys[i] == MetricsScratch::SrcKey{sources[i].positions, sources[i].colors, sources[i].normals, sources[i].count};
    }
  }
  int       r = prefetch_pairs( c, S, mp, std::min( per, nPairs ), sources, recs, 0 );
  if ( r ) { return r; }
  if ( S->skip_sources ) {
    S->cur = 0, S->prefetched = true;
    r             = metrics_range( c, mp, 0, nPairs, sources, recs, results, true );
    S->prefetched = false;
    cudaEventRecord( S->ev_free[0], c->stream );
    if ( r != RB_REBUILD ) {
      if ( r && r != RB200_ERR_TIE_OVERFLOW ) { cudaStreamSynchronize( S->copy_stream ); }
      S->skip_sources = false;
      return r;
    }
    S->skip_sources = false;  // the kept part did not fit after all: everything again, sources included
    S->cache_valid  = false;
    r               = prefetch_pairs( c, S, mp, std::min( per, nPairs ), sources, recs, 0 );
    if ( r ) { return r; }
  }
  for ( int k = 0, b = 0; b < nPairs; k++, b += per ) {
    const int e = std::min( nPairs, b + per ), set = k & 1;
    if ( e < nPairs ) {
      r = prefetch_pairs( c, S, mp, std::min( per, nPairs - e ), sources + e, recs + e, set ^ 1 );
      if ( r ) { return r; }
    }
    S->cur        = set;
    S->prefetched = true;
    r             = metrics_range( c, mp, b, e - b, sources + b, recs + b, results + b, onDevice && nChunks == 1 );
    S->prefetched = false;
    cudaEventRecord( S->ev_free[set], c->stream );
    if ( r == RB200_ERR_TIE_OVERFLOW ) {
      status = r;
    } else if ( r ) {
      cudaStreamSynchronize( S->copy_stream );  // nothing may still be reading the caller's buffers
      return r;
    }
  }
  return status;
}

int rb200_metrics_cache_sources( rb200_ctx* c, int on ) {
  if ( !c ) { return RB200_ERR_INVALID; }
  MetricsScratch* S = scratch_of( c );
  S->cache_on       = on != 0;
  S->cache_valid    = false;
  S->cache_keys.clear();
  return RB200_OK;
}

int rb200_remove_duplicates( rb200_ctx* c, const rb200_cloud_view* in, int drop, int16_t* outPos, uint8_t* outCol,
                             int64_t* outCount ) {
  if ( !c || !in || !outCount ) { return rb_fail( c, RB200_ERR_INVALID, "remove_duplicates: bad arguments" ); }
  if ( drop < 1 || drop > 2 ) { return rb_fail( c, RB200_ERR_INVALID, "remove_duplicates: drop must be 1 or 2" ); }
  cudaSetDevice( c->device );
  *outCount = 0;
  if ( in->count == 0 ) { return RB200_OK; }
  if ( !in->positions ) { return rb_fail( c, RB200_ERR_INVALID, "remove_duplicates: null positions" ); }
  MetricsScratch*      S = scratch_of( c );
  S->cache_valid         = false;
  std::vector<CloudIn> clouds{CloudIn{in, -1, in->count}};
  Batch                B{};
  std::vector<int64_t> hOff;
  int                  r = build_batch( c, S, clouds, drop, B, hOff );
  if ( r ) { return r; }
  uint32_t* h = (uint32_t*)rb_pinned( c, 64 );
  if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  RB_CUDA( cudaMemcpyAsync( h, B.ucount, 4, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  const int64_t nU = h[0];
  *outCount        = nU;
  if ( ( outPos || outCol ) && nU > 0 ) {
    RB_CUDA( S->raw.ensure( (size_t)nU * 9 + 64 ) );
    int16_t* dp = S->raw.as<int16_t>();
    uint8_t* dc = (uint8_t*)( S->raw.as<char>() + ( ( nU * 6 + 15 ) & ~15ll ) );
    RB_LAUNCH( "met_pack_unique", k_pack_unique, rb_div_up( nU, TPB ), TPB, 0, B, 0, outPos ? dp : nullptr, outCol ? dc : nullptr );
    if ( outPos ) { RB_CUDA( cudaMemcpyAsync( outPos, dp, nU * 6, cudaMemcpyDeviceToHost, c->stream ) ); }
    if ( outCol ) { RB_CUDA( cudaMemcpyAsync( outCol, dc, nU * 3, cudaMemcpyDeviceToHost, c->stream ) ); }
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
    c->stats.d2h_bytes += nU * ( ( outPos ? 6 : 0 ) + ( outCol ? 3 : 0 ) );
  }
  return RB200_OK;
}

// ---- the one exchange step of the path: fixed-size records of the per-frame accumulators ----
// record layout (RB200_METRICS_RECORD doubles): [0] frame, then per direction (q1, q2) {sse_c2c, sse_c2p, sse_color[3],
// max_c2c, max_c2p, num}, then source_points, source_after_dedup, rec_points, rec_after_dedup, tie_overflow, 0, 0
int rb200_metrics_pack( int frame, const rb200_metrics_result* r, double* record ) {
  if ( !r || !record ) { return RB200_ERR_INVALID; }
  int k       = 0;
  record[k++] = (double)frame;
  for ( const rb200_quality* q : {&r->q1, &r->q2} ) {
    record[k++] = q->sse_c2c, record[k++] = q->sse_c2p;
    for ( int c = 0; c < 3; c++ ) { record[k++] = q->sse_color[c]; }
    record[k++] = q->max_c2c, record[k++] = q->max_c2p, record[k++] = (double)q->num;
  }
  record[k++] = (double)r->source_points, record[k++] = (double)r->source_after_dedup;
  record[k++] = (double)r->rec_points, record[k++] = (double)r->rec_after_dedup, record[k++] = (double)r->tie_overflow;
  while ( k < RB200_METRICS_RECORD ) { record[k++] = 0.0; }
  return RB200_OK;
}

// the float results are derived again from the accumulators exactly as QualityMetrics::compute (PCCMetrics.cpp:204-226)
// and QualityMetrics::operator+ (:299-332) do, so every rank ends with the numbers the owning rank computed
int rb200_metrics_unpack( const double* record, const rb200_metrics_params* mp, int* frame, rb200_metrics_result* out ) {
  if ( !record || !mp || !out ) { return RB200_ERR_INVALID; }
  memset( out, 0, sizeof( *out ) );
  int k = 0;
  if ( frame ) { *frame = (int)record[0]; }
  k = 1;
  rb200_quality* qs[2] = {&out->q1, &out->q2};
  for ( int d = 0; d < 2; d++ ) {
    Acc a{};
    a.sse_c2c = (unsigned long long)record[k], a.sum_c2p = record[k + 1];
    for ( int c = 0; c < 3; c++ ) { a.sum_col[c] = record[k + 2 + c]; }
    const double maxC2c = record[k + 5], maxC2p = record[k + 6];
    const int64_t num   = (int64_t)record[k + 7];
    a.max_c2c           = (unsigned long long)maxC2c;
    memcpy( &a.max_c2p_bits, &maxC2p, 8 );
    // sse_c2p travels as the finished sum: keep it whatever the normals were
    rb200_metrics_params m2 = *mp;
    finish_quality( *qs[d], a, num, m2, true );
    qs[d]->max_c2c = maxC2c;
    if ( mp->compute_hausdorff && mp->compute_c2c ) {
      qs[d]->c2c_hausdorff      = float( maxC2c );
      qs[d]->c2c_hausdorff_psnr = get_psnr( qs[d]->c2c_hausdorff, mp->resolution, 3 );
    }
    k += 8;
  }
  combine_quality( out->qf, out->q1, out->q2, *mp );
  out->source_points = (int64_t)record[k], out->source_after_dedup = (int64_t)record[k + 1];
  out->rec_points = (int64_t)record[k + 2], out->rec_after_dedup = (int64_t)record[k + 3], out->tie_overflow = (int32_t)record[k + 4];
  return RB200_OK;
}

}  // extern "C"
