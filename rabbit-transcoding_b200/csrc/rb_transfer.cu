// rb_transfer.cu — attribute re-transfer after geometry smoothing, for every frame of the GOF (sm_100a).
//
// Restates PCCPointSet3::transferColors16bitBP (PccLibCommon/source/PCCPointSet.cpp:1126-1485) exactly as the decoder
// calls it (PccLibDecoder/source/PCCDecoder.cpp:447-465): filterType 1, searchRange 0, 8 forward / 1 backward
// neighbours, distance-weighted averages with offsets 4, skipAvgIfIdenticalSourcePointPresentFwd, and all four
// distance / colour thresholds disabled (1000 >= 512 and 256000 >= 131072 turn them into DBL_MAX, :1153-1156).
// Only the points that geometry smoothing moved (boundary type 3) change colour:
//   forward : 8-NN of the moved point in the PRE-smoothing cloud -> refinedColors1 (identical position: that colour;
//             otherwise the 1/(d^2+4)-weighted mean), and those 8 source points become "partSource" (:1165-1265);
//   backward: every partSource point looks up its 1-NN in the SMOOTHED cloud; if the colours are within 40 per
//             channel it becomes a candidate of that target (:1275-1293); candidates are std::sort-ed by distance;
//   result  : round( sum c / (sqrt(d^2)+4) / sum 1/(sqrt(d^2)+4) ) over the candidates, or refinedColors1 when
//             there is none / the attribute is RGB444 (:1319-1483 with fixWeight, w = 0, searchRange 0).
// Which 8 points tie at the 8th distance, and which of several equidistant targets is "the" 1-NN, is decided by
// nanoflann's traversal order: both searches run on the emulated trees of rb_kdtree.cu.  std::sort's order of
// equal-distance candidates (it is not stable beyond 16 elements) is reproduced with libstdc++'s introsort.
#include <algorithm>

#include "rb_common.cuh"
#include "rb_kdtree.cuh"
#include "rb_kdtree_build.cuh"

namespace {

constexpr int TPB = 256;
constexpr int KF  = 8;  // numNeighborsColorTransferFwd

struct TransferScratch {
  RbKdBuild kd;
  RbBuf     pos2, off, flags, sums, moved, part, partDist, bwdT, partCnt, refined1, candCnt, candOff, candKey, small, claim, srcList, srcNN;
};

TransferScratch* scratch_of( rb200_ctx* c ) {  // one context is driven by one host thread at a time
  if ( !c->transfer_scratch ) { c->transfer_scratch = new TransferScratch; }
  return static_cast<TransferScratch*>( c->transfer_scratch );
}

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

// the moved (type 3) points come as one bit per point from the geometry filter (rb_smooth.cu): population count per word,
// exclusive scan over the words, then every word writes the indices of its set bits and their ranks — 1/32 of the
// traffic of flagging, scanning and listing the points themselves
__global__ void k_moved_count( const uint32_t* __restrict__ bits, int64_t nWords, uint32_t* __restrict__ cnt ) {
  const int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( w > nWords ) { return; }
  cnt[w] = w < nWords ? (uint32_t)__popc( bits[w] ) : 0u;
}
__global__ void k_moved_list( const uint32_t* __restrict__ bits, const uint32_t* __restrict__ scan, int64_t nWords,
                              uint32_t* __restrict__ moved, uint32_t* __restrict__ rank ) {
  const int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( w >= nWords ) { return; }
  uint32_t m = scan[w];
  for ( uint32_t b = bits[w]; b; b &= b - 1 ) {
    const uint32_t g = (uint32_t)( w * 32 ) + (uint32_t)( __ffs( b ) - 1 );
    moved[m] = g;
    rank[g]  = m;  // (only read for type-3 points)
    m++;
  }
}

struct TArgs {
  KdForest        forest;
  int             F;
  const int64_t*  frame_off;  // [F + 1]
  const short4*   posS;       // pre-smoothing positions
  const short4*   posT;       // smoothed positions (w = boundary type)
  ushort4*        col;        // colours16 (source == target before the transfer)
  const uint32_t* moved;      // global indices of the type-3 points, ascending
  const uint32_t* rank;       // scan of the moved flags: rank[g] = position of g in `moved`
  uint32_t        nMoved;
  uint32_t*       part;       // [nMoved][KF] source index inside the frame
  uint32_t*       partDist;   // [nMoved][KF] distance of an accepted backward candidate
  uint32_t*       claim;      // one bit per point of the GOF: listed in srcList
  uint32_t*       srcList;    // the distinct partSource points (global indices), *nSrc of them
  uint32_t*       nSrc;
  uint2*          srcNN;      // [N] nearest smoothed point of a listed source point {index in frame, distance}
  uint32_t*       bwdT;       // [nMoved][KF] rank of the accepted nearest target among the moved points, or ~0
  uint8_t*        partCnt;    // [nMoved]
  ushort4*        refined1;   // [nMoved]
  uint32_t*       candCnt;    // [nMoved + 1]
  const uint32_t* candOff;
  uint64_t*       candKey;    // dist << 32 | partSource index
  int             lossless;
  uint32_t*       err;
};

// forward direction (:1165-1265)
__global__ void __launch_bounds__( 128 ) k_transfer_fwd( const TArgs a ) {
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if ( m >= a.nMoved ) { return; }
  const int64_t g    = a.moved[m];
  const int     f    = frame_of( a.frame_off, a.F, g );
  const int64_t base = a.frame_off[f];
  const short4  p    = a.posT[g];
  const int     q[3] = {p.x - a.forest.ox, p.y - a.forest.oy, p.z - a.forest.oz};
  KdResult<KF>  res;
  kd_search<KF>( a.forest, (uint32_t)f + 1u, q, res );  // kdtreeSource.search( target[index], 8 )
  a.partCnt[m] = (uint8_t)res.count;
#pragma unroll
  for ( int j = 0; j < KF; j++ ) {  // (the result set lives in registers: no dynamic indexing)
    if ( j < res.count ) { a.part[(size_t)m * KF + j] = res.idx[j]; }
  }
  ushort4 out;
  if ( res.dist[0] == 0 || res.count == 1 ) {  // identical source point present (:1186-1191) or a single neighbour (:1195-1198)
    out = a.col[base + res.idx[0]];
  } else {
    // maxColorDist2 <= DBL_MAX always holds: distance-weighted mean of all neighbours (:1212-1222, :1252-1256)
    double rc[3] = {0.0, 0.0, 0.0}, sw = 0.0;
#pragma unroll
    for ( int j = 0; j < KF; j++ ) {
      if ( j < res.count ) {
        const double  w = 1.0 / ( (double)res.dist[j] + 4.0 );
        const ushort4 c = a.col[base + res.idx[j]];
        rc[0]           = __dadd_rn( rc[0], __dmul_rn( (double)c.x, w ) );
        rc[1]           = __dadd_rn( rc[1], __dmul_rn( (double)c.y, w ) );
        rc[2]           = __dadd_rn( rc[2], __dmul_rn( (double)c.z, w ) );
        sw              = __dadd_rn( sw, w );
      }
    }
    out.x = (unsigned short)fmin( fmax( round( rc[0] / sw ), 0.0 ), 65535.0 );
    out.y = (unsigned short)fmin( fmax( round( rc[1] / sw ), 0.0 ), 65535.0 );
    out.z = (unsigned short)fmin( fmax( round( rc[2] / sw ), 0.0 ), 65535.0 );
  }
  out.w         = 0;
  a.refined1[m] = out;
}

// backward direction (:1275-1293): every partSource point looks up its nearest target.  Neighbouring moved targets share
// most of their 8 source points, so the DISTINCT source points are listed first (a claim bit per point, one append per
// warp), searched once each by full warps, and the 8 x nMoved entries then read the answer of their point.  The accepted
// ones are counted per moved target and remembered (target rank, distance); k_transfer_bwd_store files them.
__global__ void __launch_bounds__( 256 ) k_transfer_bwd_claim( const TArgs a ) {
  const uint32_t p    = blockIdx.x * blockDim.x + threadIdx.x;
  const int      lane = threadIdx.x & 31;
  bool           mine = false;
  uint32_t       gs   = 0;
  if ( p < a.nMoved * KF ) {
    const uint32_t m = p / KF, j = p % KF;
    if ( j < a.partCnt[m] ) {
      const int64_t g = a.moved[m];
      const int     f = frame_of( a.frame_off, a.F, g );
      gs              = (uint32_t)( a.frame_off[f] + a.part[p] );
      const uint32_t bit = 1u << ( gs & 31u );
      mine               = ( atomicOr( &a.claim[gs >> 5], bit ) & bit ) == 0;
    }
  }
  const uint32_t b = __ballot_sync( 0xFFFFFFFFu, mine );
  if ( b == 0 ) { return; }
  uint32_t at = 0;
  if ( lane == __ffs( b ) - 1 ) { at = atomicAdd( a.nSrc, (uint32_t)__popc( b ) ); }
  at = __shfl_sync( 0xFFFFFFFFu, at, __ffs( b ) - 1 );
  if ( mine ) { a.srcList[at + __popc( b & ( ( 1u << lane ) - 1u ) )] = gs; }
}
__global__ void __launch_bounds__( 128 ) k_transfer_bwd_search( const TArgs a ) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= *a.nSrc ) { return; }
  const uint32_t gs   = a.srcList[i];
  const int      f    = frame_of( a.frame_off, a.F, gs );
  const short4   ps   = a.posS[gs];
  const int      q[3] = {ps.x - a.forest.ox, ps.y - a.forest.oy, ps.z - a.forest.oz};
  KdResult<1>    res;
  kd_search<1>( a.forest, (uint32_t)( a.F + f ) + 1u, q, res );  // kdtreeTarget.search( partSource[index], 1 )
  a.srcNN[gs] = res.count ? make_uint2( res.idx[0], res.dist[0] ) : make_uint2( 0xFFFFFFFFu, 0u );
}
__global__ void __launch_bounds__( 256 ) k_transfer_bwd( const TArgs a ) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if ( p >= a.nMoved * KF ) { return; }
  const uint32_t m = p / KF, j = p % KF;
  a.bwdT[p]        = 0xFFFFFFFFu;
  if ( j >= a.partCnt[m] ) { return; }
  const int64_t  g    = a.moved[m];
  const int      f    = frame_of( a.frame_off, a.F, g );
  const int64_t  base = a.frame_off[f];
  const uint32_t s    = a.part[p];
  const uint2    nn   = a.srcNN[base + s];
  if ( nn.x == 0xFFFFFFFFu ) { return; }
  const int64_t r = base + nn.x;
  if ( a.posT[r].w != 3 ) { return; }  // only type-3 targets are recomputed (:1319)
  const ushort4 cs = a.col[base + s], ct = a.col[r];
  if ( abs( (int)cs.x - (int)ct.x ) < 40 && abs( (int)cs.y - (int)ct.y ) < 40 && abs( (int)cs.z - (int)ct.z ) < 40 ) {
    const uint32_t mr = a.rank[r];
    atomicAdd( &a.candCnt[mr], 1u );
    a.bwdT[p]     = mr;
    a.partDist[p] = nn.y;
  }
}
__global__ void __launch_bounds__( 256 ) k_transfer_bwd_store( const TArgs a ) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if ( p >= a.nMoved * KF ) { return; }
  const uint32_t mr = a.bwdT[p];
  if ( mr == 0xFFFFFFFFu ) { return; }
  const uint32_t slot = a.candOff[mr] + atomicAdd( &a.candCnt[mr], 1u );
  a.candKey[slot]     = ( (uint64_t)a.partDist[p] << 32 ) | p;
}

// ---- libstdc++ std::sort( first, last, dist < dist ) on an array of keys whose high 32 bits are the distance ----
__device__ __forceinline__ bool k_less( uint64_t x, uint64_t y ) { return ( x >> 32 ) < ( y >> 32 ); }
__device__ __forceinline__ void k_swap( uint64_t& x, uint64_t& y ) {
  const uint64_t t = x;
  x                = y;
  y                = t;
}
__device__ void std_unguarded_linear_insert( uint64_t* v, int last ) {
  const uint64_t val  = v[last];
  int            next = last - 1;
  while ( k_less( val, v[next] ) ) {
    v[last] = v[next];
    last    = next;
    --next;
  }
  v[last] = val;
}
__device__ void std_insertion_sort( uint64_t* v, int first, int last ) {
  if ( first == last ) { return; }
  for ( int i = first + 1; i != last; ++i ) {
    if ( k_less( v[i], v[first] ) ) {
      const uint64_t val = v[i];
      for ( int k = i; k > first; --k ) { v[k] = v[k - 1]; }
      v[first] = val;
    } else {
      std_unguarded_linear_insert( v, i );
    }
  }
}
// returns false when the depth limit would send libstdc++ into its heapsort fallback (not reproduced: fail loudly)
__device__ bool std_sort_emulated( uint64_t* v, int n ) {
  if ( n <= 1 ) { return true; }
  // __introsort_loop( first, last, 2 * lg(n) ) with an explicit stack for the recursive call on [cut, last)
  int stF[64], stL[64], stD[64];
  int sp = 0;
  stF[sp] = 0, stL[sp] = n, stD[sp] = 2 * ( 31 - __clz( n ) );
  sp++;
  while ( sp > 0 ) {
    --sp;
    int first = stF[sp], last = stL[sp], depth = stD[sp];
    while ( last - first > 16 ) {
      if ( depth == 0 ) { return false; }
      --depth;
      // __unguarded_partition_pivot
      const int mid = first + ( last - first ) / 2;
      {  // __move_median_to_first( first, first + 1, mid, last - 1 )
        const int a = first + 1, b = mid, c = last - 1;
        if ( k_less( v[a], v[b] ) ) {
          if ( k_less( v[b], v[c] ) ) {
            k_swap( v[first], v[b] );
          } else if ( k_less( v[a], v[c] ) ) {
            k_swap( v[first], v[c] );
          } else {
            k_swap( v[first], v[a] );
          }
        } else if ( k_less( v[a], v[c] ) ) {
          k_swap( v[first], v[a] );
        } else if ( k_less( v[b], v[c] ) ) {
          k_swap( v[first], v[c] );
        } else {
          k_swap( v[first], v[b] );
        }
      }
      int lo = first + 1, hi = last;  // __unguarded_partition( first + 1, last, first )
      for ( ;; ) {
        while ( k_less( v[lo], v[first] ) ) { ++lo; }
        --hi;
        while ( k_less( v[first], v[hi] ) ) { --hi; }
        if ( !( lo < hi ) ) { break; }
        k_swap( v[lo], v[hi] );
        ++lo;
      }
      const int cut = lo;
      if ( sp >= 64 ) { return false; }
      stF[sp] = cut, stL[sp] = last, stD[sp] = depth;  // __introsort_loop( cut, last, depth_limit )
      sp++;
      last = cut;
    }
  }
  // __final_insertion_sort
  if ( n > 16 ) {
    std_insertion_sort( v, 0, 16 );
    for ( int i = 16; i != n; ++i ) { std_unguarded_linear_insert( v, i ); }
  } else {
    std_insertion_sort( v, 0, n );
  }
  return true;
}

// the pending recursive calls of __introsort_loop run AFTER the current loop finishes in the real recursion order
// "loop body: recurse on [cut,last) first, then continue with [first,cut)".  Since the two ranges are disjoint and
// the algorithm only touches elements of its own range, the order in which disjoint ranges are processed does not
// change the result.

// final colours of the moved targets (:1319-1483)
__global__ void __launch_bounds__( 128 ) k_transfer_final( const TArgs a ) {
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if ( m >= a.nMoved ) { return; }
  const int64_t  g    = a.moved[m];
  const int      f    = frame_of( a.frame_off, a.F, g );
  const uint32_t beg = a.candOff[m], end = a.candOff[m + 1];
  const int      n   = (int)( end - beg );
  ushort4        out = a.refined1[m];
  out.w              = a.col[g].w;  // the layer index travels in .w
  if ( n > 0 && !a.lossless ) {
    uint64_t* v = a.candKey + beg;
    // the candidates were appended in partSource order (:1280-1292) before std::sort: restore that order first
    for ( int i = 1; i < n; i++ ) {  // insertion sort on the partSource index (low 32 bits)
      const uint64_t val = v[i];
      int            k   = i - 1;
      while ( k >= 0 && (uint32_t)v[k] > (uint32_t)val ) {
        v[k + 1] = v[k];
        k--;
      }
      v[k + 1] = val;
    }
    if ( !std_sort_emulated( v, n ) ) { atomicOr( a.err, 1u ); }
    double c2[3] = {0.0, 0.0, 0.0};
    if ( n == 1 ) {  // :1342-1348
      const uint32_t p    = (uint32_t)v[0];
      const int64_t  base = a.frame_off[frame_of( a.frame_off, a.F, a.moved[p / KF] )];
      const ushort4  c    = a.col[base + a.part[p]];
      c2[0] = c.x, c2[1] = c.y, c2[2] = c.z;
    } else {  // maxColorDist2 <= DBL_MAX: weighted mean with 1 / (sqrt(d2) + 4) (:1364-1372)
      double sw = 0.0;
      for ( int i = 0; i < n; i++ ) {
        const uint32_t p    = (uint32_t)v[i];
        const int64_t  base = a.frame_off[frame_of( a.frame_off, a.F, a.moved[p / KF] )];
        const ushort4  c    = a.col[base + a.part[p]];
        const double   w    = 1.0 / ( sqrt( (double)( v[i] >> 32 ) ) + 4.0 );
        c2[0]               = __dadd_rn( c2[0], __dmul_rn( (double)c.x, w ) );
        c2[1]               = __dadd_rn( c2[1], __dmul_rn( (double)c.y, w ) );
        c2[2]               = __dadd_rn( c2[2], __dmul_rn( (double)c.z, w ) );
        sw                  = __dadd_rn( sw, w );
      }
      c2[0] /= sw, c2[1] /= sw, c2[2] /= sw;
    }
    // fixWeight: w = 0 -> color0 = clip( round( 0 * centroid1 + 1 * centroid2 ) ); searchRange 0 keeps color0 (:1413-1466)
    const ushort4 c1 = a.refined1[m];
    out.x = (unsigned short)fmin( fmax( round( __dadd_rn( __dmul_rn( 0.0, (double)c1.x ), __dmul_rn( 1.0, c2[0] ) ) ), 0.0 ), 65535.0 );
    out.y = (unsigned short)fmin( fmax( round( __dadd_rn( __dmul_rn( 0.0, (double)c1.y ), __dmul_rn( 1.0, c2[1] ) ) ), 0.0 ), 65535.0 );
    out.z = (unsigned short)fmin( fmax( round( __dadd_rn( __dmul_rn( 0.0, (double)c1.z ), __dmul_rn( 1.0, c2[2] ) ) ), 0.0 ), 65535.0 );
  }
  (void)f;
  a.refined1[m] = out;  // written to the cloud by k_transfer_store once every target has read the old colours
}

__global__ void k_transfer_store( const TArgs a ) {
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if ( m >= a.nMoved ) { return; }
  a.col[a.moved[m]] = a.refined1[m];
}

// PCCKdTree::search for a batch of queries (test / integration entry: pins the emulation against the reference)
template <int K>
__global__ void k_knn_queries( const KdForest f, const int16_t* __restrict__ q, int64_t nq, int64_t* __restrict__ outIdx,
                               double* __restrict__ outDist ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= nq ) { return; }
  const int   qq[3] = {q[3 * i] - f.ox, q[3 * i + 1] - f.oy, q[3 * i + 2] - f.oz};
  KdResult<K> res;
  kd_search<K>( f, 1u, qq, res );
#pragma unroll
  for ( int j = 0; j < K; j++ ) {
    outIdx[i * K + j]  = j < res.count ? (int64_t)res.idx[j] : -1;
    outDist[i * K + j] = j < res.count ? (double)res.dist[j] : -1.0;
  }
}

__global__ void k_unpack_positions( const int16_t* __restrict__ in, int64_t n, short4* __restrict__ out ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  out[i] = make_short4( in[3 * i], in[3 * i + 1], in[3 * i + 2], 0 );
}

template <int K>
int launch_knn( rb200_ctx* c, const KdForest& f, const int16_t* q, int64_t nq, int64_t* oi, double* od ) {
  RB_LAUNCH( "kd_knn", k_knn_queries<K>, rb_div_up( nq, 128 ), 128, 0, f, q, nq, oi, od );
  return RB200_OK;
}

// ---- singleMapPixelInterleaving / pointLocalReconstruction: colours of the interpolated and fill points ----
// colorPointCloud (PCCCodec.cpp:1367-1374, :1429-1434): the coded point of every pixel (layer == checkerboard parity)
// reads the attribute frame and joins `source`; every other point joins `target` and gets
// PCCPointSet3::transferColorWeight (PCCPointSet.cpp:2250-2280): 5-NN in a kd-tree over `source`, the colour of an
// identical / single neighbour, else the 1/(d^2)^2-weighted mean in double, truncated to uint16.
constexpr int KW = 5;

__global__ void k_ilv_flag( const ushort4* __restrict__ col, const uint32_t* __restrict__ pix, int64_t n, int plr,
                            uint32_t* __restrict__ flags ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i > n ) { return; }
  uint32_t s = 0;
  if ( i < n ) {
    const uint32_t pv = pix[i];
    // interleaving: the layer equals the checkerboard parity (:1367-1369); PLR: f < mapCount, i.e. layer 0 (:1418)
    s = col[i].w == ( plr ? 0u : ( ( ( pv & 0xFFFFu ) + ( pv >> 16 ) ) & 1u ) ) ? 1u : 0u;
  }
  flags[i] = s;
}
__global__ void k_ilv_compact( const uint32_t* __restrict__ scan, int64_t n, const short4* __restrict__ pos,
                               short4* __restrict__ srcPos, uint32_t* __restrict__ srcIdx ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n || scan[i + 1] == scan[i] ) { return; }
  short4 p        = pos[i];
  p.w             = 0;
  srcPos[scan[i]] = p;
  srcIdx[scan[i]] = (uint32_t)i;
}
__global__ void __launch_bounds__( 128 ) k_ilv_transfer( const KdForest forest, int F, const int64_t* __restrict__ frame_off,
                                                         const uint32_t* __restrict__ scan, const int32_t* __restrict__ rootOf,
                                                         const uint32_t* __restrict__ srcIdx, const short4* __restrict__ pos,
                                                         ushort4* __restrict__ col, int64_t n ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n || scan[i + 1] != scan[i] ) { return; }  // source points keep their colour
  const int f    = frame_of( frame_off, F, i );
  const int root = rootOf[f];
  if ( root < 0 ) { return; }  // no source: transferColorWeight returns early, the colour stays 0 (:2253)
  const uint32_t  sb   = scan[frame_off[f]];  // first compacted source point of the frame
  const short4    p    = pos[i];
  const int       q[3] = {p.x - forest.ox, p.y - forest.oy, p.z - forest.oz};
  KdResult<KW>    res;
  kd_search<KW>( forest, (uint32_t)root, q, res );
  ushort4 out = col[srcIdx[sb + res.idx[0]]];
  if ( res.count > 1 && res.dist[0] != 0 ) {  // result.size() > 1 && result.dist( 0 ) > 0.0001 (:2262)
    double rc[3] = {0.0, 0.0, 0.0}, sw = 0.0;
#pragma unroll
    for ( int j = 0; j < KW; j++ ) {
      if ( j < res.count ) {
        const double  d = (double)res.dist[j];
        const double  w = __ddiv_rn( 1.0, __dmul_rn( d, d ) );  // 1.0 / pow( dist, 2.0 ), dist = squared distance
        const ushort4 c = col[srcIdx[sb + res.idx[j]]];
        rc[0]           = __dadd_rn( rc[0], __dmul_rn( (double)c.x, w ) );
        rc[1]           = __dadd_rn( rc[1], __dmul_rn( (double)c.y, w ) );
        rc[2]           = __dadd_rn( rc[2], __dmul_rn( (double)c.z, w ) );
        sw              = __dadd_rn( sw, w );
      }
    }
    out.x = (unsigned short)(int)__ddiv_rn( rc[0], sw );  // PCCVector3D -> PCCColor16bit: (uint16_t) truncation
    out.y = (unsigned short)(int)__ddiv_rn( rc[1], sw );
    out.z = (unsigned short)(int)__ddiv_rn( rc[2], sw );
  }
  out.w  = col[i].w;  // the pointToPixel layer travels in .w
  col[i] = out;
}


// ---------------------------------------------------------------------------------------------------
// PCCCodec::smoothPointCloud (PCCCodec.cpp:1106-1157), the non-grid geometry filter.  Per point: PCCKdTree::searchRadius
// (PCCKdTree.cpp:69-79) = nanoflann radiusSearch — every point with dist < radius2Smoothing in TRAVERSAL order
// (RadiusResultSet::addPoint, worstDist() == radius, nanoflann.hpp:945-952, 1207-1253), std::sort with the vendored
// IndexDist_Sorter (by distance, EQUAL DISTANCES BY INDEX, :193-200 — a total order, unlike upstream nanoflann's) and
// a cut to neighborCountSmoothing entries.  On the integer lattice the cut nearly always falls inside a shell of equal
// distances: the lower indices of the shell survive.
// ---------------------------------------------------------------------------------------------------
struct RadiusArgs {
  KdForest        forest;
  int             F;
  const int64_t*  frame_off;
  const short4*   posIn;   // the reconstruction (kept copy)
  short4*         posOut;  // smoothed positions, boundary type 1 -> 2 (:1132-1134)
  const uint32_t* part;
  int64_t         N;
  uint32_t        dLim;    // dist < radius2Smoothing        <=>  d < dLim   (d is an integer)
  uint32_t        visit;   // mindistsq <= radius2Smoothing  <=>  m <= visit
  uint32_t        bLim;    // dist2 <= radius2BoundaryDetection <=> d <= bLim
  int             maxCount;
  double          threshold;
  uint64_t*       lists;   // [threads][cap] dist << 32 | index in frame
  int             cap;
  uint32_t*       err;     // 1: list overflow
};

// ascending sort of dist << 32 | index keys: the reference's IndexDist_Sorter orders equal distances by index
// (nanoflann.hpp:193-200 as vendored), a total order, so any sorting algorithm gives std::sort's result
__device__ void sort_keys( uint64_t* v, int n ) {
  for ( int start = n / 2 - 1; start >= 0; start-- ) {  // heapsort: in place, no recursion, O(n log n) on any input
    int            root = start;
    const uint64_t val  = v[root];
    for ( int child = 2 * root + 1; child < n; child = 2 * root + 1 ) {
      if ( child + 1 < n && v[child + 1] > v[child] ) { child++; }
      if ( v[child] <= val ) { break; }
      v[root] = v[child];
      root    = child;
    }
    v[root] = val;
  }
  for ( int end = n - 1; end > 0; end-- ) {
    const uint64_t val = v[end];
    v[end]             = v[0];
    int root           = 0;
    for ( int child = 1; child < end; child = 2 * root + 1 ) {
      if ( child + 1 < end && v[child + 1] > v[child] ) { child++; }
      if ( v[child] <= val ) { break; }
      v[root] = v[child];
      root    = child;
    }
    v[root] = val;
  }
}

// findNeighbors + searchLevel with a RadiusResultSet (the walk of kd_search; the bound is the constant radius): every
// element with d < dLim, in traversal order, as dist << 32 | index into L[0, cap); returns how many there are
__device__ __forceinline__ int kd_radius( const KdForest& f, uint32_t root, const int q[3], uint32_t dLim, uint32_t visit, uint64_t* L, int cap ) {
  int      n  = 0;
  uint32_t d0 = 0, d1 = 0, d2 = 0;
  {
    const int16_t* rb = f.rootBox + (size_t)root * 6;
    int            t;
    t  = q[0] < rb[0] ? q[0] - rb[0] : ( q[0] > rb[3] ? q[0] - rb[3] : 0 );
    d0 = (uint32_t)( t * t );
    t  = q[1] < rb[1] ? q[1] - rb[1] : ( q[1] > rb[4] ? q[1] - rb[4] : 0 );
    d1 = (uint32_t)( t * t );
    t  = q[2] < rb[2] ? q[2] - rb[2] : ( q[2] > rb[5] ? q[2] - rb[5] : 0 );
    d2 = (uint32_t)( t * t );
  }
  uint32_t stN[KD_STACK], stV[KD_STACK];
  int      sp   = 0;
  uint32_t node = root;
  for ( ;; ) {
    if ( node ) {
      const uint4 nv = __ldg( reinterpret_cast<const uint4*>( f.nodes + node ) );
      if ( nv.y & KD_LEAF ) {
        const uint32_t lend = nv.x + ( nv.y & 0xFFFFu );
        for ( uint32_t e = nv.x; e < lend; e++ ) {
          const uint64_t r  = f.rec[e];
          const int      dx = q[0] - kd_coord( r, 0 ), dy = q[1] - kd_coord( r, 1 ), dz = q[2] - kd_coord( r, 2 );
          const uint32_t d  = (uint32_t)( dx * dx ) + (uint32_t)( dy * dy ) + (uint32_t)( dz * dz );
          if ( d < dLim ) {
            if ( n < cap ) { L[n] = ( (uint64_t)d << 32 ) | kd_index( r ); }
            n++;
          }
        }
        node = 0;
      } else {
        const uint32_t axis   = nv.y & 3u;
        const int      val    = axis == 0 ? q[0] : ( axis == 1 ? q[1] : q[2] );
        const int      divlow = (int16_t)( nv.z & 0xFFFFu ), divhigh = (int16_t)( nv.z >> 16 );
        const int      diff1 = val - divlow, diff2 = val - divhigh;
        const bool     nearLeft = diff1 + diff2 < 0;
        const int      cd       = nearLeft ? diff2 : diff1;
        stN[sp]                 = ( nv.x + ( nearLeft ? 1u : 0u ) ) | ( axis << 29 );
        stV[sp]                 = (uint32_t)( cd * cd );
        sp++;
        node = nv.x + ( nearLeft ? 0u : 1u );
      }
    } else {
      if ( sp == 0 ) { break; }
      sp--;
      const uint32_t e = stN[sp], v = stV[sp], axis = ( e >> 29 ) & 3u;
      if ( e >> 31 ) {
        d0 = axis == 0 ? v : d0, d1 = axis == 1 ? v : d1, d2 = axis == 2 ? v : d2;
      } else {
        const uint32_t dst = axis == 0 ? d0 : ( axis == 1 ? d1 : d2 );
        const uint32_t m   = d0 + d1 + d2 + v - dst;
        if ( m <= visit ) {
          stN[sp] = e | 0x80000000u;
          stV[sp] = dst;
          sp++;
          d0 = axis == 0 ? v : d0, d1 = axis == 1 ? v : d1, d2 = axis == 2 ? v : d2;
          node = e & ( KD_NODE_MAX - 1u );
        }
      }
    }
  }
  return n;
}

// PCCKdTree::searchRadius for explicit queries (rb200_kdtree_search_radius): sorted == 0 leaves the traversal order
__global__ void __launch_bounds__( 128 ) k_radius_query( const KdForest forest, const int16_t* __restrict__ queries, int64_t nq, uint32_t dLim,
                                                         uint32_t visit, int maxResults, int sorted, uint64_t* lists, int cap,
                                                         int64_t* __restrict__ outIdx, double* __restrict__ outDist, int32_t* __restrict__ outCount,
                                                         uint32_t* err ) {
  uint64_t* const L = lists + (size_t)( blockIdx.x * blockDim.x + threadIdx.x ) * cap;
  for ( int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x ) {
    const int q[3] = {queries[3 * i] - forest.ox, queries[3 * i + 1] - forest.oy, queries[3 * i + 2] - forest.oz};
    const int n    = kd_radius( forest, 1u, q, dLim, visit, L, cap );
    outCount[i]    = n;
    if ( n > cap ) {
      atomicOr( err, 1u );
      continue;
    }
    if ( sorted ) { sort_keys( L, n ); }
    for ( int r = 0; r < maxResults; r++ ) {
      outIdx[i * maxResults + r]  = r < n ? (int64_t)(uint32_t)L[r] : -1;
      outDist[i * maxResults + r] = r < n ? (double)(uint32_t)( L[r] >> 32 ) : -1.0;
    }
  }
}

__global__ void __launch_bounds__( 128 ) k_smooth_radius( const RadiusArgs a ) {
  uint64_t* const L = a.lists + (size_t)( blockIdx.x * blockDim.x + threadIdx.x ) * a.cap;
  for ( int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.N; i += (int64_t)gridDim.x * blockDim.x ) {
    const int      f    = frame_of( a.frame_off, a.F, i );
    const int64_t  base = a.frame_off[f];
    const short4   p    = a.posIn[i];
    const int      q[3] = {p.x - a.forest.ox, p.y - a.forest.oy, p.z - a.forest.oz};
    const int      n    = kd_radius( a.forest, (uint32_t)f + 1u, q, a.dLim, a.visit, L, a.cap );
    if ( n > a.cap ) { atomicOr( a.err, 1u ); }
    short4 out = p;
    if ( n <= a.cap ) {
      sort_keys( L, n );  // searchParams.sorted (:949-950): by distance, equal distances by index
      const int      cnt  = min( n, a.maxCount );                   // ret.resize( retSize ), PCCKdTree.cpp:75-76
      const uint32_t mine = a.part[i];
      long long      sx = 0, sy = 0, sz = 0;
      bool           other = false;
      for ( int r = 0; r < cnt; r++ ) {  // :1122-1129 (sums of int16 coordinates in double: exact)
        const uint32_t j  = (uint32_t)L[r], d = (uint32_t)( L[r] >> 32 );
        const short4   pj = a.posIn[base + j];
        sx += pj.x, sy += pj.y, sz += pj.z;
        other |= d <= a.bLim && a.part[base + j] != mine;
      }
      if ( other ) {
        if ( p.w == 1 ) { out.w = 2; }  // :1131-1133
        const double dn = (double)cnt;
        const double ex = __dsub_rn( (double)sx, __dmul_rn( dn, (double)p.x ) ), ey = __dsub_rn( (double)sy, __dmul_rn( dn, (double)p.y ) ),
                     ez = __dsub_rn( (double)sz, __dmul_rn( dn, (double)p.z ) );
        const double norm2 = __dadd_rn( __dadd_rn( __dmul_rn( ex, ex ), __dmul_rn( ey, ey ) ), __dmul_rn( ez, ez ) );
        const double dist  = (double)(long long)( __dadd_rn( norm2, dn / 2.0 ) ) / dn;  // :1136-1137
        if ( dist >= a.threshold ) {                                                   // :1141
          const double half = (double)( cnt / 2 );                                     // ( neighborCount / 2 ): size_t division
          out.x = (short)(double)(long long)( __dadd_rn( (double)sx, half ) / dn );   // :1138-1140, then PCCPoint3D( centroid )
          out.y = (short)(double)(long long)( __dadd_rn( (double)sy, half ) / dn );
          out.z = (short)(double)(long long)( __dadd_rn( (double)sz, half ) / dn );
        }
      }
    }
    a.posOut[i] = out;
  }
}

}  // namespace

int rb_interleave_colors_impl( rb200_ctx* c ) {
  const rb200_params& P = c->P;
  const int           F = c->F;
  const int64_t       N = c->h_frame_off[F];
  if ( N == 0 || P.attribute_count == 0 ) { return RB200_OK; }
  if ( P.geometry_bitdepth_3d > 12 ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "pixel interleaving: geometry bit depth above 12 is not supported by the kd-tree emulation" );
  }
  TransferScratch* S = scratch_of( c );
  RB_CUDA( S->flags.ensure( (size_t)( N + 8 ) * 4 ) );
  RB_CUDA( S->sums.ensure( rb_scan_scratch_bytes( N + 1 ) ) );
  uint32_t* flags = S->flags.as<uint32_t>();
  RB_LAUNCH( "ilv_flag", k_ilv_flag, rb_div_up( N + 1, TPB ), TPB, 0, c->d_col.as<ushort4>(), c->d_pix.as<uint32_t>(), N,
             ( P.point_local_reconstruction && !P.single_map_pixel_interleaving ) ? 1 : 0, flags );
  int r = rb_scan_u32( c, flags, flags, N + 1, S->sums.as<uint32_t>() );
  if ( r ) { return r; }
  // compacted index of the first source point of every frame (and the total)
  uint32_t* h = (uint32_t*)rb_pinned( c, (size_t)( F + 1 ) * 4 + (size_t)F * 4 + ( F + 2 ) * 8 );
  if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  for ( int f = 0; f <= F; f++ ) {
    RB_CUDA( cudaMemcpyAsync( h + f, flags + c->h_frame_off[f], 4, cudaMemcpyDeviceToHost, c->stream ) );
  }
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  const uint32_t nSrc = h[F];
  if ( nSrc == 0 || (int64_t)nSrc == N ) { return RB200_OK; }
  std::vector<int64_t> hOff{0};
  int32_t*             hRoot = (int32_t*)( h + F + 1 );
  for ( int f = 0; f < F; f++ ) {
    const int64_t ns = (int64_t)h[f + 1] - h[f], nt = ( c->h_frame_off[f + 1] - c->h_frame_off[f] ) - ns;
    hRoot[f]         = -1;
    if ( ns == 0 || nt == 0 ) { continue; }
    if ( ns < KW ) {  // PCCKdTree::search leaves result.size() at 5 whatever nanoflann found (PCCKdTree.cpp:62-67)
      return rb_fail( c, RB200_ERR_UNSUPPORTED, "pixel interleaving: frame %d has fewer than %d coded points", f, KW );
    }
    hRoot[f] = (int32_t)hOff.size();  // tree t has root t + 1
    hOff.push_back( h[f + 1] );
    if ( hOff[hOff.size() - 2] != h[f] ) {  // a frame without targets in between: its source points own no tree
      return rb_fail( c, RB200_ERR_UNSUPPORTED, "pixel interleaving: a frame without interpolated points inside the GOF is not supported" );
    }
  }
  if ( hOff.size() == 1 ) { return RB200_OK; }
  RB_CUDA( S->pos2.ensure( (size_t)nSrc * 8 ) );
  RB_CUDA( S->moved.ensure( (size_t)nSrc * 4 ) );
  RB_CUDA( S->off.ensure( hOff.size() * 8 ) );
  RB_CUDA( S->small.ensure( std::max<size_t>( 64, (size_t)F * 4 ) ) );
  {
    int64_t* hp = (int64_t*)( h + 2 * F + 2 );  // 8-byte aligned behind h[F + 1] and hRoot[F]
    memcpy( hp, hOff.data(), hOff.size() * 8 );
    RB_CUDA( cudaMemcpyAsync( S->off.p, hp, hOff.size() * 8, cudaMemcpyHostToDevice, c->stream ) );
    RB_CUDA( cudaMemcpyAsync( S->small.p, hRoot, (size_t)F * 4, cudaMemcpyHostToDevice, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
  }
  RB_LAUNCH( "ilv_compact", k_ilv_compact, rb_div_up( N, TPB ), TPB, 0, flags, N, c->d_pos.as<short4>(), S->pos2.as<short4>(),
             S->moved.as<uint32_t>() );
  r = rb_kd_build( c, S->kd, S->pos2.as<short4>(), S->off.as<int64_t>(), hOff, 0, 0, 0 );
  if ( r ) { return r; }
  RB_LAUNCH( "ilv_transfer", k_ilv_transfer, rb_div_up( N, 128 ), 128, 0, S->kd.forest, F, c->d_frame_off.as<int64_t>(), flags,
             S->small.as<int32_t>(), S->moved.as<uint32_t>(), c->d_pos.as<short4>(), c->d_col.as<ushort4>(), N );
  return RB200_OK;
}

// smoothPointCloud for every frame of the GOF: d_pos_pre keeps the reconstruction, d_pos receives the result
int rb_smooth_radius_impl( rb200_ctx* c ) {
  const rb200_params& P = c->P;
  const int           F = c->F;
  const int64_t       N = c->h_frame_off[F];
  if ( N == 0 ) { return RB200_OK; }
  if ( P.neighbor_count_smoothing < 1 || !( P.radius2_smoothing > 0.0 ) || !( P.radius2_boundary_detection >= 0.0 ) ) {
    return rb_fail( c, RB200_ERR_INVALID, "non-grid smoothing needs neighbor_count_smoothing >= 1 and radius2_smoothing > 0" );
  }
  if ( P.radius2_smoothing > 4096.0 ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "non-grid smoothing: radius2_smoothing above 4096" ); }
  if ( P.geometry_bitdepth_3d > 12 ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "non-grid smoothing: geometry bit depth above 12 is not supported by the kd-tree emulation" );
  }
  for ( int f = 0; f < F; f++ ) {
    if ( c->h_frame_off[f + 1] == c->h_frame_off[f] ) {
      return rb_fail( c, RB200_ERR_UNSUPPORTED, "non-grid smoothing: a GOF with an empty frame is not supported" );
    }
  }
  TransferScratch* S = scratch_of( c );
  RB_CUDA( c->d_pos_pre.ensure( (size_t)N * 8 ) );
  RB_CUDA( cudaMemcpyAsync( c->d_pos_pre.p, c->d_pos.p, (size_t)N * 8, cudaMemcpyDeviceToDevice, c->stream ) );
  c->pos_pre_valid = true;
  std::vector<int64_t> hOff( c->h_frame_off.begin(), c->h_frame_off.begin() + F + 1 );
  RB_CUDA( S->off.ensure( hOff.size() * 8 ) );
  {
    int64_t* hp = (int64_t*)rb_pinned( c, hOff.size() * 8 );
    if ( !hp ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
    memcpy( hp, hOff.data(), hOff.size() * 8 );
    RB_CUDA( cudaMemcpyAsync( S->off.p, hp, hOff.size() * 8, cudaMemcpyHostToDevice, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
  }
  int r = rb_kd_build( c, S->kd, c->d_pos_pre.as<short4>(), S->off.as<int64_t>(), hOff, 0, 0, 0 );
  if ( r ) { return r; }
  // every lattice point of the ball twice over; a denser neighbourhood (heavy duplication) fails loudly
  const double   r2      = P.radius2_smoothing;
  const double   rad     = sqrt( r2 ) + 1.0;
  const int      cap     = std::max( 256, 2 * (int)( 4.19 * rad * rad * rad ) );
  const int      threads = (int)std::min<int64_t>( 148 * 2 * 128, ( ( N + 127 ) / 128 ) * 128 );  // (49 KB of list per thread at r2 = 64)
  RB_CUDA( S->candKey.ensure( (size_t)threads * cap * 8 ) );
  RB_CUDA( S->small.ensure( 64 ) );
  RB_CUDA( cudaMemsetAsync( S->small.p, 0, 64, c->stream ) );
  RadiusArgs a{};
  a.forest    = S->kd.forest;
  a.F         = F;
  a.frame_off = c->d_frame_off.as<int64_t>();
  a.posIn     = c->d_pos_pre.as<short4>();
  a.posOut    = c->d_pos.as<short4>();
  a.part      = c->d_part.as<uint32_t>();
  a.N         = N;
  a.dLim      = (uint32_t)ceil( r2 );   // integer d < r2  <=>  d < ceil( r2 )
  a.visit     = (uint32_t)floor( r2 );  // integer m <= r2 <=>  m <= floor( r2 )
  a.bLim      = (uint32_t)std::min( floor( P.radius2_boundary_detection ), 4.0e9 );
  a.maxCount  = P.neighbor_count_smoothing;
  a.threshold = P.threshold_smoothing;
  a.lists     = S->candKey.as<uint64_t>();
  a.cap       = cap;
  a.err       = S->small.as<uint32_t>();
  RB_LAUNCH( "geo_radius", k_smooth_radius, threads / 128, 128, 0, a );
  uint32_t* h = (uint32_t*)rb_pinned( c, 64 );
  if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  RB_CUDA( cudaMemcpyAsync( h, S->small.p, 4, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  if ( h[0] ) {
    // nothing usable was written for the failing points: put the reconstruction back
    RB_CUDA( cudaMemcpyAsync( c->d_pos.p, c->d_pos_pre.p, (size_t)N * 8, cudaMemcpyDeviceToDevice, c->stream ) );
    return rb_fail( c, RB200_ERR_UNSUPPORTED,
                    "non-grid smoothing: more than %d points inside the smoothing radius of one point", cap );
  }
  return RB200_OK;
}

void rb_transfer_release( rb200_ctx* c ) {
  TransferScratch* s = static_cast<TransferScratch*>( c->transfer_scratch );
  if ( !s ) { return; }
  s->kd.release();
  RbBuf* b[] = {&s->pos2, &s->off, &s->flags, &s->sums, &s->moved, &s->part, &s->partDist, &s->bwdT, &s->partCnt, &s->refined1,
                &s->candCnt, &s->candOff, &s->candKey, &s->small, &s->claim, &s->srcList, &s->srcNN};
  for ( auto* x : b ) { x->release(); }
  delete s;
  c->transfer_scratch = nullptr;
}

int rb_transfer_colors_impl( rb200_ctx* c ) {
  const rb200_params& P = c->P;
  const int           F = c->F;
  const int64_t       N = c->h_frame_off[F];
  if ( N == 0 || P.attribute_count == 0 ) { return RB200_OK; }
  if ( !c->pos_pre_valid || !c->d_pos_pre.p || c->d_pos_pre.cap < (size_t)N * 8 ) {
    return rb_fail( c, RB200_ERR_STATE, "transfer_colors: the pre-smoothing cloud was not kept (attr_transfer_filter_type != 1?)" );
  }
  if ( P.geometry_bitdepth_3d > 12 ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "transfer_colors: geometry bit depth above 12 is not supported by the kd-tree emulation" );
  }
  for ( int f = 0; f < F; f++ ) {
    const int64_t n = c->h_frame_off[f + 1] - c->h_frame_off[f];
    if ( n > 0 && n < KF ) {
      return rb_fail( c, RB200_ERR_UNSUPPORTED, "transfer_colors: frame %d has fewer than %d points", f, KF );
    }
  }
  TransferScratch* S = scratch_of( c );
  // ---- moved (type 3) points ----
  const int64_t nWords = ( N + 31 ) / 32;
  if ( !c->d_moved_bits.p || c->d_moved_bits.cap < (size_t)nWords * 4 ) {
    return rb_fail( c, RB200_ERR_STATE, "transfer_colors: no record of the moved points (smooth_geometry of this reconstruction did not run)" );
  }
  RB_CUDA( S->flags.ensure( (size_t)( N + 8 ) * 4 ) );  // rank of the moved points (written at their indices only)
  RB_CUDA( S->candCnt.ensure( (size_t)( nWords + 8 ) * 4 ) );
  RB_CUDA( S->sums.ensure( rb_scan_scratch_bytes( std::max<int64_t>( N + 1, nWords + 1 ) ) ) );
  uint32_t*       flags = S->flags.as<uint32_t>();
  uint32_t*       wcnt  = S->candCnt.as<uint32_t>();
  const uint32_t* bits  = c->d_moved_bits.as<uint32_t>();
  RB_LAUNCH( "tr_moved_count", k_moved_count, rb_div_up( nWords + 1, TPB ), TPB, 0, bits, nWords, wcnt );
  int r = rb_scan_u32( c, wcnt, wcnt, nWords + 1, S->sums.as<uint32_t>() );
  if ( r ) { return r; }
  uint32_t* h = (uint32_t*)rb_pinned( c, 64 );
  if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  RB_CUDA( cudaMemcpyAsync( h, wcnt + nWords, 4, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  const uint32_t M = h[0];
  if ( M == 0 ) { return RB200_OK; }  // nothing moved: every colour stays (:1163-1164 with type != 3)
  RB_CUDA( S->moved.ensure( (size_t)M * 4 ) );
  RB_LAUNCH( "tr_list_moved", k_moved_list, rb_div_up( nWords, TPB ), TPB, 0, bits, wcnt, nWords, S->moved.as<uint32_t>(), flags );
  // ---- forest: trees 1..F over the pre-smoothing frames, F+1..2F over the smoothed frames ----
  std::vector<int64_t> hOff;
  std::vector<int>     treeOfS( F, -1 ), treeOfT( F, -1 );
  hOff.push_back( 0 );
  // empty frames own no tree; the roots of the others are numbered consecutively, so keep a frame -> root map
  for ( int pass = 0; pass < 2; pass++ ) {
    for ( int f = 0; f < F; f++ ) {
      const int64_t n = c->h_frame_off[f + 1] - c->h_frame_off[f];
      if ( n == 0 ) { continue; }
      ( pass == 0 ? treeOfS : treeOfT )[f] = (int)hOff.size() - 1;
      hOff.push_back( hOff.back() + n );
    }
  }
  bool dense = true;  // no empty frame: tree ids are f and F + f (the kernels assume this layout)
  for ( int f = 0; f < F; f++ ) { dense &= ( treeOfS[f] == f && treeOfT[f] == F + f ); }
  if ( !dense ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "transfer_colors: a GOF with an empty frame next to smoothed frames is not supported" );
  }
  RB_CUDA( S->pos2.ensure( (size_t)2 * N * 8 ) );
  RB_CUDA( cudaMemcpyAsync( S->pos2.p, c->d_pos_pre.p, (size_t)N * 8, cudaMemcpyDeviceToDevice, c->stream ) );
  RB_CUDA( cudaMemcpyAsync( S->pos2.as<char>() + (size_t)N * 8, c->d_pos.p, (size_t)N * 8, cudaMemcpyDeviceToDevice, c->stream ) );
  RB_CUDA( S->off.ensure( hOff.size() * 8 ) );
  {
    int64_t* hp = (int64_t*)rb_pinned( c, hOff.size() * 8 );
    if ( !hp ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
    memcpy( hp, hOff.data(), hOff.size() * 8 );
    RB_CUDA( cudaMemcpyAsync( S->off.p, hp, hOff.size() * 8, cudaMemcpyHostToDevice, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
  }
  r = rb_kd_build( c, S->kd, S->pos2.as<short4>(), S->off.as<int64_t>(), hOff, 0, 0, 0 );
  if ( r ) { return r; }
  // ---- transfer ----
  RB_CUDA( S->part.ensure( (size_t)M * KF * 4 ) );
  RB_CUDA( S->partDist.ensure( (size_t)M * KF * 4 ) );
  RB_CUDA( S->bwdT.ensure( (size_t)M * KF * 4 ) );
  RB_CUDA( S->partCnt.ensure( (size_t)M ) );
  RB_CUDA( S->refined1.ensure( (size_t)M * 8 ) );
  RB_CUDA( S->candCnt.ensure( (size_t)( M + 8 ) * 4 ) );
  RB_CUDA( S->candOff.ensure( (size_t)( M + 8 ) * 4 ) );
  RB_CUDA( S->candKey.ensure( (size_t)M * KF * 8 ) );
  RB_CUDA( S->small.ensure( 64 ) );
  RB_CUDA( cudaMemsetAsync( S->small.p, 0, 64, c->stream ) );
  RB_CUDA( S->claim.ensure( (size_t)( N / 32 + 2 ) * 4 ) );
  RB_CUDA( S->srcList.ensure( (size_t)M * KF * 4 ) );
  RB_CUDA( S->srcNN.ensure( (size_t)N * 8 ) );
  RB_CUDA( cudaMemsetAsync( S->claim.p, 0, (size_t)( N / 32 + 2 ) * 4, c->stream ) );
  TArgs a{};
  a.forest    = S->kd.forest;
  a.F         = F;
  a.frame_off = c->d_frame_off.as<int64_t>();
  a.posS      = c->d_pos_pre.as<short4>();
  a.posT      = c->d_pos.as<short4>();
  a.col       = c->d_col.as<ushort4>();
  a.moved     = S->moved.as<uint32_t>();
  a.rank      = flags;
  a.nMoved    = M;
  a.part      = S->part.as<uint32_t>();
  a.partDist  = S->partDist.as<uint32_t>();
  a.bwdT      = S->bwdT.as<uint32_t>();
  a.partCnt   = S->partCnt.as<uint8_t>();
  a.refined1  = S->refined1.as<ushort4>();
  a.candCnt   = S->candCnt.as<uint32_t>();
  a.candOff   = S->candOff.as<uint32_t>();
  a.candKey   = S->candKey.as<uint64_t>();
  a.lossless  = P.attribute_rgb444;
  a.err       = S->small.as<uint32_t>();
  a.claim     = S->claim.as<uint32_t>();
  a.srcList   = S->srcList.as<uint32_t>();
  a.nSrc      = S->small.as<uint32_t>() + 4;
  a.srcNN     = S->srcNN.as<uint2>();
  RB_LAUNCH( "tr_forward", k_transfer_fwd, rb_div_up( M, 128 ), 128, 0, a );
  RB_CUDA( cudaMemsetAsync( a.candCnt, 0, (size_t)( M + 1 ) * 4, c->stream ) );
  RB_LAUNCH( "tr_backward_claim", k_transfer_bwd_claim, rb_div_up( (int64_t)M * KF, 256 ), 256, 0, a );
  // (the number of distinct source points stays on the device: the grid covers the worst case, surplus CTAs leave)
  RB_LAUNCH( "tr_backward_search", k_transfer_bwd_search, rb_div_up( (int64_t)M * KF, 128 ), 128, 0, a );
  RB_LAUNCH( "tr_backward", k_transfer_bwd, rb_div_up( (int64_t)M * KF, 256 ), 256, 0, a );
  r = rb_scan_u32( c, a.candCnt, S->candOff.as<uint32_t>(), M + 1, S->sums.as<uint32_t>() );
  if ( r ) { return r; }
  RB_CUDA( cudaMemsetAsync( a.candCnt, 0, (size_t)( M + 1 ) * 4, c->stream ) );
  RB_LAUNCH( "tr_backward_store", k_transfer_bwd_store, rb_div_up( (int64_t)M * KF, 256 ), 256, 0, a );
  RB_LAUNCH( "tr_final", k_transfer_final, rb_div_up( M, 128 ), 128, 0, a );
  RB_LAUNCH( "tr_store", k_transfer_store, rb_div_up( M, 128 ), 128, 0, a );
  RB_CUDA( cudaMemcpyAsync( h, a.err, 4, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  if ( h[0] ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "transfer_colors: a candidate list drove std::sort into its heapsort fallback "
                                              "(not reproduced)" );
  }
  return RB200_OK;
}

extern "C" int rb200_kdtree_search( rb200_ctx* c, const int16_t* cloud, int64_t n, const int16_t* queries, int64_t nq, int k,
                                    int64_t* outIdx, double* outDist ) {
  if ( !c || !cloud || !queries || !outIdx || !outDist || n <= 0 || nq < 0 ) {
    return rb_fail( c, RB200_ERR_INVALID, "kdtree_search: bad arguments" );
  }
  if ( k < 1 || k > 8 ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "kdtree_search: k must be 1..8" ); }
  cudaSetDevice( c->device );
  if ( nq == 0 ) { return RB200_OK; }
  TransferScratch* S = scratch_of( c );
  // origin = per-axis minimum over cloud and queries (host side: the arrays are the caller's host or device memory)
  RB_CUDA( S->pos2.ensure( (size_t)n * 8 + (size_t)( n + nq ) * 6 + (size_t)nq * k * 16 + 256 ) );
  char*    base = S->pos2.as<char>();
  short4*  dPos = (short4*)base;
  int16_t* dRaw = (int16_t*)( base + (size_t)n * 8 );
  int16_t* dQ   = dRaw + 3 * n;
  char*    dOut = (char*)( ( (uintptr_t)( dQ + 3 * nq ) + 15 ) & ~(uintptr_t)15 );
  int64_t* dIdx = (int64_t*)dOut;
  double*  dDst = (double*)( dOut + (size_t)nq * k * 8 );
  RB_CUDA( cudaMemcpyAsync( dRaw, cloud, (size_t)n * 6, cudaMemcpyDefault, c->stream ) );
  RB_CUDA( cudaMemcpyAsync( dQ, queries, (size_t)nq * 6, cudaMemcpyDefault, c->stream ) );
  RB_LAUNCH( "kd_unpack", k_unpack_positions, rb_div_up( n, TPB ), TPB, 0, dRaw, n, dPos );
  // the origin must make every coordinate of the cloud non-negative: take it from a host-side pass when the cloud
  // is host memory, else assume 0 (decoded clouds are non-negative)
  int ox = 0, oy = 0, oz = 0;
  cudaPointerAttributes attr{};
  if ( cudaPointerGetAttributes( &attr, cloud ) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered ||
       attr.type == cudaMemoryTypeHost ) {
    cudaGetLastError();
    ox = oy = oz = 32767;
    for ( int64_t i = 0; i < n; i++ ) {
      ox = std::min<int>( ox, cloud[3 * i] );
      oy = std::min<int>( oy, cloud[3 * i + 1] );
      oz = std::min<int>( oz, cloud[3 * i + 2] );
    }
  }
  std::vector<int64_t> hOff{0, n};
  RB_CUDA( S->off.ensure( 16 ) );
  {
    int64_t* hp = (int64_t*)rb_pinned( c, 16 );
    if ( !hp ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
    hp[0] = 0, hp[1] = n;
    RB_CUDA( cudaMemcpyAsync( S->off.p, hp, 16, cudaMemcpyHostToDevice, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
  }
  int r = rb_kd_build( c, S->kd, dPos, S->off.as<int64_t>(), hOff, ox, oy, oz );
  if ( r ) { return r; }
  switch ( k ) {
    case 1: r = launch_knn<1>( c, S->kd.forest, dQ, nq, dIdx, dDst ); break;
    case 2: r = launch_knn<2>( c, S->kd.forest, dQ, nq, dIdx, dDst ); break;
    case 3: r = launch_knn<3>( c, S->kd.forest, dQ, nq, dIdx, dDst ); break;
    case 4: r = launch_knn<4>( c, S->kd.forest, dQ, nq, dIdx, dDst ); break;
    case 5: r = launch_knn<5>( c, S->kd.forest, dQ, nq, dIdx, dDst ); break;
    case 6: r = launch_knn<6>( c, S->kd.forest, dQ, nq, dIdx, dDst ); break;
    case 7: r = launch_knn<7>( c, S->kd.forest, dQ, nq, dIdx, dDst ); break;
    default: r = launch_knn<8>( c, S->kd.forest, dQ, nq, dIdx, dDst ); break;
  }
  if ( r ) { return r; }
  RB_CUDA( cudaMemcpyAsync( outIdx, dIdx, (size_t)nq * k * 8, cudaMemcpyDefault, c->stream ) );
  RB_CUDA( cudaMemcpyAsync( outDist, dDst, (size_t)nq * k * 8, cudaMemcpyDefault, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  return RB200_OK;
}

// PCCKdTree::searchRadius (PCCKdTree.cpp:69-79) for explicit queries; sorted == 0 returns nanoflann's traversal order
// (radiusSearch with SearchParams::sorted == false) — the building block of the non-grid smoothPointCloud
extern "C" int rb200_kdtree_search_radius( rb200_ctx* c, const int16_t* cloud, int64_t n, const int16_t* queries, int64_t nq, double radius2,
                                           int maxResults, int sorted, int64_t* outIdx, double* outDist, int32_t* outCount ) {
  if ( !c || !cloud || !queries || !outIdx || !outDist || !outCount || n <= 0 || nq < 0 || maxResults < 1 || !( radius2 > 0.0 ) ||
       radius2 > 4096.0 ) {
    return rb_fail( c, RB200_ERR_INVALID, "kdtree_search_radius: bad arguments" );
  }
  cudaSetDevice( c->device );
  if ( nq == 0 ) { return RB200_OK; }
  TransferScratch* S = scratch_of( c );
  const size_t outB = (size_t)nq * maxResults * 16 + (size_t)nq * 4;
  RB_CUDA( S->pos2.ensure( (size_t)n * 8 + (size_t)( n + nq ) * 6 + outB + 256 ) );
  char*    base = S->pos2.as<char>();
  short4*  dPos = (short4*)base;
  int16_t* dRaw = (int16_t*)( base + (size_t)n * 8 );
  int16_t* dQ   = dRaw + 3 * n;
  char*    dOut = (char*)( ( (uintptr_t)( dQ + 3 * nq ) + 15 ) & ~(uintptr_t)15 );
  int64_t* dIdx = (int64_t*)dOut;
  double*  dDst = (double*)( dOut + (size_t)nq * maxResults * 8 );
  int32_t* dCnt = (int32_t*)( dOut + (size_t)nq * maxResults * 16 );
  RB_CUDA( cudaMemcpyAsync( dRaw, cloud, (size_t)n * 6, cudaMemcpyDefault, c->stream ) );
  RB_CUDA( cudaMemcpyAsync( dQ, queries, (size_t)nq * 6, cudaMemcpyDefault, c->stream ) );
  RB_LAUNCH( "kd_unpack", k_unpack_positions, rb_div_up( n, TPB ), TPB, 0, dRaw, n, dPos );
  int ox = 0, oy = 0, oz = 0;
  cudaPointerAttributes attr{};
  if ( cudaPointerGetAttributes( &attr, cloud ) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered ||
       attr.type == cudaMemoryTypeHost ) {
    cudaGetLastError();
    ox = oy = oz = 32767;
    for ( int64_t i = 0; i < n; i++ ) {
      ox = std::min<int>( ox, cloud[3 * i] );
      oy = std::min<int>( oy, cloud[3 * i + 1] );
      oz = std::min<int>( oz, cloud[3 * i + 2] );
    }
  }
  std::vector<int64_t> hOff{0, n};
  RB_CUDA( S->off.ensure( 16 ) );
  {
    int64_t* hp = (int64_t*)rb_pinned( c, 16 );
    if ( !hp ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
    hp[0] = 0, hp[1] = n;
    RB_CUDA( cudaMemcpyAsync( S->off.p, hp, 16, cudaMemcpyHostToDevice, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
  }
  int r = rb_kd_build( c, S->kd, dPos, S->off.as<int64_t>(), hOff, ox, oy, oz );
  if ( r ) { return r; }
  const double rad     = sqrt( radius2 ) + 1.0;
  const int    cap     = std::max( std::max( 256, maxResults ), 2 * (int)( 4.19 * rad * rad * rad ) );
  const int    threads = (int)std::min<int64_t>( 148 * 4 * 128, ( ( nq + 127 ) / 128 ) * 128 );
  RB_CUDA( S->candKey.ensure( (size_t)threads * cap * 8 ) );
  RB_CUDA( S->small.ensure( 64 ) );
  RB_CUDA( cudaMemsetAsync( S->small.p, 0, 64, c->stream ) );
  RB_LAUNCH( "kd_radius", k_radius_query, threads / 128, 128, 0, S->kd.forest, dQ, nq, (uint32_t)ceil( radius2 ), (uint32_t)floor( radius2 ),
             maxResults, sorted, S->candKey.as<uint64_t>(), cap, dIdx, dDst, dCnt, S->small.as<uint32_t>() );
  uint32_t* h = (uint32_t*)rb_pinned( c, 64 );
  if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  RB_CUDA( cudaMemcpyAsync( h, S->small.p, 4, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaMemcpyAsync( outIdx, dIdx, (size_t)nq * maxResults * 8, cudaMemcpyDefault, c->stream ) );
  RB_CUDA( cudaMemcpyAsync( outDist, dDst, (size_t)nq * maxResults * 8, cudaMemcpyDefault, c->stream ) );
  RB_CUDA( cudaMemcpyAsync( outCount, dCnt, (size_t)nq * 4, cudaMemcpyDefault, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  if ( h[0] ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "kdtree_search_radius: more than %d points inside the radius of one query", cap );
  }
  return RB200_OK;
}
