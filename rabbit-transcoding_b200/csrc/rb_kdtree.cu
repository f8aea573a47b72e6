// rb_kdtree.cu — builds, on the GPU, exactly the kd-trees nanoflann would build on the CPU (see rb_kdtree.cuh).
//
// Restates KDTreeSingleIndexAdaptor::buildIndex / computeBoundingBox / divideTree / middleSplit_ / planeSplit /
// computeMinMax (dependencies/nanoflann/nanoflann.hpp:858-866, 1009-1181) for a FOREST of trees (one per cloud of a
// GOF).  The element order inside every leaf is part of the result (ties of a kNN query come back in traversal
// order), so planeSplit's two Hoare passes are reproduced as the permutation they are: with lim1 = #(v < cut),
//   pass 1 swaps the i-th element >= cut of [0, lim1) (ascending) with the i-th element < cut of [lim1, n) (descending),
//   pass 2 does the same on [lim1, n) with "<= cut" and lim2 = #(v <= cut).
// Two phases:
//  * level phase, nodes of more than KS_CAP elements: every node is cut into chunks of GT consecutive elements, one CTA
//    per chunk, so no pass has to look up "which node does this element belong to".  The predicates of a level are
//    written ONCE as bit masks (2 bits per element); all ranks (prefix counts inside a node) come from the masks and a
//    per-chunk prefix — 1/32 of the element traffic.  A Hoare pass is "misplaced elements to a staging array at their
//    rank" + "every misplaced position reads its partner by rank": coalesced on both sides, no pair lists, no scans over
//    the elements.  The last pass of a level also reduces the tight boxes of the two children (the next level's
//    computeMinMax).  ~44 bytes per element and level.
//  * subtree phase, nodes of at most KS_CAP elements: one WARP owns a subtree in its slice of shared memory and builds it
//    depth first (it splits a node, keeps the left child and pushes the right one on its own stack); every pass is a
//    lane-strided loop with ballots — no CTA barrier, no lock, no atomics on the elements.  Warps are persistent and
//    take subtrees from a global counter, so an SM always holds ~20 independent instruction streams.
// divlow / divhigh come from the children's tight boxes (divideTree :1080-1081).
#include <algorithm>

#include "rb_common.cuh"
#include "rb_kdtree.cuh"
#include "rb_kdtree_build.cuh"

namespace {

constexpr int TPB    = 256;
constexpr int GT     = 2048;          // elements per chunk of the level phase
constexpr int GWORDS = GT / 32;       // mask words per chunk
constexpr int GEPT   = GT / TPB;      // elements per thread
constexpr int KS_CAP   = 512;         // largest node of the subtree phase
constexpr int KS_WARPS = 4;           // warps (= independent subtrees) per CTA
constexpr int KS_STACK = 56;          // pending right children of one subtree: <= its depth (<= 36 + 9, see below)

// counters[]: [0] next free level-phase node, [1] small roots, [2] split nodes of the level, [3] pool exhausted,
//             [4] depth of the deepest subtree, [5] coordinate range error, [6] chunks of the level, [7] subtree stack
//             overflow, [8] next subtree to build (work counter of the persistent warps)
enum { C_NEXT = 0, C_SMALL = 1, C_BIG = 2, C_POOL = 3, C_DEPTH = 4, C_RANGE = 5, C_CHUNKS = 6, C_STACK = 7, C_WORK = 8, C_LARGE = 9 };
constexpr uint32_t P2_WARP_MAX = 16384;  // nodes up to this size get a warp for their second Hoare pass, larger ones a CTA

// build-time node of the level phase
struct GNode {
  uint32_t left, right;   // element range [left, right) in rec
  uint32_t child1;        // 0: not split here
  uint32_t lim1, lim2;    // elements < cutval, <= cutval
  uint32_t m1, m2;        // misplaced pairs of the two Hoare passes
  uint32_t idx;           // the left child takes [left, left + idx)
  uint32_t firstChunk;
  int16_t  lo[3], hi[3];      // the (loose) box handed down by the parent (divideTree's bbox argument)
  int16_t  tmin[3], tmax[3];  // tight box of the node's points (what divideTree hands back up)
  int16_t  cutval;
  int8_t   cutfeat;
  int8_t   state;             // 0 new, 1 split by the level phase, 2 root of a shared-memory subtree
};

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm( "mov.u32 %0, %%lanemask_lt;" : "=r"( m ) );
  return m;
}

// middleSplit_ (nanoflann.hpp:1103-1142): cut axis and cut value from the loose box lo/hi and the tight box tmin/tmax
__device__ __forceinline__ void kd_choose_split( const int lo[3], const int hi[3], const int tmin[3], const int tmax[3], const int o[3],
                                                 int& cutfeat, int& cutval ) {
  int max_span = hi[0] - lo[0];
  for ( int i = 1; i < 3; i++ ) { max_span = max( max_span, hi[i] - lo[i] ); }
  int cf = 0, max_spread = -1;
  for ( int i = 0; i < 3; i++ ) {
    const int span = hi[i] - lo[i];
    // span > (1 - EPS) * max_span in double (:1112): for integer spans below 100000 the right side lies strictly between
    // max_span - 1 and max_span, so the test is span == max_span — and false for every axis when max_span == 0
    if ( span == max_span && max_span > 0 ) {
      const int spread = tmax[i] - tmin[i];
      if ( spread > max_spread ) {
        cf         = i;
        max_spread = spread;
      }
    }
  }
  const int loc = cf == 0 ? lo[0] : ( cf == 1 ? lo[1] : lo[2] ), hic = cf == 0 ? hi[0] : ( cf == 1 ? hi[1] : hi[2] );
  const int tmn = cf == 0 ? tmin[0] : ( cf == 1 ? tmin[1] : tmin[2] ), tmx = cf == 0 ? tmax[0] : ( cf == 1 ? tmax[1] : tmax[2] );
  const int oc  = cf == 0 ? o[0] : ( cf == 1 ? o[1] : o[2] );
  // split_val = (bbox.low + bbox.high) / 2 is an int division of the ABSOLUTE coordinates (truncation toward zero)
  const int split = ( ( loc + oc ) + ( hic + oc ) ) / 2 - oc;
  cutfeat         = cf;
  cutval          = split < tmn ? tmn : ( split > tmx ? tmx : split );
}

// :1137-1139
__device__ __forceinline__ uint32_t kd_split_index( uint32_t count, uint32_t lim1, uint32_t lim2 ) {
  if ( lim1 > count / 2 ) { return lim1; }
  if ( lim2 < count / 2 ) { return lim2; }
  return count / 2;
}

// ---------------------------------------------------------------------------------------------------
// level phase
// ---------------------------------------------------------------------------------------------------

// element records from positions (all clouds of the forest are concatenated; tree t owns [off[t], off[t+1])) and the
// tight boxes of the roots (computeBoundingBox, :1009-1024): one RED per CTA and box side in the common case
__global__ void __launch_bounds__( TPB ) k_g_init( const short4* __restrict__ pos, const int64_t* __restrict__ off, int nTrees, int64_t E,
                                                   int ox, int oy, int oz, uint64_t* __restrict__ rec, int32_t* __restrict__ st,
                                                   uint32_t* __restrict__ counters ) {
  __shared__ int sTree[TPB / 32];
  __shared__ int sBox[TPB / 32][6];
  const int64_t e0   = (int64_t)blockIdx.x * ( TPB * GEPT );
  const int     lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // the tree of the first element of the CTA; a CTA rarely crosses into the next tree
  int t0;
  {
    int lo = 0, hi = nTrees - 1;
    while ( lo < hi ) {
      const int mid = ( lo + hi + 1 ) >> 1;
      if ( off[mid] <= e0 ) {
        lo = mid;
      } else {
        hi = mid - 1;
      }
    }
    t0 = lo;
  }
  int  mn[3] = {1 << 20, 1 << 20, 1 << 20}, mx[3] = {-( 1 << 20 ), -( 1 << 20 ), -( 1 << 20 )};
  int  cur   = t0;         // tree the running box belongs to
  bool bad   = false;
#pragma unroll
  for ( int q = 0; q < GEPT; q++ ) {
    const int64_t e = e0 + q * TPB + threadIdx.x;
    if ( e >= E ) { break; }
    int t = cur;
    while ( e >= off[t + 1] ) { t++; }  // empty clouds are rejected by the host
    if ( t != cur ) {  // this thread walked into the next tree: flush what it has
      if ( mn[0] <= mx[0] ) {
        for ( int k = 0; k < 3; k++ ) {
          atomicMin( &st[(size_t)( cur + 1 ) * 6 + k], mn[k] );
          atomicMax( &st[(size_t)( cur + 1 ) * 6 + 3 + k], mx[k] );
        }
      }
      for ( int k = 0; k < 3; k++ ) { mn[k] = 1 << 20, mx[k] = -( 1 << 20 ); }
      cur = t;
    }
    const short4 p = pos[e];
    const int    x = p.x - ox, y = p.y - oy, z = p.z - oz;
    bad |= (unsigned)x > 4095u || (unsigned)y > 4095u || (unsigned)z > 4095u;
    rec[e] = (uint64_t)( x & 0xFFF ) | ( (uint64_t)( y & 0xFFF ) << 12 ) | ( (uint64_t)( z & 0xFFF ) << 24 ) |
             ( (uint64_t)( e - off[t] ) << 36 );
    mn[0] = min( mn[0], x ), mn[1] = min( mn[1], y ), mn[2] = min( mn[2], z );
    mx[0] = max( mx[0], x ), mx[1] = max( mx[1], y ), mx[2] = max( mx[2], z );
  }
  if ( bad ) { atomicOr( &counters[C_RANGE], 1u ); }
  // combine: warps whose lanes all ended in the same tree reduce with shuffles; the others flush per lane
  const bool     have = mn[0] <= mx[0];
  const uint32_t act  = __ballot_sync( 0xFFFFFFFFu, have );
  int            same = 0;
  if ( act == 0xFFFFFFFFu ) { __match_all_sync( 0xFFFFFFFFu, cur, &same ); }
  if ( same ) {
#pragma unroll
    for ( int k = 0; k < 3; k++ ) {
      mn[k] = __reduce_min_sync( 0xFFFFFFFFu, mn[k] );
      mx[k] = __reduce_max_sync( 0xFFFFFFFFu, mx[k] );
    }
    if ( lane == 0 ) {
      sTree[w] = cur;
      for ( int k = 0; k < 3; k++ ) { sBox[w][k] = mn[k], sBox[w][3 + k] = mx[k]; }
    }
  } else {
    if ( lane == 0 ) { sTree[w] = -1; }
    if ( have ) {
      for ( int k = 0; k < 3; k++ ) {
        atomicMin( &st[(size_t)( cur + 1 ) * 6 + k], mn[k] );
        atomicMax( &st[(size_t)( cur + 1 ) * 6 + 3 + k], mx[k] );
      }
    }
  }
  __syncthreads();
  if ( threadIdx.x < TPB / 32 ) {  // a run of warps with the same tree is flushed by its first warp
    const int i = threadIdx.x, t = sTree[i];
    if ( t >= 0 && ( i == 0 || sTree[i - 1] != t ) ) {
      int b[6];
      for ( int k = 0; k < 6; k++ ) { b[k] = sBox[i][k]; }
      for ( int j = i + 1; j < TPB / 32 && sTree[j] == t; j++ ) {
        for ( int k = 0; k < 3; k++ ) { b[k] = min( b[k], sBox[j][k] ), b[3 + k] = max( b[3 + k], sBox[j][3 + k] ); }
      }
      for ( int k = 0; k < 3; k++ ) {
        atomicMin( &st[(size_t)( t + 1 ) * 6 + k], b[k] );
        atomicMax( &st[(size_t)( t + 1 ) * 6 + 3 + k], b[3 + k] );
      }
    }
  }
}

__global__ void k_g_roots( GNode* __restrict__ nodes, const int64_t* __restrict__ off, int nTrees, int32_t* __restrict__ st ) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if ( t >= nTrees ) { return; }
  for ( int k = 0; k < 3; k++ ) {
    st[(size_t)( t + 1 ) * 6 + k]     = 0x7FFFFFFF;
    st[(size_t)( t + 1 ) * 6 + 3 + k] = (int32_t)0x80000000;
  }
  GNode n{};
  n.left       = (uint32_t)off[t];
  n.right      = (uint32_t)off[t + 1];
  nodes[t + 1] = n;
  if ( t == 0 ) {
    GNode z{};
    nodes[0] = z;
  }
}

// One chunk of GT consecutive elements of a node of the current level: everything an element pass needs, in one
// 48-byte record (written by the CTA of k_g_count, completed by the node scans), so the CTAs of the other passes start
// with ONE metadata round trip.
struct GChunk {
  uint32_t e0, left, cnt;  // first element (global), first element of the node, elements in the chunk
  uint32_t node;
  uint32_t cutcf;          // (uint16) cutval | cutfeat << 16
  uint32_t lim1, lim2;     // node-wide: elements < cutval, <= cutval               (k_g_nodescan<false>)
  uint32_t m, pre;         // misplaced pairs of the running Hoare pass and the exclusive prefix of its mask in the node
  uint32_t idx, child1;    // the left child takes positions [0, idx); children are child1, child1 + 1
  uint32_t pad;
};
constexpr int BIG = 1 << 20;

// the nodes of this level: tight box from the statistics, small root / split decision (middleSplit_), chunk records and
// the ids of the children.  One thread per node; chunks and child ids are handed out with one atomic per warp (the
// numbering of nodes and chunks is free: only the tree they describe is the result).  C_BIG / C_CHUNKS are zeroed by
// the host before the launch; nothing else allocates from C_NEXT while this kernel runs.
__global__ void __launch_bounds__( TPB ) k_g_setup( GNode* __restrict__ nodes, const int32_t* __restrict__ st, uint32_t lvlBegin,
                                                    uint32_t lvlEnd, int isRoot, int ox, int oy, int oz,
                                                    uint32_t* __restrict__ smallRoots, uint32_t* __restrict__ counters,
                                                    uint32_t* __restrict__ chunkNode, uint32_t chunkCap, uint32_t nodeCap,
                                                    int32_t* __restrict__ cls, uint32_t* __restrict__ largeList, uint32_t ksCap ) {
  const int      lane = threadIdx.x & 31;
  const uint32_t i    = lvlBegin + blockIdx.x * TPB + threadIdx.x;
  const int      o[3] = {ox, oy, oz};
  uint32_t       nch = 0, big = 0, small = 0;
  int            cf = 0, cv = 0;
  if ( i < lvlEnd ) {
    GNode& n = nodes[i];
    int    lo[3], hi[3], tmin[3], tmax[3];
    for ( int k = 0; k < 3; k++ ) {
      tmin[k]   = st[(size_t)i * 6 + k];
      tmax[k]   = st[(size_t)i * 6 + 3 + k];
      n.tmin[k] = (int16_t)tmin[k];
      n.tmax[k] = (int16_t)tmax[k];
      if ( isRoot ) {  // divideTree( 0, N, root_bbox ) starts from the tight box
        n.lo[k] = (int16_t)tmin[k];
        n.hi[k] = (int16_t)tmax[k];
      }
      lo[k] = n.lo[k], hi[k] = n.hi[k];
    }
    const uint32_t count = n.right - n.left;
    if ( count <= ksCap ) {
      n.state = 2;
      small   = 1;
    } else {
      kd_choose_split( lo, hi, tmin, tmax, o, cf, cv );
      n.cutfeat = (int8_t)cf;
      n.cutval  = (int16_t)cv;
      n.state   = 1;
      nch       = ( count + GT - 1 ) / GT;
      big       = 1;
    }
  }
  // warp-wide exclusive scans, one allocation per warp
  uint32_t ia = nch, ib = big, is = small;
#pragma unroll
  for ( int d = 1; d < 32; d <<= 1 ) {
    const uint32_t ta = __shfl_up_sync( 0xFFFFFFFFu, ia, d ), tb = __shfl_up_sync( 0xFFFFFFFFu, ib, d ),
                   ts = __shfl_up_sync( 0xFFFFFFFFu, is, d );
    if ( lane >= d ) { ia += ta, ib += tb, is += ts; }
  }
  uint32_t baseC = 0, baseB = 0, baseS = 0;
  if ( lane == 31 ) {
    if ( ia ) { baseC = atomicAdd( &counters[C_CHUNKS], ia ); }
    if ( ib ) {
      baseB = atomicAdd( &counters[C_NEXT], 2u * ib );
      atomicAdd( &counters[C_BIG], ib );
    }
    if ( is ) { baseS = atomicAdd( &counters[C_SMALL], is ); }
  }
  baseC = __shfl_sync( 0xFFFFFFFFu, baseC, 31 );
  baseB = __shfl_sync( 0xFFFFFFFFu, baseB, 31 );
  baseS = __shfl_sync( 0xFFFFFFFFu, baseS, 31 );
  if ( small ) { smallRoots[baseS + is - 1u] = i; }
  if ( big ) {
    GNode&         n    = nodes[i];
    const uint32_t offC = baseC + ia - nch, c1 = baseB + 2u * ( ib - 1u );
    if ( c1 + 2u > nodeCap || offC + nch > chunkCap ) {
      counters[C_POOL] = 1;
      n.child1         = 0;
      n.firstChunk     = 0;
    } else {
      n.child1     = c1;
      n.firstChunk = offC;
      for ( uint32_t k = 0; k < nch; k++ ) { chunkNode[offC + k] = i; }
      if ( n.right - n.left > P2_WARP_MAX ) { largeList[atomicAdd( &counters[C_LARGE], 1u )] = i; }
      // boxes of the classes "< cut" and "> cut" of planeSplit, filled by k_g_count
      for ( int j = 0; j < 12; j++ ) { cls[(size_t)i * 12 + j] = ( j % 6 ) < 3 ? BIG : -BIG; }
    }
  }
}

// pass A: the predicates of planeSplit for every element of the level as bit masks, their per-chunk counts, and the
// boxes of the classes "< cut" and "> cut": the children's tight boxes follow from those (the few elements == cut are
// added by k_g_apply2 once their final positions are known) without another pass over the elements.  The CTA also
// writes the chunk record the later passes start from.
__global__ void __launch_bounds__( TPB ) k_g_count( const uint64_t* __restrict__ rec, const GNode* __restrict__ nodes,
                                                    const uint32_t* __restrict__ chunkNode, GChunk* __restrict__ chunks,
                                                    uint32_t* __restrict__ wA, uint32_t* __restrict__ wB,
                                                    uint32_t* __restrict__ cA, uint32_t* __restrict__ cB, int32_t* __restrict__ cls ) {
  __shared__ uint32_t sA[TPB / 32], sB[TPB / 32];
  __shared__ int      sBox[TPB / 32][12];
  const uint32_t c = blockIdx.x, node = chunkNode[c];
  const GNode&   n = nodes[node];
  const uint32_t e0 = n.left + ( c - n.firstChunk ) * (uint32_t)GT, cnt = min( (uint32_t)GT, n.right - e0 );
  const int      cf = n.cutfeat, cut = n.cutval;
  const int      lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if ( threadIdx.x == 0 ) {
    GChunk d{};
    d.e0 = e0, d.left = n.left, d.cnt = cnt, d.node = node, d.cutcf = ( (uint32_t)cut & 0xFFFFu ) | ( (uint32_t)cf << 16 );
    d.child1  = n.child1;
    chunks[c] = d;
  }
  uint64_t r[GEPT];
#pragma unroll
  for ( int q = 0; q < GEPT; q++ ) {
    const uint32_t i = q * TPB + threadIdx.x;
    r[q]             = i < cnt ? rec[e0 + i] : ~0ull;
  }
  int bx[12];  // {min[3], max[3]} of "< cut", then of "> cut"
#pragma unroll
  for ( int j = 0; j < 12; j++ ) { bx[j] = ( j % 6 ) < 3 ? BIG : -BIG; }
  uint32_t ca = 0, cb = 0;
#pragma unroll
  for ( int q = 0; q < GEPT; q++ ) {
    const uint32_t i  = q * TPB + threadIdx.x;
    const bool     ok = i < cnt;
    const int      x = kd_coord( r[q], 0 ), y = kd_coord( r[q], 1 ), z = kd_coord( r[q], 2 );
    const int      v  = cf == 0 ? x : ( cf == 1 ? y : z );
    const bool     lt = ok && v < cut, le = ok && v <= cut;
    const uint32_t ma = __ballot_sync( 0xFFFFFFFFu, lt ), mb = __ballot_sync( 0xFFFFFFFFu, le );
    if ( lane == 0 ) {
      wA[(size_t)c * GWORDS + q * ( TPB / 32 ) + w] = ma;
      wB[(size_t)c * GWORDS + q * ( TPB / 32 ) + w] = mb;
    }
    ca += __popc( ma ), cb += __popc( mb );
    if ( lt ) {
      bx[0] = min( bx[0], x ), bx[1] = min( bx[1], y ), bx[2] = min( bx[2], z );
      bx[3] = max( bx[3], x ), bx[4] = max( bx[4], y ), bx[5] = max( bx[5], z );
    } else if ( ok && !le ) {
      bx[6] = min( bx[6], x ), bx[7] = min( bx[7], y ), bx[8] = min( bx[8], z );
      bx[9] = max( bx[9], x ), bx[10] = max( bx[10], y ), bx[11] = max( bx[11], z );
    }
  }
#pragma unroll
  for ( int j = 0; j < 12; j++ ) {
    bx[j] = ( j % 6 ) < 3 ? __reduce_min_sync( 0xFFFFFFFFu, bx[j] ) : __reduce_max_sync( 0xFFFFFFFFu, bx[j] );
  }
  if ( lane == 0 ) {
    sA[w] = ca, sB[w] = cb;
#pragma unroll
    for ( int j = 0; j < 12; j++ ) { sBox[w][j] = bx[j]; }
  }
  __syncthreads();
  if ( threadIdx.x == 0 ) {
    uint32_t ta = 0, tb = 0;
    for ( int k = 0; k < TPB / 32; k++ ) { ta += sA[k], tb += sB[k]; }
    cA[c] = ta, cB[c] = tb;
  }
  if ( threadIdx.x >= 32 && threadIdx.x < 32 + 12 ) {
    const int  j     = threadIdx.x - 32;
    const bool isMin = ( j % 6 ) < 3;
    int        v     = sBox[0][j];
    for ( int k = 1; k < TPB / 32; k++ ) { v = isMin ? min( v, sBox[k][j] ) : max( v, sBox[k][j] ); }
    if ( isMin ) {
      if ( v < BIG ) { atomicMin( &cls[(size_t)node * 12 + j], v ); }
    } else {
      if ( v > -BIG ) { atomicMax( &cls[(size_t)node * 12 + j], v ); }
    }
  }
}

// number of set bits of the node's mask `words` in positions [0, pos): chunk prefix + whole words + partial word
__device__ __forceinline__ uint32_t mask_prefix_at( const uint32_t* __restrict__ words, uint32_t firstChunk, uint32_t chunkExcl,
                                                    uint32_t pos, int lane ) {
  const uint32_t  ch = pos / GT, wi = ( pos % GT ) >> 5, bit = pos & 31u;
  const uint32_t* W  = words + (size_t)( firstChunk + ch ) * GWORDS;
  uint32_t        s  = 0;
  for ( uint32_t j = lane; j <= wi; j += 32 ) {
    const uint32_t v = W[j];
    s += j < wi ? __popc( v ) : __popc( v & ( ( 1u << bit ) - 1u ) );
  }
  return chunkExcl + __reduce_add_sync( 0xFFFFFFFFu, s );
}

// pass B / E: one warp per node of the level.  Exclusive prefix of the chunk counts, lim1 / lim2, the number of
// misplaced pairs — completed into the chunk records — and (SECOND == false) the children (divideTree :1070-1078) with
// their tight boxes from the class boxes of k_g_count
template <bool SECOND>
__global__ void __launch_bounds__( TPB ) k_g_nodescan( GNode* __restrict__ nodes, uint32_t lvlBegin, uint32_t lvlEnd,
                                                       GChunk* __restrict__ chunks, uint32_t* __restrict__ cA,
                                                       uint32_t* __restrict__ cB, const uint32_t* __restrict__ wA,
                                                       const uint32_t* __restrict__ wB, int32_t* __restrict__ st,
                                                       const int32_t* __restrict__ cls ) {
  const uint32_t i    = lvlBegin + ( blockIdx.x * TPB + threadIdx.x ) / 32;
  const int      lane = threadIdx.x & 31;
  if ( i >= lvlEnd ) { return; }
  GNode& n = nodes[i];
  if ( n.state != 1 || n.child1 == 0 ) { return; }
  const uint32_t count = n.right - n.left, nch = ( count + GT - 1 ) / GT, first = n.firstChunk;
  if ( !SECOND ) {
    uint32_t runA = 0, runB = 0;
    for ( uint32_t k0 = 0; k0 < nch; k0 += 32 ) {
      const uint32_t k = k0 + lane;
      const uint32_t a = k < nch ? cA[first + k] : 0u, b = k < nch ? cB[first + k] : 0u;
      uint32_t       ia = a, ib = b;
#pragma unroll
      for ( int d = 1; d < 32; d <<= 1 ) {
        const uint32_t ta = __shfl_up_sync( 0xFFFFFFFFu, ia, d ), tb = __shfl_up_sync( 0xFFFFFFFFu, ib, d );
        if ( lane >= d ) { ia += ta, ib += tb; }
      }
      if ( k < nch ) { cA[first + k] = runA + ia - a, cB[first + k] = runB + ib - b; }
      runA += __shfl_sync( 0xFFFFFFFFu, ia, 31 );
      runB += __shfl_sync( 0xFFFFFFFFu, ib, 31 );
    }
    __syncwarp();
    const uint32_t lim1 = runA, lim2 = runB;
    // misplaced on the left of pass 1 = elements >= cut among the first lim1 = lim1 - #(v < cut in [0, lim1))
    uint32_t m1 = 0;
    if ( lim1 > 0 && lim1 < count ) { m1 = lim1 - mask_prefix_at( wA, first, cA[first + lim1 / GT], lim1, lane ); }
    const uint32_t idx = kd_split_index( count, lim1, lim2 );
    for ( uint32_t k = lane; k < nch; k += 32 ) {
      GChunk& d = chunks[first + k];
      d.lim1 = lim1, d.lim2 = lim2, d.m = m1, d.pre = cA[first + k], d.idx = idx;
    }
    if ( lane == 0 ) {
      n.lim1 = lim1, n.lim2 = lim2, n.m1 = m1, n.idx = idx;
      GNode a{}, b{};
      a.left = n.left, a.right = n.left + idx, b.left = n.left + idx, b.right = n.right;
      for ( int k = 0; k < 3; k++ ) {
        a.lo[k] = b.lo[k] = n.lo[k];
        a.hi[k] = b.hi[k] = n.hi[k];
      }
      a.hi[n.cutfeat]     = n.cutval;  // left_bbox[cutfeat].high = cutval
      b.lo[n.cutfeat]     = n.cutval;  // right_bbox[cutfeat].low = cutval
      nodes[n.child1]     = a;
      nodes[n.child1 + 1] = b;
      // tight boxes of the children so far: the left child holds the class "< cut", the right child the class "> cut";
      // k_g_apply2 adds the elements == cut (positions [lim1, lim2) after pass 2) to the side they end up on
      const int32_t* C = cls + (size_t)i * 12;
      for ( int k = 0; k < 3; k++ ) {
        st[(size_t)n.child1 * 6 + k]             = C[k] >= BIG ? 0x7FFFFFFF : C[k];
        st[(size_t)n.child1 * 6 + 3 + k]         = C[3 + k] <= -BIG ? (int32_t)0x80000000 : C[3 + k];
        st[(size_t)( n.child1 + 1 ) * 6 + k]     = C[6 + k] >= BIG ? 0x7FFFFFFF : C[6 + k];
        st[(size_t)( n.child1 + 1 ) * 6 + 3 + k] = C[9 + k] <= -BIG ? (int32_t)0x80000000 : C[9 + k];
      }
    }
  } else {
    // cB now holds the per-chunk counts of "<= cut" among the positions >= lim1 AFTER pass 1
    uint32_t run = 0;
    for ( uint32_t k0 = 0; k0 < nch; k0 += 32 ) {
      const uint32_t k = k0 + lane;
      const uint32_t b = k < nch ? cB[first + k] : 0u;
      uint32_t       ib = b;
#pragma unroll
      for ( int d = 1; d < 32; d <<= 1 ) {
        const uint32_t tb = __shfl_up_sync( 0xFFFFFFFFu, ib, d );
        if ( lane >= d ) { ib += tb; }
      }
      if ( k < nch ) { cB[first + k] = run + ib - b; }
      run += __shfl_sync( 0xFFFFFFFFu, ib, 31 );
    }
    __syncwarp();
    const uint32_t lim1 = n.lim1, lim2 = n.lim2;
    uint32_t       m2 = 0;
    if ( lim2 > lim1 && lim2 < count ) { m2 = ( lim2 - lim1 ) - mask_prefix_at( wB, first, cB[first + lim2 / GT], lim2, lane ); }
    for ( uint32_t k = lane; k < nch; k += 32 ) {
      GChunk& d = chunks[first + k];
      d.m = m2, d.pre = cB[first + k];
    }
    if ( lane == 0 ) { n.m2 = m2; }
  }
}

// per-chunk prefix of the mask words in shared memory: pre[w] = set bits in words [0, w) of the chunk
__device__ __forceinline__ void chunk_word_prefix( const uint32_t* __restrict__ words, uint32_t* sWord, uint32_t* sPre ) {
  if ( threadIdx.x < GWORDS ) { sWord[threadIdx.x] = words[threadIdx.x]; }
  __syncthreads();
  if ( threadIdx.x < 32 ) {  // GWORDS == 64: two words per lane
    const int      l = threadIdx.x;
    const uint32_t a = __popc( sWord[2 * l] ), b = __popc( sWord[2 * l + 1] );
    uint32_t       in = a + b;
#pragma unroll
    for ( int d = 1; d < 32; d <<= 1 ) {
      const uint32_t t = __shfl_up_sync( 0xFFFFFFFFu, in, d );
      if ( l >= d ) { in += t; }
    }
    sPre[2 * l]     = in - a - b;
    sPre[2 * l + 1] = in - b;
  }
  __syncthreads();
}
static_assert( GWORDS == 64, "chunk_word_prefix handles two words per lane" );

// pass C / F: the misplaced elements of a Hoare pass go to the staging array at their rank.
//   pass 1 (SECOND == false): left = positions < lim1 with v >= cut -> tmp[left + i]; right = positions >= lim1 with
//                             v < cut -> tmp[left + lim1 + r] (r = rank from the left)
//   pass 2 (SECOND == true) : left = positions in [lim1, lim2) with v > cut -> tmp[left + lim1 + i]; right = positions >= lim2
//                             with v <= cut -> tmp[left + lim2 + r]
template <bool SECOND>
__global__ void __launch_bounds__( TPB ) k_g_stage( const uint64_t* __restrict__ rec, const GChunk* __restrict__ chunks,
                                                    const uint32_t* __restrict__ words, uint64_t* __restrict__ tmp ) {
  __shared__ uint32_t sWord[GWORDS], sPre[GWORDS];
  const uint32_t c = blockIdx.x;
  const GChunk&  D = chunks[c];
  const uint32_t m = D.m;
  if ( m == 0 ) { return; }
  const uint32_t e0 = D.e0, left = D.left, cnt = D.cnt, p0 = e0 - left, lim1 = D.lim1, lim2 = D.lim2, base = D.pre;
  // nothing of this chunk takes part in pass 2 when it lies below lim1
  if ( SECOND && p0 + cnt <= lim1 ) { return; }
  chunk_word_prefix( words + (size_t)c * GWORDS, sWord, sPre );
  const int      lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt   = lanemask_lt();
#pragma unroll
  for ( int q = 0; q < GEPT; q++ ) {
    const uint32_t i = q * TPB + threadIdx.x;
    if ( i >= cnt ) { continue; }
    const int      wi  = q * ( TPB / 32 ) + w;
    const uint32_t wd  = sWord[wi];
    const bool     bit = ( wd >> lane ) & 1u;
    const uint32_t pre = base + sPre[wi] + __popc( wd & lt );  // set bits of the node's mask before this position
    const uint32_t p   = p0 + i;
    if ( !SECOND ) {
      if ( p < lim1 ) {
        if ( !bit ) { tmp[left + ( p - pre )] = rec[e0 + i]; }
      } else if ( bit ) {
        tmp[left + lim1 + ( pre - ( lim1 - m ) )] = rec[e0 + i];
      }
    } else if ( p >= lim1 ) {
      if ( p < lim2 ) {
        if ( !bit ) { tmp[left + lim1 + ( ( p - lim1 ) - pre )] = rec[e0 + i]; }
      } else if ( bit ) {
        tmp[left + lim2 + ( pre - ( ( lim2 - lim1 ) - m ) )] = rec[e0 + i];
      }
    }
  }
}

// pass D: every misplaced position of pass 1 takes its partner; the "<= cut" mask of the positions >= lim1 is written
// for the new arrangement (pass 2 works on it: wC) together with its per-chunk counts.  All loads of the chunk are
// issued before the first ballot, so a CTA pays one memory round trip, not one per 256 elements.
__global__ void __launch_bounds__( TPB ) k_g_apply1( uint64_t* __restrict__ rec, const GChunk* __restrict__ chunks,
                                                     const uint32_t* __restrict__ wA, const uint32_t* __restrict__ wB,
                                                     uint32_t* __restrict__ wC, uint32_t* __restrict__ cB,
                                                     const uint64_t* __restrict__ tmp ) {
  __shared__ uint32_t sWord[GWORDS], sPre[GWORDS], sWordB[GWORDS];
  __shared__ uint32_t sCnt[TPB / 32];
  const uint32_t c = blockIdx.x;
  const GChunk&  D = chunks[c];
  const uint32_t e0 = D.e0, left = D.left, cnt = D.cnt, p0 = e0 - left, lim1 = D.lim1, m = D.m, base = D.pre;
  const int      cf = ( D.cutcf >> 16 ) & 3, cut = (int16_t)( D.cutcf & 0xFFFFu );
  if ( threadIdx.x < GWORDS ) { sWordB[threadIdx.x] = wB[(size_t)c * GWORDS + threadIdx.x]; }
  chunk_word_prefix( wA + (size_t)c * GWORDS, sWord, sPre );
  const int      lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt   = lanemask_lt();
  uint64_t       nv[GEPT];   // the partner of a misplaced position
  uint32_t       kind = 0;   // bit q: position q of this thread is misplaced
#pragma unroll
  for ( int q = 0; q < GEPT; q++ ) {
    const uint32_t i   = q * TPB + threadIdx.x;
    const int      wi  = q * ( TPB / 32 ) + w;
    const uint32_t wd  = sWord[wi];
    const bool     bit = ( wd >> lane ) & 1u;
    const uint32_t pre = base + sPre[wi] + __popc( wd & lt );
    const uint32_t p   = p0 + i;
    nv[q]              = 0;
    if ( i < cnt && m ) {
      if ( p < lim1 ) {
        if ( !bit ) { nv[q] = tmp[left + lim1 + ( m - 1u - ( p - pre ) )], kind |= 1u << q; }
      } else if ( bit ) {
        nv[q] = tmp[left + ( m - 1u - ( pre - ( lim1 - m ) ) )], kind |= 1u << q;
      }
    }
  }
  uint32_t cntB = 0;
#pragma unroll
  for ( int q = 0; q < GEPT; q++ ) {
    const uint32_t i  = q * TPB + threadIdx.x;
    const int      wi = q * ( TPB / 32 ) + w;
    const uint32_t p  = p0 + i;
    bool           b2 = i < cnt && p >= lim1 && ( ( sWordB[wi] >> lane ) & 1u );
    if ( kind >> q & 1u ) {
      rec[e0 + i] = nv[q];
      if ( p >= lim1 ) { b2 = kd_coord( nv[q], cf ) <= cut; }
    }
    const uint32_t mb = __ballot_sync( 0xFFFFFFFFu, b2 );
    if ( lane == 0 ) { wC[(size_t)c * GWORDS + wi] = mb; }
    cntB += __popc( mb );
  }
  if ( lane == 0 ) { sCnt[w] = cntB; }
  __syncthreads();
  if ( threadIdx.x == 0 ) {
    uint32_t t = 0;
    for ( int k = 0; k < TPB / 32; k++ ) { t += sCnt[k]; }
    cB[c] = t;
  }
}

// pass E-G for one node (one CTA per node of the level): the second Hoare pass of planeSplit only concerns the elements
// == cut, a thin slab of the node.  After pass 1 they sit somewhere in [lim1, n); "<= cut" of that range is the mask wC.
// The CTA lists the positions of [lim1, lim2) that hold an element > cut (ascending) and the positions of [lim2, n) that
// hold an element == cut (ascending), swaps the i-th of the first list with the i-th from the END of the second —
// exactly the two-pointer loop (:1169-1181) — and then adds the slab [lim1, lim2), which now holds all elements == cut,
// to the tight box of the child each position belongs to.  Lists longer than P2_CAP are handled in rounds.
constexpr int P2_CAP = 2048;
__device__ __forceinline__ uint32_t node_word( const uint32_t* __restrict__ W, uint32_t wi, uint32_t lo, uint32_t hi, bool invert ) {
  // word wi of the node's mask restricted to positions [lo, hi), optionally inverted
  uint32_t       v    = invert ? ~W[wi] : W[wi];
  const uint32_t base = wi << 5;
  if ( base < lo ) { v &= lo - base >= 32 ? 0u : ( 0xFFFFFFFFu << ( lo - base ) ); }
  if ( base + 32 > hi ) { v &= hi <= base ? 0u : ( 0xFFFFFFFFu >> ( base + 32 - hi ) ); }
  return v;
}
// ordered positions of the set bits of the (restricted) mask with ranks in [r0, r0 + P2_CAP) -> out[rank - r0]; returns the
// total number of set bits.  Every thread owns a contiguous slice of the words.
__device__ __forceinline__ uint32_t node_list( const uint32_t* __restrict__ W, uint32_t lo, uint32_t hi, bool invert, uint32_t r0,
                                               uint32_t* out, uint32_t* sScan ) {
  const uint32_t w0 = lo >> 5, w1 = ( hi + 31 ) >> 5, nw = w1 > w0 ? w1 - w0 : 0, per = ( nw + TPB - 1 ) / TPB;
  const uint32_t a = w0 + threadIdx.x * per, b = min( w1, a + per );
  uint32_t       cnt = 0;
  for ( uint32_t wi = a; wi < b; wi++ ) { cnt += __popc( node_word( W, wi, lo, hi, invert ) ); }
  // CTA exclusive scan of the 256 slice counts
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t  inc  = cnt;
#pragma unroll
  for ( int d = 1; d < 32; d <<= 1 ) {
    const uint32_t t = __shfl_up_sync( 0xFFFFFFFFu, inc, d );
    if ( lane >= d ) { inc += t; }
  }
  __syncthreads();  // sScan may still be read from the previous call
  if ( lane == 31 ) { sScan[w] = inc; }
  __syncthreads();
  uint32_t off = inc - cnt, total = 0;
  for ( int k = 0; k < TPB / 32; k++ ) {
    const uint32_t x = sScan[k];
    if ( k < w ) { off += x; }
    total += x;
  }
  for ( uint32_t wi = a; wi < b; wi++ ) {
    for ( uint32_t v = node_word( W, wi, lo, hi, invert ); v; v &= v - 1 ) {
      if ( off >= r0 && off - r0 < (uint32_t)P2_CAP ) { out[off - r0] = ( wi << 5 ) + ( __ffs( v ) - 1 ); }
      off++;
    }
  }
  return total;
}

__global__ void __launch_bounds__( TPB ) k_g_pass2_node( uint64_t* __restrict__ rec, GNode* __restrict__ nodes,
                                                         const uint32_t* __restrict__ largeList, const uint32_t* __restrict__ wC,
                                                         int32_t* __restrict__ st ) {
  __shared__ uint32_t sL[P2_CAP], sR[P2_CAP], sScan[TPB / 32];
  __shared__ int      sBox[TPB / 32][12];
  const uint32_t i = largeList[blockIdx.x];
  GNode&         n = nodes[i];
  if ( n.state != 1 || n.child1 == 0 ) { return; }
  const uint32_t  left = n.left, count = n.right - n.left, lim1 = n.lim1, lim2 = n.lim2, idx = n.idx, child1 = n.child1;
  const uint32_t* W = wC + (size_t)n.firstChunk * GWORDS;  // chunks are node-aligned: position p is bit p & 31 of word p >> 5
  uint32_t        m = 0;
  if ( lim2 > lim1 && lim2 < count ) {
    for ( uint32_t r0 = 0;; r0 += P2_CAP ) {
      // misplaced on the left: zeros of [lim1, lim2), ranks r0.. ascending; on the right: ones of [lim2, n), ranks from the END
      m = node_list( W, lim1, lim2, true, r0, sL, sScan );
      if ( m == 0 || r0 >= m ) { break; }
      const uint32_t take = min( (uint32_t)P2_CAP, m - r0 );
      // the partners of ranks [r0, r0 + take) are the right ranks m - 1 - r0 down to m - r0 - take
      node_list( W, lim2, count, false, m - r0 - take, sR, sScan );
      __syncthreads();
      for ( uint32_t k = threadIdx.x; k < take; k += TPB ) {
        const uint32_t pl = left + sL[k], pr = left + sR[take - 1 - k];
        const uint64_t a = rec[pl], b = rec[pr];
        rec[pl] = b, rec[pr] = a;
      }
      __syncthreads();
      if ( r0 + take >= m ) { break; }
    }
  }
  if ( threadIdx.x == 0 ) { n.m2 = m; }
  if ( lim2 == lim1 ) { return; }
  // the slab [lim1, lim2): every position holds an element == cut now
  __threadfence_block();
  __syncthreads();
  int bx[12];
#pragma unroll
  for ( int k = 0; k < 3; k++ ) { bx[k] = bx[6 + k] = BIG, bx[3 + k] = bx[9 + k] = -BIG; }
  for ( uint32_t p = lim1 + threadIdx.x; p < lim2; p += TPB ) {
    const uint64_t v  = rec[left + p];
    const int      cx = kd_coord( v, 0 ), cy = kd_coord( v, 1 ), cz = kd_coord( v, 2 );
    if ( p < idx ) {
      bx[0] = min( bx[0], cx ), bx[1] = min( bx[1], cy ), bx[2] = min( bx[2], cz );
      bx[3] = max( bx[3], cx ), bx[4] = max( bx[4], cy ), bx[5] = max( bx[5], cz );
    } else {
      bx[6] = min( bx[6], cx ), bx[7] = min( bx[7], cy ), bx[8] = min( bx[8], cz );
      bx[9] = max( bx[9], cx ), bx[10] = max( bx[10], cy ), bx[11] = max( bx[11], cz );
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for ( int k = 0; k < 12; k++ ) {
    bx[k] = ( k % 6 ) < 3 ? __reduce_min_sync( 0xFFFFFFFFu, bx[k] ) : __reduce_max_sync( 0xFFFFFFFFu, bx[k] );
  }
  if ( lane == 0 ) {
#pragma unroll
    for ( int k = 0; k < 12; k++ ) { sBox[w][k] = bx[k]; }
  }
  __syncthreads();
  if ( threadIdx.x < 12 ) {
    const int  k     = threadIdx.x;
    const bool isMin = ( k % 6 ) < 3;
    int        v     = sBox[0][k];
    for ( int j = 1; j < TPB / 32; j++ ) { v = isMin ? min( v, sBox[j][k] ) : max( v, sBox[j][k] ); }
    int32_t* dst = st + (size_t)( child1 + ( k >= 6 ? 1 : 0 ) ) * 6 + ( k % 6 );
    if ( isMin ) {
      if ( v < BIG ) { atomicMin( dst, v ); }
    } else {
      if ( v > -BIG ) { atomicMax( dst, v ); }
    }
  }
}

// the same for the many nodes of at most P2_WARP_MAX elements: one warp per node (its mask is at most 512 words)
constexpr int P2_WCAP = 256;
__device__ __forceinline__ uint32_t node_list_warp( const uint32_t* __restrict__ W, uint32_t lo, uint32_t hi, bool invert, uint32_t r0,
                                                    uint32_t* out, int lane ) {
  const uint32_t w0 = lo >> 5, w1 = ( hi + 31 ) >> 5, nw = w1 > w0 ? w1 - w0 : 0, per = ( nw + 31 ) / 32;
  const uint32_t a = w0 + lane * per, b = min( w1, a + per );
  uint32_t       cnt = 0;
  for ( uint32_t wi = a; wi < b; wi++ ) { cnt += __popc( node_word( W, wi, lo, hi, invert ) ); }
  uint32_t inc = cnt;
#pragma unroll
  for ( int d = 1; d < 32; d <<= 1 ) {
    const uint32_t t = __shfl_up_sync( 0xFFFFFFFFu, inc, d );
    if ( lane >= d ) { inc += t; }
  }
  const uint32_t total = __shfl_sync( 0xFFFFFFFFu, inc, 31 );
  uint32_t       off   = inc - cnt;
  if ( total == 0 ) { return 0; }
  for ( uint32_t wi = a; wi < b; wi++ ) {
    for ( uint32_t v = node_word( W, wi, lo, hi, invert ); v; v &= v - 1 ) {
      if ( off >= r0 && off - r0 < (uint32_t)P2_WCAP ) { out[off - r0] = ( wi << 5 ) + ( __ffs( v ) - 1 ); }
      off++;
    }
  }
  return total;
}
__global__ void __launch_bounds__( TPB ) k_g_pass2_warp( uint64_t* __restrict__ rec, GNode* __restrict__ nodes, uint32_t lvlBegin,
                                                         uint32_t lvlEnd, const uint32_t* __restrict__ wC, int32_t* __restrict__ st ) {
  __shared__ uint32_t sL[TPB / 32][P2_WCAP], sR[TPB / 32][P2_WCAP];
  const int      lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t i    = lvlBegin + blockIdx.x * ( TPB / 32 ) + w;
  if ( i >= lvlEnd ) { return; }
  GNode& n = nodes[i];
  if ( n.state != 1 || n.child1 == 0 ) { return; }
  const uint32_t left = n.left, count = n.right - n.left, lim1 = n.lim1, lim2 = n.lim2, idx = n.idx, child1 = n.child1;
  if ( count > P2_WARP_MAX || lim2 == lim1 ) { return; }
  const uint32_t* W = wC + (size_t)n.firstChunk * GWORDS;
  if ( lim2 < count ) {
    for ( uint32_t r0 = 0;; r0 += P2_WCAP ) {
      const uint32_t m = node_list_warp( W, lim1, lim2, true, r0, sL[w], lane );
      if ( m == 0 || r0 >= m ) { break; }
      const uint32_t take = min( (uint32_t)P2_WCAP, m - r0 );
      node_list_warp( W, lim2, count, false, m - r0 - take, sR[w], lane );
      __syncwarp();
      for ( uint32_t k = lane; k < take; k += 32 ) {
        const uint32_t pl = left + sL[w][k], pr = left + sR[w][take - 1 - k];
        const uint64_t a = rec[pl], b = rec[pr];
        rec[pl] = b, rec[pr] = a;
      }
      __syncwarp();
      if ( r0 + take >= m ) { break; }
    }
  }
  __threadfence_block();
  __syncwarp();
  int bx[12];
#pragma unroll
  for ( int k = 0; k < 3; k++ ) { bx[k] = bx[6 + k] = BIG, bx[3 + k] = bx[9 + k] = -BIG; }
  for ( uint32_t p = lim1 + lane; p < lim2; p += 32 ) {
    const uint64_t v  = rec[left + p];
    const int      cx = kd_coord( v, 0 ), cy = kd_coord( v, 1 ), cz = kd_coord( v, 2 );
    if ( p < idx ) {
      bx[0] = min( bx[0], cx ), bx[1] = min( bx[1], cy ), bx[2] = min( bx[2], cz );
      bx[3] = max( bx[3], cx ), bx[4] = max( bx[4], cy ), bx[5] = max( bx[5], cz );
    } else {
      bx[6] = min( bx[6], cx ), bx[7] = min( bx[7], cy ), bx[8] = min( bx[8], cz );
      bx[9] = max( bx[9], cx ), bx[10] = max( bx[10], cy ), bx[11] = max( bx[11], cz );
    }
  }
#pragma unroll
  for ( int k = 0; k < 12; k++ ) {
    bx[k] = ( k % 6 ) < 3 ? __reduce_min_sync( 0xFFFFFFFFu, bx[k] ) : __reduce_max_sync( 0xFFFFFFFFu, bx[k] );
  }
  if ( lane < 12 ) {
    int v = 0;
#pragma unroll
    for ( int k = 0; k < 12; k++ ) {
      if ( lane == k ) { v = bx[k]; }
    }
    int32_t* dst = st + (size_t)( child1 + ( lane >= 6 ? 1 : 0 ) ) * 6 + ( lane % 6 );
    if ( ( lane % 6 ) < 3 ) {
      if ( v < BIG ) { atomicMin( dst, v ); }
    } else {
      if ( v > -BIG ) { atomicMax( dst, v ); }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// subtree phase: one WARP builds a whole subtree of at most KS_CAP elements in its slice of shared memory.
//  * nodes of more than KS_SMALL elements are split by the whole warp, depth first (lane-strided passes, ballots);
//  * nodes of at most KS_SMALL elements are only LISTED by that phase and then finished one per LANE: every lane runs
//    nanoflann's own sequential algorithm (computeMinMax, middleSplit_, the two-pointer loops of planeSplit) on its
//    node's slice of the shared records.  A warp-wide pass over <= 32 elements is one instruction of data work under
//    ~400 instructions of warp-uniform bookkeeping; per lane the bookkeeping runs for 32 nodes at once.
// A node is written complete by whoever splits it: with lim1 <= idx <= lim2 the left child holds an element == cut
// exactly when idx > lim1 (else its largest coordinate is the largest one < cut), and the same on the right, so
// divlow / divhigh (divideTree :1080-1081) need no report from the children.  The ids of a node's children follow from
// the split position ( every position between two elements of the subtree is cut exactly once ): no id counter.
// Depth of such a subtree: every split halves the loose box along its longest side, 36 splits reduce a 4096^3 box to
// a single lattice point, and from there on cutval == tmin == tmax gives lim1 = 0, lim2 = n, idx = n / 2 — so at most
// 36 + log2( KS_CAP ) levels, and the stack of pending right children is never deeper than that.
// ---------------------------------------------------------------------------------------------------
constexpr int KS_SMALL      = 32;                   // largest node finished by a single lane
constexpr int KS_LEAF       = 10;                   // leaf_max_size (PCCKdTree.cpp:58)
constexpr int KS_LSTACK     = 3;                    // pending nodes of one lane: disjoint, > KS_LEAF elements each, the node is <= 43

struct KsItem {
  uint32_t gid;          // node id
  uint16_t left, right;  // element range inside the subtree
  int16_t  lo[3], hi[3]; // loose box (divideTree's bbox argument)
  uint16_t depth, pad;
};
static_assert( sizeof( KsItem ) == 24, "KsItem is copied as 6 words" );

template <int CAP>
struct KsWarpT {
  uint64_t rec[CAP];
  uint64_t tmp[CAP];  // warp phase: staging of the Hoare passes; lane phase: the lanes' stacks (word-interleaved)
  KsItem   stack[KS_STACK];
  KsItem   small[CAP / ( KS_LEAF + 1 ) + 2];  // listed nodes hold more than KS_LEAF elements each
};
static_assert( KS_LSTACK * 5 * 32 * 4 <= KS_CAP * 8 && KS_CAP <= 2047 && KS_SMALL < ( KS_LSTACK + 1 ) * ( KS_LEAF + 1 ), "lane stacks live in the staging array" );

// one Hoare pass of planeSplit (:1154-1181) on positions [begin, right) of the warp's subtree: the positions before
// `lim` (node-relative) must hold the elements with kd_coord < bound (INCL: <= bound)
template <bool INCL, typename KsWarp>
__device__ __forceinline__ void ks_hoare( KsWarp& S, uint32_t left, uint32_t begin, uint32_t right, uint32_t lim, int cf, int bound,
                                          int lane, uint32_t lt ) {
  // misplaced elements to the staging array at their rank: left ones at [begin, ...), right ones at [left + lim, ...)
  uint32_t ml = 0, mr = 0;
  for ( uint32_t p0 = begin; p0 < right; p0 += 32 ) {
    const uint32_t p   = p0 + lane;
    const bool     ok  = p < right;
    const uint64_t r   = ok ? S.rec[p] : 0ull;
    const int      v   = kd_coord( r, cf );
    const bool     in  = INCL ? v <= bound : v < bound;
    const bool     isL = ok && p - left < lim && !in, isR = ok && p - left >= lim && in;
    const uint32_t bl = __ballot_sync( 0xFFFFFFFFu, isL ), br = __ballot_sync( 0xFFFFFFFFu, isR );
    if ( isL ) { S.tmp[begin + ml + __popc( bl & lt )] = r; }
    if ( isR ) { S.tmp[left + lim + mr + __popc( br & lt )] = r; }
    ml += __popc( bl ), mr += __popc( br );
  }
  const uint32_t m = ml;
  __syncwarp();
  if ( m == 0 ) { return; }
  // the i-th misplaced element from the left meets the i-th misplaced element from the right
  ml = mr = 0;
  for ( uint32_t p0 = begin; p0 < right; p0 += 32 ) {
    const uint32_t p   = p0 + lane;
    const bool     ok  = p < right;
    const int      v   = kd_coord( ok ? S.rec[p] : 0ull, cf );
    const bool     in  = INCL ? v <= bound : v < bound;
    const bool     isL = ok && p - left < lim && !in, isR = ok && p - left >= lim && in;
    const uint32_t bl = __ballot_sync( 0xFFFFFFFFu, isL ), br = __ballot_sync( 0xFFFFFFFFu, isR );
    if ( isL ) { S.rec[p] = S.tmp[left + lim + ( m - 1u - ( ml + __popc( bl & lt ) ) )]; }
    if ( isR ) { S.rec[p] = S.tmp[begin + ( m - 1u - ( mr + __popc( br & lt ) ) )]; }
    ml += __popc( bl ), mr += __popc( br );
  }
  __syncwarp();
}

// lane phase: this lane finishes the node `it` (more than KS_LEAF, at most KS_SMALL elements) and everything below it.
// `stk` is the lane's stack: word w of entry d at stk[( d * 5 + w ) * 32].  Returns the depth of the deepest node.
template <typename KsWarp>
__device__ __forceinline__ int ks_lane_subtree( KsWarp& S, KdNode* __restrict__ nodes, uint32_t base, uint32_t idBase, KsItem it,
                                                bool have, uint32_t* __restrict__ stk, const int o[3] ) {
  uint32_t gid = it.gid;
  int      l = it.left, r = it.right, depth = it.depth, deepest = 0, sp = 0;
  int      lo0 = it.lo[0], lo1 = it.lo[1], lo2 = it.lo[2], hi0 = it.hi[0], hi1 = it.hi[1], hi2 = it.hi[2];
  while ( have ) {
    const int n = r - l;
    deepest     = max( deepest, depth );
    // computeMinMax (:1092-1101)
    int mn[3] = {4096, 4096, 4096}, mx[3] = {-1, -1, -1};
    for ( int p = l; p < r; p++ ) {
      const uint64_t q = S.rec[p];
      const int      x = kd_coord( q, 0 ), y = kd_coord( q, 1 ), z = kd_coord( q, 2 );
      mn[0] = min( mn[0], x ), mn[1] = min( mn[1], y ), mn[2] = min( mn[2], z );
      mx[0] = max( mx[0], x ), mx[1] = max( mx[1], y ), mx[2] = max( mx[2], z );
    }
    int cf, cut;
    {
      const int lo[3] = {lo0, lo1, lo2}, hi[3] = {hi0, hi1, hi2};
      kd_choose_split( lo, hi, mn, mx, o, cf, cut );
    }
    const int sh = 12 * cf;
    // planeSplit (:1154-1181), both pointers stepped together: a pointer that may move does, two blocked pointers swap.
    // (`right &&` of the original only keeps an unsigned index from wrapping; the indices are signed here.)
    int a = l, z = r - 1, maxLT = -1, minGT = 4096;
    while ( a <= z ) {
      const uint64_t ra = S.rec[a], rz = S.rec[z];
      const int      va = (int)( ra >> sh ) & 0xFFF, vz = (int)( rz >> sh ) & 0xFFF;
      const bool     okA = va < cut, okZ = vz >= cut;
      if ( okA ) { maxLT = max( maxLT, va ); }
      if ( !okZ ) { maxLT = max( maxLT, vz ); }
      if ( !okA && !okZ ) {
        S.rec[a] = rz, S.rec[z] = ra;
        a++, z--;
      } else {
        a += okA, z -= okZ;
      }
    }
    const int lim1 = a - l;
    z              = r - 1;
    while ( a <= z ) {
      const uint64_t ra = S.rec[a], rz = S.rec[z];
      const int      va = (int)( ra >> sh ) & 0xFFF, vz = (int)( rz >> sh ) & 0xFFF;
      const bool     okA = va <= cut, okZ = vz > cut;
      if ( !okA ) { minGT = min( minGT, va ); }
      if ( okZ ) { minGT = min( minGT, vz ); }
      if ( !okA && !okZ ) {
        S.rec[a] = rz, S.rec[z] = ra;
        a++, z--;
      } else {
        a += okA, z -= okZ;
      }
    }
    const int lim2 = a - l;
    const int idx  = (int)kd_split_index( (uint32_t)n, (uint32_t)lim1, (uint32_t)lim2 );
    const int s    = l + idx;  // split position inside the subtree, 1 <= s < total
    const uint32_t c1     = idBase + 2u * (uint32_t)( s - 1 );
    const int      divlow = idx > lim1 ? cut : maxLT, divhigh = idx < lim2 ? cut : minGT;
    *reinterpret_cast<uint4*>( &nodes[gid] ) = make_uint4( c1, (uint32_t)cf, ( (uint32_t)divlow & 0xFFFFu ) | ( (uint32_t)divhigh << 16 ), 0u );
    const int nl = idx, nr = n - idx;
    if ( nl <= KS_LEAF ) { *reinterpret_cast<uint2*>( &nodes[c1] ) = make_uint2( base + (uint32_t)l, KD_LEAF | (uint32_t)nl ); }
    if ( nr <= KS_LEAF ) { *reinterpret_cast<uint2*>( &nodes[c1 + 1] ) = make_uint2( base + (uint32_t)s, KD_LEAF | (uint32_t)nr ); }
    depth++;
    if ( nl > KS_LEAF ) {
      if ( nr > KS_LEAF && sp < KS_LSTACK ) {  // right child: [s, r), box with low[cf] = cut (sp < KS_LSTACK always holds, see above)
        uint32_t* e = stk + sp * 5 * 32;
        e[0]        = c1 + 1;
        e[32]       = (uint32_t)s | ( (uint32_t)r << 11 ) | ( (uint32_t)depth << 22 );
        e[64]       = (uint32_t)( cf == 0 ? cut : lo0 ) | ( (uint32_t)( cf == 1 ? cut : lo1 ) << 16 );
        e[96]       = (uint32_t)( cf == 2 ? cut : lo2 ) | ( (uint32_t)hi0 << 16 );
        e[128]      = (uint32_t)hi1 | ( (uint32_t)hi2 << 16 );
        sp++;
      }
      gid = c1, r = s;  // left child: [l, s), box with high[cf] = cut
      hi0 = cf == 0 ? cut : hi0, hi1 = cf == 1 ? cut : hi1, hi2 = cf == 2 ? cut : hi2;
    } else if ( nr > KS_LEAF ) {
      gid = c1 + 1, l = s;
      lo0 = cf == 0 ? cut : lo0, lo1 = cf == 1 ? cut : lo1, lo2 = cf == 2 ? cut : lo2;
    } else if ( sp > 0 ) {
      sp--;
      const uint32_t* e = stk + sp * 5 * 32;
      gid               = e[0];
      l = (int)( e[32] & 0x7FFu ), r = (int)( ( e[32] >> 11 ) & 0x7FFu ), depth = (int)( e[32] >> 22 );
      lo0 = (int)( e[64] & 0xFFFFu ), lo1 = (int)( e[64] >> 16 ), lo2 = (int)( e[96] & 0xFFFFu );
      hi0 = (int)( e[96] >> 16 ), hi1 = (int)( e[128] & 0xFFFFu ), hi2 = (int)( e[128] >> 16 );
    } else {
      have = false;
    }
  }
  return deepest;
}

template <int CAP>
__global__ void __launch_bounds__( KS_WARPS * 32 ) k_kd_subtree( uint64_t* __restrict__ grec, const GNode* __restrict__ gnodes,
                                                                 KdNode* __restrict__ nodes, const uint32_t* __restrict__ smallRoots,
                                                                 uint32_t nRoots, uint32_t* __restrict__ counters, uint32_t subBase,
                                                                 int ox, int oy, int oz, uint32_t ksSmall ) {
  extern __shared__ __align__( 16 ) unsigned char ks_smem[];
  const int      lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  using KsWarp        = KsWarpT<CAP>;
  KsWarp&        S    = reinterpret_cast<KsWarp*>( ks_smem )[w];
  const int      o[3] = {ox, oy, oz};
  const uint32_t lt   = lanemask_lt();
  int            maxDepth = 0;
  bool           overflow = false;
  for ( ;; ) {  // persistent warp: next subtree from the global counter
    uint32_t job = 0;
    if ( lane == 0 ) { job = atomicAdd( &counters[C_WORK], 1u ); }
    job = __shfl_sync( 0xFFFFFFFFu, job, 0 );
    if ( job >= nRoots ) { break; }
    const uint32_t rootId = smallRoots[job];
    const GNode&   root   = gnodes[rootId];
    const uint32_t base = root.left, total = root.right - base;
    // node ids of this subtree: a contiguous block behind the level-phase nodes, the children of the node that cuts
    // between elements s - 1 and s are idBase + 2 ( s - 1 ) and the id after it
    const uint32_t idBase = subBase + 2u * base;
    for ( uint32_t i = lane; i < total; i += 32 ) { S.rec[i] = grec[base + i]; }
    KsItem cur{};
    cur.gid = rootId, cur.left = 0, cur.right = (uint16_t)total, cur.depth = 0;
    for ( int k = 0; k < 3; k++ ) { cur.lo[k] = root.lo[k], cur.hi[k] = root.hi[k]; }
    // the root's tight box is the level phase's (k_g_setup)
    int mn[3] = {root.tmin[0], root.tmin[1], root.tmin[2]}, mx[3] = {root.tmax[0], root.tmax[1], root.tmax[2]};
    int sp = 0, nSmall = 0;
    bool haveBox = true;
    __syncwarp();
    if ( total <= (uint32_t)KS_LEAF ) {  // only the root of a subtree can arrive here as a leaf
      if ( lane == 0 ) { *reinterpret_cast<uint2*>( &nodes[rootId] ) = make_uint2( base, KD_LEAF | total ); }
    } else if ( total <= ksSmall ) {
      if ( lane == 0 ) { S.small[0] = cur; }
      nSmall = 1;
    } else {
      // ---- warp phase: nodes of more than KS_SMALL elements ----
      for ( ;; ) {
        const uint32_t left = cur.left, right = cur.right, n = right - left;
        // ---- tight box (computeMinMax) ----
        if ( !haveBox ) {
          mn[0] = mn[1] = mn[2] = 4096, mx[0] = mx[1] = mx[2] = -1;
          for ( uint32_t p = left + lane; p < right; p += 32 ) {
            const uint64_t q = S.rec[p];
            const int      x = kd_coord( q, 0 ), y = kd_coord( q, 1 ), z = kd_coord( q, 2 );
            mn[0] = min( mn[0], x ), mn[1] = min( mn[1], y ), mn[2] = min( mn[2], z );
            mx[0] = max( mx[0], x ), mx[1] = max( mx[1], y ), mx[2] = max( mx[2], z );
          }
#pragma unroll
          for ( int k = 0; k < 3; k++ ) {
            mn[k] = __reduce_min_sync( 0xFFFFFFFFu, mn[k] );
            mx[k] = __reduce_max_sync( 0xFFFFFFFFu, mx[k] );
          }
        }
        haveBox = false;
        int cf, cut;
        {
          const int lo[3] = {cur.lo[0], cur.lo[1], cur.lo[2]}, hi[3] = {cur.hi[0], cur.hi[1], cur.hi[2]};
          kd_choose_split( lo, hi, mn, mx, o, cf, cut );
        }
        // ---- lim1 / lim2, and the coordinates next to the cut on both sides ----
        uint32_t lim1 = 0, lim2 = 0;
        int      maxLT = -1, minGT = 4096;
        for ( uint32_t p = left + lane; p < right; p += 32 ) {
          const int v = kd_coord( S.rec[p], cf );
          lim1 += v < cut, lim2 += v <= cut;
          if ( v < cut ) { maxLT = max( maxLT, v ); }
          if ( v > cut ) { minGT = min( minGT, v ); }
        }
        lim1 = __reduce_add_sync( 0xFFFFFFFFu, lim1 );
        lim2 = __reduce_add_sync( 0xFFFFFFFFu, lim2 );
        // ---- the two Hoare passes of planeSplit ----
        if ( lim1 > 0 && lim1 < n ) { ks_hoare<false, KsWarp>( S, left, left, right, lim1, cf, cut, lane, lt ); }
        if ( lim2 > lim1 && lim2 < n ) { ks_hoare<true, KsWarp>( S, left, left + lim1, right, lim2, cf, cut, lane, lt ); }
        const uint32_t idx = kd_split_index( n, lim1, lim2 );
        const uint32_t s   = left + idx, c1 = idBase + 2u * ( s - 1u );
        int divlow = cut, divhigh = cut;
        if ( idx == lim1 ) { divlow = __reduce_max_sync( 0xFFFFFFFFu, maxLT ); }
        if ( idx == lim2 ) { divhigh = __reduce_min_sync( 0xFFFFFFFFu, minGT ); }
        if ( lane == 0 ) {
          *reinterpret_cast<uint4*>( &nodes[cur.gid] ) =
              make_uint4( c1, (uint32_t)cf, ( (uint32_t)divlow & 0xFFFFu ) | ( (uint32_t)divhigh << 16 ), 0u );
        }
        // ---- children (divideTree :1070-1078): leaves are written, small ones listed, the left large one is kept ----
        const uint32_t nl = idx, nr = n - idx;
        KsItem lft = cur, rgt = cur;
        lft.gid = c1, lft.right = (uint16_t)s, lft.depth = rgt.depth = (uint16_t)( cur.depth + 1 );
        rgt.gid = c1 + 1, rgt.left = (uint16_t)s;
        if ( cf == 0 ) {  // left_bbox[cutfeat].high = right_bbox[cutfeat].low = cutval
          lft.hi[0] = rgt.lo[0] = (int16_t)cut;
        } else if ( cf == 1 ) {
          lft.hi[1] = rgt.lo[1] = (int16_t)cut;
        } else {
          lft.hi[2] = rgt.lo[2] = (int16_t)cut;
        }
        if ( lane == 0 ) {
          if ( nl <= (uint32_t)KS_LEAF ) {
            *reinterpret_cast<uint2*>( &nodes[c1] ) = make_uint2( base + left, KD_LEAF | nl );
          } else if ( nl <= ksSmall ) {
            S.small[nSmall] = lft;
          }
        }
        if ( nl > (uint32_t)KS_LEAF && nl <= ksSmall ) { nSmall++; }
        if ( lane == 0 ) {
          if ( nr <= (uint32_t)KS_LEAF ) {
            *reinterpret_cast<uint2*>( &nodes[c1 + 1] ) = make_uint2( base + s, KD_LEAF | nr );
          } else if ( nr <= ksSmall ) {
            S.small[nSmall] = rgt;
          }
        }
        if ( nr > (uint32_t)KS_LEAF && nr <= ksSmall ) { nSmall++; }
        maxDepth = max( maxDepth, (int)cur.depth );
        const bool bigL = nl > ksSmall, bigR = nr > ksSmall;
        if ( bigL ) {
          if ( bigR ) {
            if ( sp >= KS_STACK ) {  // cannot happen for 12-bit coordinates (see above); fail loudly
              overflow = true;
              break;
            }
            if ( lane == 0 ) { S.stack[sp] = rgt; }
            sp++;
          }
          cur = lft;
        } else if ( bigR ) {
          cur = rgt;
        } else {
          if ( sp == 0 ) { break; }
          __syncwarp();
          cur = S.stack[--sp];
        }
      }
    }
    __syncwarp();
    // ---- lane phase: the listed nodes, one per lane ----
    for ( int b = 0; b < nSmall; b += 32 ) {
      const bool have = b + lane < nSmall;
      KsItem     it{};
      if ( have ) { it = S.small[b + lane]; }
      const int d = ks_lane_subtree( S, nodes, base, idBase, it, have, reinterpret_cast<uint32_t*>( S.tmp ) + lane, o );
      maxDepth    = max( maxDepth, d );
    }
    __syncwarp();
    for ( uint32_t i = lane; i < total; i += 32 ) { grec[base + i] = S.rec[i]; }
    __syncwarp();
  }
  maxDepth = __reduce_max_sync( 0xFFFFFFFFu, maxDepth );
  if ( lane == 0 ) {
    atomicMax( &counters[C_DEPTH], (uint32_t)maxDepth + 1u );
    if ( overflow ) { atomicOr( &counters[C_STACK], 1u ); }
  }
}

// level-phase nodes -> search nodes: divlow / divhigh from the children's tight boxes (divideTree :1080-1081), and the
// root boxes for computeInitialDistances
__global__ void k_g_finalize( const GNode* __restrict__ gnodes, uint32_t nLevelNodes, int nTrees, KdNode* __restrict__ nodes,
                              int16_t* __restrict__ rootBox ) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= nLevelNodes || i == 0 ) { return; }
  const GNode& n = gnodes[i];
  if ( i <= (uint32_t)nTrees ) {
    for ( int k = 0; k < 3; k++ ) { rootBox[(size_t)i * 6 + k] = n.tmin[k], rootBox[(size_t)i * 6 + 3 + k] = n.tmax[k]; }
  }
  if ( n.state != 1 ) { return; }  // roots of shared-memory subtrees wrote their own search node
  KdNode s{};
  s.a       = n.child1;
  s.b       = (uint32_t)n.cutfeat;
  s.divlow  = gnodes[n.child1].tmax[n.cutfeat];
  s.divhigh = gnodes[n.child1 + 1].tmin[n.cutfeat];
  nodes[i]  = s;
}

template <int CAP>
int launch_subtrees( rb200_ctx* c, uint64_t* rec, const GNode* gnodes, KdNode* nodes, const uint32_t* roots, uint32_t nRoots,
                     uint32_t* counters, uint32_t subBase, int ox, int oy, int oz, uint32_t ksSmall ) {
  const size_t smem = sizeof( KsWarpT<CAP> ) * KS_WARPS;
  int          perSM = 1, nSM = 148;
  RB_CUDA( cudaFuncSetAttribute( k_kd_subtree<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem ) );
  RB_CUDA( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSM, k_kd_subtree<CAP>, KS_WARPS * 32, smem ) );
  RB_CUDA( cudaDeviceGetAttribute( &nSM, cudaDevAttrMultiProcessorCount, c->device ) );
  const uint32_t grid = (uint32_t)std::min<int64_t>( (int64_t)std::max( perSM, 1 ) * nSM, ( (int64_t)nRoots + KS_WARPS - 1 ) / KS_WARPS );
  RB_LAUNCH( "kd_subtree", k_kd_subtree<CAP>, grid, KS_WARPS * 32, smem, rec, gnodes, nodes, roots, nRoots, counters, subBase, ox, oy,
             oz, ksSmall );
  return RB200_OK;
}

}  // namespace

int rb_kd_build( rb200_ctx* c, RbKdBuild& B, const short4* pos, const int64_t* dOff, const std::vector<int64_t>& hOff, int ox,
                 int oy, int oz ) {
  // measured alternatives on the 32-frame vox10 GOF (subtree kernel, ms): cap 512 / lane nodes <= 32: 3.78; lane nodes
  // <= 16 / 24 / 48 / 64: 4.38 / 3.85 / 3.93 / 4.44; cap 1024 (one level less in the level phase, half the warps): 7.15
  constexpr uint32_t ksCap = KS_CAP, ksSmall = KS_SMALL;
  const int     nTrees = (int)hOff.size() - 1;
  const int64_t E      = hOff[nTrees];
  if ( E <= 0 || nTrees <= 0 ) { return rb_fail( c, RB200_ERR_INVALID, "kd build: empty forest" ); }
  if ( E >= ( 1ll << 30 ) ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "kd build: more than 2^30 points" ); }
  for ( int t = 0; t < nTrees; t++ ) {
    if ( hOff[t + 1] - hOff[t] <= 0 ) { return rb_fail( c, RB200_ERR_INVALID, "kd build: empty cloud %d", t ); }
    if ( hOff[t + 1] - hOff[t] >= ( 1ll << 28 ) ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "kd build: cloud too large" ); }
  }
  // level-phase nodes: every split node holds more than KS_CAP elements, so a level has at most E / KS_CAP of them
  const uint32_t gCap     = (uint32_t)( E / 16 + 64ll * nTrees + 4096 );
  // chunks of a level: every node of more than ksCap elements has at most one that is not full
  const uint32_t chunkCap = (uint32_t)( E / GT + E / ksCap + 2ll * nTrees + 64 );
  const size_t   nodeCap  = (size_t)gCap + 2 * (size_t)E + 2;  // + the id blocks of the subtrees (2 ids per element)
  if ( nodeCap >= (size_t)KD_NODE_MAX ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "kd build: %lld elements need more node ids than a search stack entry holds", (long long)E );
  }
  RB_CUDA( B.rec.ensure( (size_t)E * 8 ) );
  RB_CUDA( B.tmp.ensure( (size_t)E * 8 ) );
  RB_CUDA( B.gnodes.ensure( (size_t)gCap * sizeof( GNode ) ) );
  RB_CUDA( B.nodes.ensure( nodeCap * sizeof( KdNode ) ) );
  RB_CUDA( B.rootBox.ensure( (size_t)( nTrees + 1 ) * 12 ) );
  RB_CUDA( B.chunkNode.ensure( (size_t)chunkCap * 4 ) );
  RB_CUDA( B.chunks.ensure( (size_t)chunkCap * sizeof( GChunk ) ) );
  RB_CUDA( B.cls.ensure( (size_t)gCap * 12 * 4 ) );
  RB_CUDA( B.wA.ensure( (size_t)chunkCap * GWORDS * 4 ) );
  RB_CUDA( B.wB.ensure( (size_t)chunkCap * GWORDS * 4 ) );
  RB_CUDA( B.wC.ensure( (size_t)chunkCap * GWORDS * 4 ) );
  RB_CUDA( B.cA.ensure( (size_t)chunkCap * 4 ) );
  RB_CUDA( B.cB.ensure( (size_t)chunkCap * 4 ) );
  RB_CUDA( B.smallRoots.ensure( (size_t)gCap * 4 ) );
  RB_CUDA( B.largeList.ensure( (size_t)( E / P2_WARP_MAX + nTrees + 64 ) * 4 ) );
  RB_CUDA( B.counters.ensure( 64 ) );
  RB_CUDA( B.stats.ensure( (size_t)gCap * 24 ) );
  int32_t*  st        = B.stats.as<int32_t>();
  uint64_t* rec       = B.rec.as<uint64_t>();
  uint64_t* tmp       = B.tmp.as<uint64_t>();
  GNode*    gnodes    = B.gnodes.as<GNode>();
  KdNode*   nodes     = B.nodes.as<KdNode>();
  uint32_t* counters  = B.counters.as<uint32_t>();
  uint32_t* chunkNode = B.chunkNode.as<uint32_t>();
  GChunk*   chunks    = B.chunks.as<GChunk>();
  int32_t*  cls       = B.cls.as<int32_t>();
  uint32_t *wA = B.wA.as<uint32_t>(), *wB = B.wB.as<uint32_t>(), *wC = B.wC.as<uint32_t>(), *cA = B.cA.as<uint32_t>(),
           *cB = B.cB.as<uint32_t>();
  uint32_t* h = (uint32_t*)rb_pinned( c, 64 );
  if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  for ( int k = 0; k < 16; k++ ) { h[k] = 0; }
  h[C_NEXT] = (uint32_t)nTrees + 1;
  RB_CUDA( cudaMemcpyAsync( counters, h, 64, cudaMemcpyHostToDevice, c->stream ) );
  RB_LAUNCH( "kd_roots", k_g_roots, rb_div_up( nTrees, 128 ), 128, 0, gnodes, dOff, nTrees, st );
  RB_LAUNCH( "kd_init", k_g_init, rb_div_up( E, TPB * GEPT ), TPB, 0, pos, dOff, nTrees, E, ox, oy, oz, rec, st, counters );
  uint32_t lvlBegin = 1, lvlEnd = (uint32_t)nTrees + 1;
  int      level = 0;
  for ( ;; level++ ) {
    if ( level > 200 ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "kd build: tree deeper than 200 levels" ); }
    const uint32_t nLvl = lvlEnd - lvlBegin;
    RB_CUDA( cudaMemsetAsync( counters + C_BIG, 0, 4, c->stream ) );
    RB_CUDA( cudaMemsetAsync( counters + C_CHUNKS, 0, 4, c->stream ) );
    RB_CUDA( cudaMemsetAsync( counters + C_LARGE, 0, 4, c->stream ) );
    RB_LAUNCH( "kd_setup", k_g_setup, rb_div_up( nLvl, TPB ), TPB, 0, gnodes, st, lvlBegin, lvlEnd, level == 0 ? 1 : 0, ox, oy, oz,
               B.smallRoots.as<uint32_t>(), counters, chunkNode, chunkCap, gCap, cls, B.largeList.as<uint32_t>(), ksCap );
    RB_CUDA( cudaMemcpyAsync( h, counters, 64, cudaMemcpyDeviceToHost, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
    const uint32_t nLarge = h[C_LARGE];
    if ( h[C_RANGE] ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "kd build: coordinate range of the clouds exceeds 4096" ); }
    if ( h[C_POOL] ) { return rb_fail( c, RB200_ERR_NOMEM, "kd build: node pool exhausted" ); }
    const uint32_t nBig = h[C_BIG], nChunks = h[C_CHUNKS];
    if ( nBig == 0 ) { break; }  // no node of this level is large enough for the level phase
    const int GW = rb_div_up( (int64_t)nLvl * 32, TPB );
    RB_LAUNCH( "kd_count", k_g_count, nChunks, TPB, 0, rec, gnodes, chunkNode, chunks, wA, wB, cA, cB, cls );
    RB_LAUNCH( "kd_nodescan", k_g_nodescan<false>, GW, TPB, 0, gnodes, lvlBegin, lvlEnd, chunks, cA, cB, wA, wB, st, cls );
    RB_LAUNCH( "kd_stage1", k_g_stage<false>, nChunks, TPB, 0, rec, chunks, wA, tmp );
    RB_LAUNCH( "kd_apply1", k_g_apply1, nChunks, TPB, 0, rec, chunks, wA, wB, wC, cB, tmp );
    if ( nLarge ) { RB_LAUNCH( "kd_pass2", k_g_pass2_node, nLarge, TPB, 0, rec, gnodes, B.largeList.as<uint32_t>(), wC, st ); }
    RB_LAUNCH( "kd_pass2", k_g_pass2_warp, rb_div_up( nLvl, TPB / 32 ), TPB, 0, rec, gnodes, lvlBegin, lvlEnd, wC, st );
    lvlBegin = lvlEnd;  // the children were numbered consecutively behind the nodes that existed
    lvlEnd   = h[C_NEXT];
  }
  const uint32_t nRoots = h[C_SMALL], nLevelNodes = h[C_NEXT];  // nodes [1, nLevelNodes) were created by the level phase
  if ( nRoots ) {
    int r = launch_subtrees<KS_CAP>( c, rec, gnodes, nodes, B.smallRoots.as<uint32_t>(), nRoots, counters, gCap, ox, oy, oz, ksSmall );
    if ( r ) { return r; }
  }
  RB_LAUNCH( "kd_finalize", k_g_finalize, rb_div_up( nLevelNodes, TPB ), TPB, 0, gnodes, nLevelNodes, nTrees, nodes,
             B.rootBox.as<int16_t>() );
  RB_CUDA( cudaMemcpyAsync( h, counters, 64, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  if ( h[C_STACK] ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "kd build: a subtree is deeper than its work stack" ); }
  if ( level + (int)h[C_DEPTH] + 2 >= KD_STACK ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "kd build: tree depth %d exceeds the traversal stack", level + (int)h[C_DEPTH] );
  }
  B.forest.rec     = rec;
  B.forest.nodes   = nodes;
  B.forest.rootBox = B.rootBox.as<int16_t>();
  B.forest.ox      = ox;
  B.forest.oy      = oy;
  B.forest.oz      = oz;
  B.nNodes         = nLevelNodes;
  B.nTrees         = nTrees;
  B.levels         = level + (int)h[C_DEPTH];
  return RB200_OK;
}
