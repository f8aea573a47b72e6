// rb_kdtree.cu — builds, on the GPU, exactly the kd-trees nanoflann would build on the CPU (see rb_kdtree.cuh).
//
// Restates KDTreeSingleIndexAdaptor::buildIndex / computeBoundingBox / divideTree / middleSplit_ / planeSplit /
// computeMinMax (dependencies/nanoflann/nanoflann.hpp:858-866, 1009-1181) for a FOREST of trees (one per cloud of a
// GOF) in two phases:
//  * level-parallel phase for nodes with more than KB_CAP (2048) points: every level is a handful of passes over the
//    element records of all trees (per-node min/max and counts with warp-aggregated atomics, then the two Hoare
//    passes of planeSplit as "rank the misplaced elements with one prefix sum, swap the i-th misplaced element from
//    the left with the i-th misplaced element from the right");
//  * block phase for subtrees of at most KB_CAP points: one CTA per subtree runs the same level-synchronous
//    formulation entirely in shared memory, down to the leaves.
// divlow / divhigh come from the children's tight boxes (divideTree :1080-1081).
#include <algorithm>

#include "rb_common.cuh"
#include "rb_kdtree.cuh"
#include "rb_kdtree_build.cuh"

namespace {

constexpr int TPB      = 256;

// per-node {min[3], max[3]} of the level-parallel phase: plain int32 so the updates are single RED instructions
__device__ __forceinline__ void stat_update( int32_t* __restrict__ st, uint32_t node, const int mn[3], const int mx[3] ) {
  int32_t* p = st + (size_t)node * 6;
#pragma unroll
  for ( int k = 0; k < 3; k++ ) {
    atomicMin( p + k, mn[k] );
    atomicMax( p + 3 + k, mx[k] );
  }
}

// records from positions (all clouds of the forest are concatenated; tree t owns [off[t], off[t+1]))
__global__ void k_kd_init( const short4* __restrict__ pos, const int64_t* __restrict__ off, int nTrees, int64_t E, int ox, int oy,
                           int oz, uint64_t* __restrict__ rec, uint32_t* __restrict__ nid, uint32_t* __restrict__ err ) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( e >= E ) { return; }
  int lo = 0, hi = nTrees - 1;
  while ( lo < hi ) {
    const int mid = ( lo + hi + 1 ) >> 1;
    if ( off[mid] <= e ) {
      lo = mid;
    } else {
      hi = mid - 1;
    }
  }
  const short4 p = pos[e];
  const int    x = p.x - ox, y = p.y - oy, z = p.z - oz;
  if ( (unsigned)x > 4095u || (unsigned)y > 4095u || (unsigned)z > 4095u ) { atomicOr( err, 1u ); }
  rec[e] = (uint64_t)( x & 0xFFF ) | ( (uint64_t)( y & 0xFFF ) << 12 ) | ( (uint64_t)( z & 0xFFF ) << 24 ) |
           ( (uint64_t)( e - off[lo] ) << 36 );
  nid[e] = (uint32_t)lo + 1u;
}

__global__ void k_kd_roots( KdNode* __restrict__ nodes, const int64_t* __restrict__ off, int nTrees, int32_t* __restrict__ st ) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if ( t >= nTrees ) { return; }
  for ( int k = 0; k < 3; k++ ) {
    st[(size_t)( t + 1 ) * 6 + k]     = 0x7FFFFFFF;
    st[(size_t)( t + 1 ) * 6 + 3 + k] = (int32_t)0x80000000;
  }
  KdNode n{};
  n.left  = (uint32_t)off[t];
  n.right = (uint32_t)off[t + 1];
  for ( int k = 0; k < 3; k++ ) {
    n.tmin[k] = 32767;
    n.tmax[k] = -32768;
  }
  n.state      = 0;
  nodes[t + 1] = n;
  if ( t == 0 ) {
    KdNode z{};
    nodes[0] = z;
  }
}

// per-node tight bounding box of the nodes of this level (computeMinMax for all three axes at once):
// warp shuffle reduction -> CTA combine in shared memory -> one RED per CTA and node in the common case
// ASSIGN: the elements of the split nodes of level [lvlBegin, lvlEnd) first move to their child (what k_kd_assign
// does), and the boxes are those of the children — the statistics pass of the next level rides on the assignment pass
template <bool ASSIGN>
__global__ void __launch_bounds__( TPB ) k_kd_stats( const uint64_t* __restrict__ rec, uint32_t* __restrict__ nid,
                                                     const KdNode* __restrict__ nodes, int32_t* __restrict__ st, int64_t E,
                                                     uint32_t lvlBegin, uint32_t lvlEnd ) {
  __shared__ uint32_t wNode[TPB / 32];
  __shared__ int      wMin[TPB / 32][3], wMax[TPB / 32][3];
  const int64_t e    = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int     lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t      node = 0;
  int           c[3] = {0, 0, 0};
  bool          on   = false;
  if ( e < E ) {
    node = nid[e];
    on   = node >= lvlBegin && node < lvlEnd;
    if ( ASSIGN && on ) {
      const KdNode& n = nodes[node];
      on              = n.state == 1 && n.child1 != 0;
      if ( on ) {
        node   = ( (uint32_t)e < nodes[n.child1].right ) ? n.child1 : n.child1 + 1;
        nid[e] = node;
      }
    }
    if ( on ) {
      const uint64_t r = rec[e];
      c[0] = kd_coord( r, 0 ), c[1] = kd_coord( r, 1 ), c[2] = kd_coord( r, 2 );
    }
  }
  const uint32_t act     = __ballot_sync( 0xFFFFFFFFu, on );
  const uint32_t peers   = on ? __match_any_sync( act, node ) : 0u;
  const bool     uniform = on && peers == act && act == 0xFFFFFFFFu;  // warp-uniform predicate (all lanes agree)
  const bool     wuni    = __all_sync( 0xFFFFFFFFu, uniform );
  int            mn[3] = {c[0], c[1], c[2]}, mx[3] = {c[0], c[1], c[2]};
  if ( wuni ) {
#pragma unroll
    for ( int k = 0; k < 3; k++ ) {
#pragma unroll
      for ( int d = 16; d > 0; d >>= 1 ) {
        mn[k] = min( mn[k], __shfl_xor_sync( 0xFFFFFFFFu, mn[k], d ) );
        mx[k] = max( mx[k], __shfl_xor_sync( 0xFFFFFFFFu, mx[k], d ) );
      }
    }
    if ( lane == 0 ) {
      wNode[w] = node;
      for ( int k = 0; k < 3; k++ ) { wMin[w][k] = mn[k], wMax[w][k] = mx[k]; }
    }
  } else {
    if ( lane == 0 ) { wNode[w] = 0xFFFFFFFFu; }
    if ( on ) {  // mixed warp: reduce inside every peer group through its leader
      const int leader = __ffs( peers ) - 1;
      for ( uint32_t m = peers & ~( 1u << leader ); m; m &= m - 1 ) {
        const int src = __ffs( m ) - 1;
#pragma unroll
        for ( int k = 0; k < 3; k++ ) {
          const int v = __shfl_sync( peers, c[k], src );
          mn[k] = min( mn[k], v );
          mx[k] = max( mx[k], v );
        }
      }
      if ( lane == leader ) { stat_update( st, node, mn, mx ); }
    }
  }
  __syncthreads();
  if ( threadIdx.x < TPB / 32 ) {  // combine the uniform warps of this CTA: a run of equal nodes is flushed by its first warp
    const int      i  = threadIdx.x;
    const uint32_t nd = wNode[i];
    if ( nd != 0xFFFFFFFFu && ( i == 0 || wNode[i - 1] != nd ) ) {
      int a[3] = {wMin[i][0], wMin[i][1], wMin[i][2]}, b[3] = {wMax[i][0], wMax[i][1], wMax[i][2]};
      for ( int j = i + 1; j < TPB / 32 && wNode[j] == nd; j++ ) {
        for ( int k = 0; k < 3; k++ ) {
          a[k] = min( a[k], wMin[j][k] );
          b[k] = max( b[k], wMax[j][k] );
        }
      }
      stat_update( st, nd, a, b );
    }
  }
}

// middleSplit_ (nanoflann.hpp:1103-1142) for one node: cut axis and cut value
__device__ __forceinline__ void kd_choose_split( KdNode& n, int o[3] ) {
  int max_span = n.hi[0] - n.lo[0];
  for ( int i = 1; i < 3; i++ ) { max_span = max( max_span, n.hi[i] - n.lo[i] ); }
  int cutfeat = 0, max_spread = -1;
  for ( int i = 0; i < 3; i++ ) {
    const int span = n.hi[i] - n.lo[i];
    if ( (double)span > ( 1.0 - 0.00001 ) * (double)max_span ) {  // span > (1 - EPS) * max_span in double
      const int spread = n.tmax[i] - n.tmin[i];
      if ( spread > max_spread ) {
        cutfeat    = i;
        max_spread = spread;
      }
    }
  }
  // split_val = (bbox.low + bbox.high) / 2 is an int division of the ABSOLUTE coordinates (truncation toward zero)
  const int split = ( ( n.lo[cutfeat] + o[cutfeat] ) + ( n.hi[cutfeat] + o[cutfeat] ) ) / 2 - o[cutfeat];
  int       cut;
  if ( split < n.tmin[cutfeat] ) {
    cut = n.tmin[cutfeat];
  } else if ( split > n.tmax[cutfeat] ) {
    cut = n.tmax[cutfeat];
  } else {
    cut = split;
  }
  n.cutfeat = (int8_t)cutfeat;
  n.cutval  = (int16_t)cut;
}

// the nodes of this level: leaf / small root / split decision
__global__ void k_kd_split( KdNode* __restrict__ nodes, uint32_t lvlBegin, uint32_t lvlEnd, int small, int ox, int oy, int oz,
                            uint32_t* __restrict__ smallRoots, uint32_t* __restrict__ counters, int isRoot,
                            const int32_t* __restrict__ st ) {
  const uint32_t i = lvlBegin + blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= lvlEnd ) { return; }
  KdNode& n = nodes[i];
  for ( int k = 0; k < 3; k++ ) {
    n.tmin[k] = (int16_t)st[(size_t)i * 6 + k];
    n.tmax[k] = (int16_t)st[(size_t)i * 6 + 3 + k];
  }
  if ( isRoot ) {  // a root: divideTree( 0, N, root_bbox ) starts from the tight box
    for ( int k = 0; k < 3; k++ ) {
      n.lo[k] = n.tmin[k];
      n.hi[k] = n.tmax[k];
    }
  }
  const uint32_t count = n.right - n.left;
  if ( count <= (uint32_t)small ) {
    n.state                                = 2;
    smallRoots[atomicAdd( &counters[1], 1u )] = i;
    return;
  }
  int o[3] = {ox, oy, oz};
  kd_choose_split( n, o );
  n.lt = n.le = 0;
  n.state     = 1;
  atomicAdd( &counters[2], 1u );  // big nodes of this level
}

// lim1 / lim2 of planeSplit: elements < cutval and <= cutval
__global__ void k_kd_count( const uint64_t* __restrict__ rec, const uint32_t* __restrict__ nid, KdNode* __restrict__ nodes,
                            int64_t E, uint32_t lvlBegin, uint32_t lvlEnd ) {
  const int64_t e    = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  uint32_t      node = 0;
  bool          on = false, lt = false, le = false;
  if ( e < E ) {
    node = nid[e];
    if ( node >= lvlBegin && node < lvlEnd ) {
      const KdNode& n = nodes[node];
      if ( n.state == 1 ) {
        on          = true;
        const int v = kd_coord( rec[e], n.cutfeat );
        lt          = v < n.cutval;
        le          = v <= n.cutval;
      }
    }
  }
  const uint32_t act = __ballot_sync( 0xFFFFFFFFu, on );
  if ( !on ) { return; }
  const uint32_t peers = __match_any_sync( act, node );
  const uint32_t bl = __ballot_sync( act, lt ), be = __ballot_sync( act, le );
  if ( ( threadIdx.x & 31 ) == __ffs( peers ) - 1 ) {
    const uint32_t a = __popc( bl & peers ), b = __popc( be & peers );
    if ( a ) { atomicAdd( &nodes[node].lt, a ); }
    if ( b ) { atomicAdd( &nodes[node].le, b ); }
  }
}

// misplaced-element flags of one Hoare pass.  pass 0: [0, count) split at lim1 by (v < cutval);
// pass 1: [lim1, count) split at lim2 by (v <= cutval)
__global__ void k_kd_flag( const uint64_t* __restrict__ rec, const uint32_t* __restrict__ nid, const KdNode* __restrict__ nodes,
                           int64_t E, uint32_t lvlBegin, uint32_t lvlEnd, int pass, uint32_t* __restrict__ flags ) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( e > E ) { return; }
  uint32_t f = 0;
  if ( e < E ) {
    const uint32_t node = nid[e];
    if ( node >= lvlBegin && node < lvlEnd ) {
      const KdNode& n = nodes[node];
      if ( n.state == 1 ) {
        const uint32_t p = (uint32_t)e - n.left;
        const int      v = kd_coord( rec[e], n.cutfeat );
        if ( pass == 0 ) {
          const bool in = v < n.cutval;
          f             = ( p < n.lt ) ? !in : in;
        } else if ( p >= n.lt ) {
          const bool in = v <= n.cutval;
          f             = ( p < n.le ) ? !in : in;
        }
      }
    }
  }
  flags[e] = f;
}

// pair lists: the i-th misplaced element from the left (ascending position) meets the i-th misplaced element
// from the right (descending position) — exactly the swaps of the two-pointer loop (:1154-1181)
__global__ void k_kd_pairs( const uint32_t* __restrict__ nid, const KdNode* __restrict__ nodes, int64_t E, uint32_t lvlBegin,
                            uint32_t lvlEnd, int pass, const uint32_t* __restrict__ scan, const uint32_t* __restrict__ flags_unused,
                            uint32_t* __restrict__ pairL, uint32_t* __restrict__ pairR ) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( e >= E ) { return; }
  if ( scan[e + 1] == scan[e] ) { return; }
  const KdNode&  n     = nodes[nid[e]];
  const uint32_t begin = pass == 0 ? n.left : n.left + n.lt;
  const uint32_t lim   = pass == 0 ? n.left + n.lt : n.left + n.le;
  const uint32_t m     = ( scan[n.right] - scan[begin] ) >> 1;  // misplaced on each side
  const uint32_t r     = scan[e] - scan[begin];
  if ( (uint32_t)e < lim ) {
    pairL[begin + r] = (uint32_t)e;
  } else {
    pairR[begin + ( m - 1 - ( r - m ) )] = (uint32_t)e;
  }
}

__global__ void k_kd_swap( uint64_t* __restrict__ rec, const uint32_t* __restrict__ nid, const KdNode* __restrict__ nodes, int64_t E,
                           uint32_t lvlBegin, uint32_t lvlEnd, int pass, const uint32_t* __restrict__ scan,
                           const uint32_t* __restrict__ pairL, const uint32_t* __restrict__ pairR ) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( e >= E ) { return; }
  const uint32_t node = nid[e];
  if ( node < lvlBegin || node >= lvlEnd ) { return; }
  const KdNode& n = nodes[node];
  if ( n.state != 1 ) { return; }
  const uint32_t begin = pass == 0 ? n.left : n.left + n.lt;
  if ( (uint32_t)e < begin ) { return; }
  const uint32_t m = ( scan[n.right] - scan[begin] ) >> 1;
  const uint32_t k = (uint32_t)e - begin;
  if ( k >= m ) { return; }
  const uint32_t a = pairL[begin + k], b = pairR[begin + k];
  const uint64_t ra = rec[a], rb = rec[b];
  rec[a] = rb;
  rec[b] = ra;
}

// children of the split nodes of this level (divideTree :1070-1078)
__global__ void k_kd_children( KdNode* __restrict__ nodes, uint32_t lvlBegin, uint32_t lvlEnd, uint32_t* __restrict__ counters,
                               uint32_t nodeCap, int32_t* __restrict__ st ) {
  const uint32_t i = lvlBegin + blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= lvlEnd ) { return; }
  KdNode& n = nodes[i];
  if ( n.state != 1 ) { return; }
  const uint32_t count = n.right - n.left;
  uint32_t       idx;  // :1137-1139
  if ( n.lt > count / 2 ) {
    idx = n.lt;
  } else if ( n.le < count / 2 ) {
    idx = n.le;
  } else {
    idx = count / 2;
  }
  const uint32_t c1 = atomicAdd( &counters[0], 2u );
  if ( c1 + 2 > nodeCap ) {
    counters[3] = 1;  // node pool exhausted
    n.state     = 3;
    n.child1    = 0;
    return;
  }
  n.child1 = c1;
  KdNode a{}, b{};
  a.left  = n.left;
  a.right = n.left + idx;
  b.left  = n.left + idx;
  b.right = n.right;
  for ( int k = 0; k < 3; k++ ) {
    a.lo[k] = b.lo[k] = n.lo[k];
    a.hi[k] = b.hi[k] = n.hi[k];
    a.tmin[k] = b.tmin[k] = 32767;
    a.tmax[k] = b.tmax[k] = -32768;
  }
  a.hi[n.cutfeat] = n.cutval;  // left_bbox[cutfeat].high = cutval
  b.lo[n.cutfeat] = n.cutval;  // right_bbox[cutfeat].low = cutval
  nodes[c1]       = a;
  nodes[c1 + 1]   = b;
  for ( int k = 0; k < 3; k++ ) {
    st[(size_t)c1 * 6 + k] = st[(size_t)( c1 + 1 ) * 6 + k] = 0x7FFFFFFF;
    st[(size_t)c1 * 6 + 3 + k] = st[(size_t)( c1 + 1 ) * 6 + 3 + k] = (int32_t)0x80000000;
  }
}

// ---------------------------------------------------------------------------------------------------
// block phase: one CTA builds a whole subtree of at most KB_CAP elements, level by level, in shared memory.
// Same formulation as the level-parallel phase (tight boxes with atomics, counts, the two Hoare passes as
// rank-the-misplaced + pairwise swap), but every pass is a few hundred cycles on 2048 shared-memory records.
// ---------------------------------------------------------------------------------------------------
constexpr int KB_CAP   = 2048;  // elements per CTA subtree
constexpr int KB_LEVEL = 384;   // nodes per level: children only come from nodes with > 10 elements, so <= 2 * 2048 / 11 = 372
constexpr int KB_TPB   = 256;
constexpr int KB_EPT   = KB_CAP / KB_TPB;

struct KbNode {  // a node of the level being processed
  int32_t  tmin[3], tmax[3];
  uint32_t gid, pgid, lt, le;
  uint16_t left, right, child;  // element range inside the CTA subtree; child: index of child1 in the next level
  int16_t  lo[3], hi[3], cutval;
  uint8_t  cutfeat, pfeat, state, side;  // side: 0 left child, 1 right child of pgid (pgid == 0: subtree root)
};
struct KbNext {  // a node of the next level, as created by its parent
  uint32_t gid, pgid;
  uint16_t left, right;
  int16_t  lo[3], hi[3];
  uint8_t  pfeat, side;
};

struct KbShared {
  uint64_t rec[KB_CAP];
  KbNode   cur[KB_LEVEL];
  KbNext   nxt[KB_LEVEL];
  uint16_t nid[KB_CAP];
  uint16_t scan[KB_CAP + 2];
  uint16_t pairL[KB_CAP], pairR[KB_CAP];
  uint32_t warpSum[KB_TPB / 32];
  uint32_t nNext, gBase, big;
  uint8_t  active[KB_CAP / 32];  // chunk of 32 consecutive elements still has an element in an unfinished node
};

__global__ void __launch_bounds__( KB_TPB ) k_kd_block( uint64_t* __restrict__ grec, KdNode* __restrict__ nodes,
                                                        const uint32_t* __restrict__ smallRoots, uint32_t nRoots,
                                                        uint32_t* __restrict__ counters, uint32_t nodeCap, int ox, int oy, int oz ) {
  extern __shared__ __align__( 16 ) unsigned char kb_smem[];
  KbShared&  S     = *reinterpret_cast<KbShared*>( kb_smem );
  uint64_t*  rec   = S.rec;
  uint16_t*  nid   = S.nid;
  uint16_t*  scan  = S.scan;
  uint16_t*  pairL = S.pairL;
  uint16_t*  pairR = S.pairR;
  KbNode*    cur   = S.cur;
  KbNext*    nxt   = S.nxt;
  uint32_t*  warpSum = S.warpSum;
  uint32_t&  sNNext = S.nNext;
  uint32_t&  sGBase = S.gBase;
  uint32_t&  sBig   = S.big;
  uint8_t*   active = S.active;
  const int      t = threadIdx.x, lane = t & 31, w = t >> 5;
  const uint32_t rootId = smallRoots[blockIdx.x];
  const uint32_t base = nodes[rootId].left, total = nodes[rootId].right - base;
  int            o[3] = {ox, oy, oz};
  for ( uint32_t i = t; i < total; i += KB_TPB ) {
    rec[i] = grec[base + i];
    nid[i] = 0;
  }
  if ( t < KB_CAP / 32 ) { active[t] = (uint32_t)t * 32 < total; }
  if ( t == 0 ) {
    const KdNode r = nodes[rootId];
    KbNext       n{};
    n.gid = rootId, n.pgid = 0, n.left = 0, n.right = (uint16_t)total, n.pfeat = 0, n.side = 0;
    for ( int k = 0; k < 3; k++ ) { n.lo[k] = r.lo[k], n.hi[k] = r.hi[k]; }
    nxt[0] = n;
    sNNext = 1;
  }
  __syncthreads();
  int level = 0;
  for ( ;; level++ ) {
    // ---- this level's nodes ----
    const uint32_t nl = sNNext;
    __syncthreads();
    if ( nl == 0 ) { break; }
    for ( uint32_t j = t; j < nl; j += KB_TPB ) {
      const KbNext x = nxt[j];
      KbNode       n{};
      n.gid = x.gid, n.pgid = x.pgid, n.left = x.left, n.right = x.right, n.pfeat = x.pfeat, n.side = x.side;
      for ( int k = 0; k < 3; k++ ) {
        n.lo[k] = x.lo[k], n.hi[k] = x.hi[k];
        n.tmin[k] = 0x7FFFFFFF, n.tmax[k] = (int32_t)0x80000000;
      }
      cur[j] = n;
    }
    if ( t == 0 ) { sNNext = 0, sBig = 0; }
    __syncthreads();
    // ---- tight boxes ----
#pragma unroll
    for ( int q = 0; q < KB_EPT; q++ ) {
      const uint32_t e = q * KB_TPB + t;
      if ( !active[e >> 5] ) { continue; }  // warp-uniform
      const uint32_t j = e < total ? nid[e] : 0xFFFFu;
      const bool     on = j != 0xFFFFu;
      int            c[3] = {0, 0, 0};
      if ( on ) {
        const uint64_t r = rec[e];
        c[0] = kd_coord( r, 0 ), c[1] = kd_coord( r, 1 ), c[2] = kd_coord( r, 2 );
      }
      const uint32_t act = __ballot_sync( 0xFFFFFFFFu, on );
      int allSame = 0;
      if ( act == 0xFFFFFFFFu ) { __match_all_sync( 0xFFFFFFFFu, j, &allSame ); }
      if ( allSame ) {
        int mn[3], mx[3];
#pragma unroll
        for ( int k = 0; k < 3; k++ ) {
          mn[k] = __reduce_min_sync( 0xFFFFFFFFu, c[k] );
          mx[k] = __reduce_max_sync( 0xFFFFFFFFu, c[k] );
        }
        if ( lane == 0 ) {
          for ( int k = 0; k < 3; k++ ) {
            atomicMin( &cur[j].tmin[k], mn[k] );
            atomicMax( &cur[j].tmax[k], mx[k] );
          }
        }
      } else if ( on ) {
        for ( int k = 0; k < 3; k++ ) {
          atomicMin( &cur[j].tmin[k], c[k] );
          atomicMax( &cur[j].tmax[k], c[k] );
        }
      }
    }
    __syncthreads();
    // ---- leaf / split decision; the node reports its tight bound to its parent (divideTree :1080-1081) ----
    for ( uint32_t j = t; j < nl; j += KB_TPB ) {
      KbNode&        n     = cur[j];
      const uint32_t count = n.right - n.left;
      if ( n.pgid ) {
        if ( n.side == 0 ) {
          nodes[n.pgid].divlow = (int16_t)n.tmax[n.pfeat];
        } else {
          nodes[n.pgid].divhigh = (int16_t)n.tmin[n.pfeat];
        }
      }
      if ( count <= 10 ) {  // leaf_max_size, PCCKdTree.cpp:58
        KdNode g{};
        g.left = base + n.left, g.right = base + n.right, g.child1 = 0, g.state = 3;
        for ( int k = 0; k < 3; k++ ) {
          g.lo[k] = n.lo[k], g.hi[k] = n.hi[k];
          g.tmin[k] = (int16_t)n.tmin[k], g.tmax[k] = (int16_t)n.tmax[k];
        }
        nodes[n.gid] = g;
        n.state      = 3;
      } else {
        KdNode tmp{};
        for ( int k = 0; k < 3; k++ ) {
          tmp.lo[k] = n.lo[k], tmp.hi[k] = n.hi[k];
          tmp.tmin[k] = (int16_t)n.tmin[k], tmp.tmax[k] = (int16_t)n.tmax[k];
        }
        kd_choose_split( tmp, o );
        n.cutfeat = (uint8_t)tmp.cutfeat, n.cutval = tmp.cutval;
        n.lt = n.le = 0;
        n.state     = 1;
        atomicAdd( &sBig, 1u );
      }
    }
    __syncthreads();
    if ( sBig == 0 ) { break; }
    // ---- lim1 / lim2 (one shared-memory atomic per warp when the warp sits inside one node) ----
#pragma unroll
    for ( int q = 0; q < KB_EPT; q++ ) {
      const uint32_t e  = q * KB_TPB + t;
      if ( !active[e >> 5] ) { continue; }
      const uint32_t j  = e < total ? nid[e] : 0xFFFFu;
      const bool     on = j != 0xFFFFu && cur[j].state == 1;
      bool           lt = false, le = false;
      if ( on ) {
        const int v = kd_coord( rec[e], cur[j].cutfeat ), cut = cur[j].cutval;
        lt = v < cut, le = v <= cut;
      }
      int allSame = 0;
      __match_all_sync( 0xFFFFFFFFu, on ? j : 0xFFFFu, &allSame );
      if ( allSame ) {
        const uint32_t bl = __ballot_sync( 0xFFFFFFFFu, lt ), be = __ballot_sync( 0xFFFFFFFFu, le );
        if ( on && lane == 0 ) {
          if ( bl ) { atomicAdd( &cur[j].lt, (uint32_t)__popc( bl ) ); }
          if ( be ) { atomicAdd( &cur[j].le, (uint32_t)__popc( be ) ); }
        }
      } else if ( on ) {
        if ( lt ) { atomicAdd( &cur[j].lt, 1u ); }
        if ( le ) { atomicAdd( &cur[j].le, 1u ); }
      }
    }
    __syncthreads();
    // ---- the two Hoare passes of planeSplit (:1154-1181) ----
    for ( int pass = 0; pass < 2; pass++ ) {
      // flags + block exclusive scan.  Warp w owns the 256 consecutive elements [256 w, 256 w + 256) and walks them 32
      // at a time (consecutive lanes = consecutive elements: conflict-free), carrying its running count in a register
      uint32_t fbits = 0, run = 0;
      uint16_t pre[KB_EPT];
#pragma unroll
      for ( int q = 0; q < KB_EPT; q++ ) {
        const uint32_t e = w * ( KB_EPT * 32 ) + q * 32 + lane;
        uint32_t       m = 0;
        const uint32_t j = ( active[e >> 5] && e < total ) ? nid[e] : 0xFFFFu;
        if ( j != 0xFFFFu && cur[j].state == 1 ) {
          const KbNode&  n = cur[j];
          const uint32_t p = e - n.left;
          const int      v = kd_coord( rec[e], n.cutfeat );
          if ( pass == 0 ) {
            const bool in = v < n.cutval;
            m             = ( p < n.lt ) ? !in : in;
          } else if ( p >= n.lt ) {
            const bool in = v <= n.cutval;
            m             = ( p < n.le ) ? !in : in;
          }
        }
        const uint32_t bal = __ballot_sync( 0xFFFFFFFFu, m != 0 );
        pre[q]             = (uint16_t)( run + __popc( bal & ( ( 1u << lane ) - 1u ) ) );
        run += __popc( bal );
        fbits |= m << q;
      }
      if ( lane == 0 ) { warpSum[w] = run; }
      __syncthreads();
      uint32_t wbase = 0;
      for ( int k = 0; k < w; k++ ) { wbase += warpSum[k]; }
#pragma unroll
      for ( int q = 0; q < KB_EPT; q++ ) { scan[w * ( KB_EPT * 32 ) + q * 32 + lane] = (uint16_t)( wbase + pre[q] ); }
      if ( t == KB_TPB - 1 ) { scan[KB_CAP] = (uint16_t)( wbase + run ); }
      __syncthreads();
      // pair lists
#pragma unroll
      for ( int q = 0; q < KB_EPT; q++ ) {
        const uint32_t e = w * ( KB_EPT * 32 ) + q * 32 + lane;
        if ( ( fbits >> q ) & 1u ) {
          const KbNode&  n     = cur[nid[e]];
          const uint32_t begin = pass == 0 ? n.left : n.left + n.lt;
          const uint32_t lim   = pass == 0 ? n.left + n.lt : n.left + n.le;
          const uint32_t m     = ( (uint32_t)scan[n.right] - scan[begin] ) >> 1;
          const uint32_t r     = (uint32_t)scan[e] - scan[begin];
          if ( e < lim ) {
            pairL[begin + r] = (uint16_t)e;
          } else {
            pairR[begin + ( m - 1 - ( r - m ) )] = (uint16_t)e;
          }
        }
      }
      __syncthreads();
      // swaps
#pragma unroll
      for ( int q = 0; q < KB_EPT; q++ ) {
        const uint32_t e = q * KB_TPB + t;
        if ( !active[e >> 5] ) { continue; }
        const uint32_t j = e < total ? nid[e] : 0xFFFFu;
        if ( j != 0xFFFFu && cur[j].state == 1 ) {
          const KbNode&  n     = cur[j];
          const uint32_t begin = pass == 0 ? n.left : n.left + n.lt;
          if ( e >= begin ) {
            const uint32_t m = ( (uint32_t)scan[n.right] - scan[begin] ) >> 1, k = e - begin;
            if ( k < m ) {
              const uint32_t a = pairL[begin + k], b = pairR[begin + k];
              const uint64_t ra = rec[a], rb = rec[b];
              rec[a] = rb;
              rec[b] = ra;
            }
          }
        }
      }
      __syncthreads();
    }
    // ---- children (divideTree :1070-1078) ----
    for ( uint32_t j = t; j < nl; j += KB_TPB ) {
      KbNode& n = cur[j];
      if ( n.state == 1 ) { n.child = (uint16_t)atomicAdd( &sNNext, 2u ); }
    }
    __syncthreads();
    if ( t == 0 ) { sGBase = atomicAdd( &counters[0], sNNext ); }
    __syncthreads();
    const uint32_t gbase = sGBase;
    if ( gbase + sNNext > nodeCap ) {  // node pool exhausted: flag it, leave the subtree as a (wrong) leaf — the host fails
      if ( t == 0 ) { counters[3] = 1; }
      break;
    }
    for ( uint32_t j = t; j < nl; j += KB_TPB ) {
      KbNode& n = cur[j];
      if ( n.state != 1 ) { continue; }
      const uint32_t count = n.right - n.left;
      uint32_t       idx;  // :1137-1139
      if ( n.lt > count / 2 ) {
        idx = n.lt;
      } else if ( n.le < count / 2 ) {
        idx = n.le;
      } else {
        idx = count / 2;
      }
      KbNext a{}, b{};
      a.gid = gbase + n.child, b.gid = gbase + n.child + 1;
      a.pgid = b.pgid = n.gid;
      a.pfeat = b.pfeat = n.cutfeat;
      a.side = 0, b.side = 1;
      a.left = n.left, a.right = (uint16_t)( n.left + idx ), b.left = (uint16_t)( n.left + idx ), b.right = n.right;
      for ( int k = 0; k < 3; k++ ) {
        a.lo[k] = b.lo[k] = n.lo[k];
        a.hi[k] = b.hi[k] = n.hi[k];
      }
      a.hi[n.cutfeat] = n.cutval;  // left_bbox[cutfeat].high = cutval
      b.lo[n.cutfeat] = n.cutval;  // right_bbox[cutfeat].low = cutval
      nxt[n.child]     = a;
      nxt[n.child + 1] = b;
      KdNode g{};
      g.left = base + n.left, g.right = base + n.right, g.child1 = a.gid, g.lt = n.lt, g.le = n.le;
      g.cutfeat = (int8_t)n.cutfeat, g.cutval = n.cutval, g.state = 1;
      for ( int k = 0; k < 3; k++ ) {
        g.lo[k] = n.lo[k], g.hi[k] = n.hi[k];
        g.tmin[k] = (int16_t)n.tmin[k], g.tmax[k] = (int16_t)n.tmax[k];
      }
      nodes[n.gid] = g;  // divlow / divhigh are written by the children at the next level
    }
    __syncthreads();
    // ---- elements move to their child; elements of finished leaves drop out ----
#pragma unroll
    for ( int q = 0; q < KB_EPT; q++ ) {
      const uint32_t e = q * KB_TPB + t;
      if ( !active[e >> 5] ) { continue; }
      bool still = false;
      if ( e < total ) {
        const uint32_t j = nid[e];
        if ( j != 0xFFFFu ) {
          const KbNode& n = cur[j];
          still           = n.state == 1;
          nid[e]          = still ? (uint16_t)( e < nxt[n.child].right ? n.child : n.child + 1 ) : (uint16_t)0xFFFFu;
        }
      }
      const uint32_t any = __ballot_sync( 0xFFFFFFFFu, still );
      if ( lane == 0 && any == 0 ) { active[e >> 5] = 0; }
    }
    __syncthreads();
  }
  __syncthreads();
  for ( uint32_t i = t; i < total; i += KB_TPB ) { grec[base + i] = rec[i]; }
  if ( t == 0 ) { atomicMax( &counters[4], (uint32_t)level + 1u ); }
}

// divlow / divhigh from the children's tight boxes (divideTree :1080-1081); depth bookkeeping
__global__ void k_kd_finalize( KdNode* __restrict__ nodes, uint32_t nNodes ) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if ( i >= nNodes || i == 0 ) { return; }
  KdNode& n = nodes[i];
  if ( n.child1 == 0 ) { return; }
  n.divlow  = nodes[n.child1].tmax[n.cutfeat];
  n.divhigh = nodes[n.child1 + 1].tmin[n.cutfeat];
}

}  // namespace

int rb_kd_build( rb200_ctx* c, RbKdBuild& B, const short4* pos, const int64_t* dOff, const std::vector<int64_t>& hOff, int ox,
                 int oy, int oz ) {
  const int     nTrees = (int)hOff.size() - 1;
  const int64_t E      = hOff[nTrees];
  if ( E <= 0 || nTrees <= 0 ) { return rb_fail( c, RB200_ERR_INVALID, "kd build: empty forest" ); }
  if ( E >= ( 1ll << 31 ) ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "kd build: more than 2^31 points" ); }
  for ( int t = 0; t < nTrees; t++ ) {
    if ( hOff[t + 1] - hOff[t] <= 0 ) { return rb_fail( c, RB200_ERR_INVALID, "kd build: empty cloud %d", t ); }
    if ( hOff[t + 1] - hOff[t] >= ( 1ll << 28 ) ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "kd build: cloud too large" ); }
  }
  const uint32_t nodeCap = (uint32_t)std::min<int64_t>( 2 * E + nTrees + 16, ( E * 3 ) / 4 + 64ll * nTrees + 4096 );
  RB_CUDA( B.rec.ensure( (size_t)E * 8 ) );
  RB_CUDA( B.nid.ensure( (size_t)E * 4 ) );
  RB_CUDA( B.nodes.ensure( (size_t)nodeCap * sizeof( KdNode ) ) );
  RB_CUDA( B.flags.ensure( (size_t)( E + 8 ) * 4 ) );
  RB_CUDA( B.pairL.ensure( (size_t)E * 4 ) );
  RB_CUDA( B.pairR.ensure( (size_t)E * 4 ) );
  RB_CUDA( B.sums.ensure( rb_scan_scratch_bytes( E + 1 ) ) );
  RB_CUDA( B.smallRoots.ensure( (size_t)( E + nTrees ) * 4 ) );
  RB_CUDA( B.counters.ensure( 64 ) );
  const uint32_t statCap = (uint32_t)std::min<int64_t>( nodeCap, E / 16 + 64ll * nTrees + 4096 );
  RB_CUDA( B.stats.ensure( (size_t)statCap * 24 ) );
  int32_t* st = B.stats.as<int32_t>();
  uint64_t* rec      = B.rec.as<uint64_t>();
  uint32_t* nid      = B.nid.as<uint32_t>();
  KdNode*   nodes    = B.nodes.as<KdNode>();
  uint32_t* flags    = B.flags.as<uint32_t>();
  uint32_t* counters = B.counters.as<uint32_t>();
  uint32_t* h        = (uint32_t*)rb_pinned( c, 64 );
  if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  // counters: [0] next free node, [1] small roots, [2] big nodes of the level, [3] pool exhausted, [4] depth, [5] range error
  h[0] = (uint32_t)nTrees + 1;
  h[1] = h[2] = h[3] = h[4] = h[5] = 0;
  RB_CUDA( cudaMemcpyAsync( counters, h, 32, cudaMemcpyHostToDevice, c->stream ) );
  const int G = rb_div_up( E, TPB );
  RB_LAUNCH( "kd_init", k_kd_init, G, TPB, 0, pos, dOff, nTrees, E, ox, oy, oz, rec, nid, counters + 5 );
  RB_LAUNCH( "kd_roots", k_kd_roots, rb_div_up( nTrees, 128 ), 128, 0, nodes, dOff, nTrees, st );
  uint32_t lvlBegin = 1, lvlEnd = (uint32_t)nTrees + 1;
  int      level = 0;
  bool     roots = true;
  for ( ;; level++ ) {
    if ( level > 200 ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "kd build: tree deeper than 200 levels" ); }
    const uint32_t nLvl = lvlEnd - lvlBegin;
    if ( level == 0 ) { RB_LAUNCH( "kd_stats", k_kd_stats<false>, G, TPB, 0, rec, nid, nodes, st, E, lvlBegin, lvlEnd ); }
    RB_CUDA( cudaMemsetAsync( counters + 2, 0, 4, c->stream ) );
    RB_LAUNCH( "kd_split", k_kd_split, rb_div_up( nLvl, 128 ), 128, 0, nodes, lvlBegin, lvlEnd, KB_CAP, ox, oy, oz,
               B.smallRoots.as<uint32_t>(), counters, roots ? 1 : 0, st );
    roots = false;
    RB_CUDA( cudaMemcpyAsync( h, counters, 32, cudaMemcpyDeviceToHost, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
    if ( h[5] ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "kd build: coordinate range of the clouds exceeds 4096" ); }
    if ( h[2] == 0 ) { break; }  // no node of this level is large enough for the level-parallel phase
    RB_LAUNCH( "kd_count", k_kd_count, G, TPB, 0, rec, nid, nodes, E, lvlBegin, lvlEnd );
    for ( int pass = 0; pass < 2; pass++ ) {
      RB_LAUNCH( "kd_flag", k_kd_flag, rb_div_up( E + 1, TPB ), TPB, 0, rec, nid, nodes, E, lvlBegin, lvlEnd, pass, flags );
      int r = rb_scan_u32( c, flags, flags, E + 1, B.sums.as<uint32_t>() );
      if ( r ) { return r; }
      RB_LAUNCH( "kd_pairs", k_kd_pairs, G, TPB, 0, nid, nodes, E, lvlBegin, lvlEnd, pass, flags, nullptr, B.pairL.as<uint32_t>(),
                 B.pairR.as<uint32_t>() );
      RB_LAUNCH( "kd_swap", k_kd_swap, G, TPB, 0, rec, nid, nodes, E, lvlBegin, lvlEnd, pass, flags, B.pairL.as<uint32_t>(),
                 B.pairR.as<uint32_t>() );
    }
    const uint32_t before = h[0];
    RB_LAUNCH( "kd_children", k_kd_children, rb_div_up( nLvl, 128 ), 128, 0, nodes, lvlBegin, lvlEnd, counters,
               std::min( nodeCap, statCap ), st );
    // assignment to the children + the children's tight boxes (the next level's statistics) in one pass
    RB_LAUNCH( "kd_assign", k_kd_stats<true>, G, TPB, 0, rec, nid, nodes, st, E, lvlBegin, lvlEnd );
    lvlBegin = before;
    lvlEnd   = before + 2 * h[2];
    if ( lvlEnd > std::min( nodeCap, statCap ) ) { return rb_fail( c, RB200_ERR_NOMEM, "kd build: node pool exhausted" ); }
  }
  const uint32_t nRoots = h[1], nLevelNodes = h[0];  // nodes [1, nLevelNodes) were created by the level-parallel phase
  if ( nRoots ) {
    const size_t smem = sizeof( KbShared );
    static bool  attr = false;
    if ( !attr ) {
      RB_CUDA( cudaFuncSetAttribute( k_kd_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem ) );
      attr = true;
    }
    RB_LAUNCH( "kd_block", k_kd_block, nRoots, KB_TPB, smem, rec, nodes, B.smallRoots.as<uint32_t>(), nRoots, counters, nodeCap,
               ox, oy, oz );
  }
  RB_CUDA( cudaMemcpyAsync( h, counters, 32, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  if ( h[3] ) { return rb_fail( c, RB200_ERR_NOMEM, "kd build: node pool exhausted" ); }
  const uint32_t nNodes = h[0];
  RB_LAUNCH( "kd_finalize", k_kd_finalize, rb_div_up( nLevelNodes, TPB ), TPB, 0, nodes, nLevelNodes );
  if ( level + (int)h[4] + 2 >= KD_STACK ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "kd build: tree depth %d exceeds the traversal stack", level + (int)h[4] );
  }
  B.forest.rec   = rec;
  B.forest.nodes = nodes;
  B.forest.ox    = ox;
  B.forest.oy    = oy;
  B.forest.oz    = oz;
  B.nNodes       = nNodes;
  B.nTrees       = nTrees;
  B.levels       = level + (int)h[4];
  return RB200_OK;
}
