// rb_kdtree.cuh — bit-exact emulation of the reference's nearest-neighbour ORDER (sm_100a).
//
// PCCKdTree (PccLibCommon/source/PCCKdTree.cpp:42-79) is nanoflann's KDTreeSingleIndexAdaptor<L2_Simple_Adaptor<int16_t,
// ..., double>, ..., 3, size_t> with leaf_max_size 10 (dependencies/nanoflann/nanoflann.hpp, NANOFLANN_VERSION 0x123,
// vendored in the reference tree).  knnSearch returns equal-distance neighbours in *tree traversal order*
// (nanoflann.hpp:110-133, 1207-1253), and on an integer lattice ties at the k-th slot are the norm, so
// PCCPointSet3::transferColors16bitBP (k = 8 forward, k = 1 backward) is only reproducible bit-exactly if the tree
// itself — every split plane, the Hoare-style permutation of planeSplit, the leaf order — is reproduced.
// This header holds the forest layout and the per-thread searchLevel emulation; rb_kdtree.cu builds the forest.
#pragma once
#include <stdint.h>

// search node, 16 bytes = one vector load per visit.  The build-time state of a node (boxes, counts) lives in the
// builder's own scratch (rb_kdtree.cu); what nanoflann's searchLevel reads is all that is kept:
//   leaf     : a = first element in KdForest::rec, b = KD_LEAF | count            (node.lr.left / right)
//   interior : a = child1 (children are a and a + 1), b = cut axis, divlow / divhigh (node.sub)
struct KdNode {
  uint32_t a, b;
  int16_t  divlow, divhigh;
  uint32_t pad;
};
constexpr uint32_t KD_LEAF = 0x80000000u;

// element record: x | y << 12 | z << 24 | index << 36 (coordinates relative to the forest origin, < 4096;
// index = position of the point inside its cloud, < 2^28)
__host__ __device__ __forceinline__ int      kd_coord( uint64_t r, int axis ) { return (int)( ( r >> ( 12 * axis ) ) & 0xFFFu ); }
__host__ __device__ __forceinline__ uint32_t kd_index( uint64_t r ) { return (uint32_t)( r >> 36 ); }

struct KdForest {
  const uint64_t* rec;      // [E] permuted element records of all trees
  const KdNode*   nodes;    // node 0 is unused; tree t has root `t + 1`
  const int16_t*  rootBox;  // [nTrees + 1][6] tight box {min[3], max[3]} of every root (root_bbox, nanoflann.hpp:1009-1024)
  int             ox, oy, oz;  // origin subtracted from every coordinate
};

constexpr int      KD_STACK    = 96;
constexpr uint32_t KD_INF      = 0xFFFFFFFFu;
constexpr uint32_t KD_NODE_MAX = 1u << 29;  // node ids share a stack word with the axis and the entry kind

// KNNResultSet::addPoint (nanoflann.hpp:110-133): insertion from the back, only entries with a LARGER distance shift,
// so ties keep first-seen order and a candidate equal to the current k-th distance is dropped when the set is full.
// Unused slots hold KD_INF (no real distance reaches it), which makes the insertion a fixed sequence of K selects on
// registers: slot i takes the candidate when dist[i] > d >= dist[i - 1], the old dist[i - 1] when that is larger too.
template <int K>
struct KdResult {
  uint32_t dist[K];
  uint32_t idx[K];
  int      count;
  __device__ __forceinline__ void init() {
    count = 0;
#pragma unroll
    for ( int i = 0; i < K; i++ ) { dist[i] = KD_INF, idx[i] = 0; }  // worstDist() = (std::numeric_limits<DistanceType>::max)(), :96
  }
  __device__ __forceinline__ uint32_t worst() const { return dist[K - 1]; }
  __device__ __forceinline__ void     add( uint32_t d, uint32_t index ) {
#pragma unroll
    for ( int i = K - 1; i >= 1; --i ) {
      if ( dist[i] > d ) {
        const bool sh = dist[i - 1] > d;
        dist[i]       = sh ? dist[i - 1] : d;
        idx[i]        = sh ? idx[i - 1] : index;
      }
    }
    if ( dist[0] > d ) { dist[0] = d, idx[0] = index; }
    if ( count < K ) { count++; }
  }
};

// findNeighbors + searchLevel (nanoflann.hpp:901-915, 1207-1253) for the tree rooted at `root`; the query is given in
// forest-relative coordinates (may lie outside [0, 4096): it is an int).  All distances are exact integers.
// One loop, one step per iteration — a node visit (interior: push the far child and go to the near one; leaf: scan) or
// one entry off the stack — so the threads of a warp meet again after every step.  mindistsq is always the sum of
// dists[] (:911, :1246-1251), so an entry is two words: far child | axis << 29 | kind << 31, and cut_dist (kind 0,
// "the far child is pending") or the saved dists[axis] (kind 1, "restore it", :1251).
template <int K>
__device__ __forceinline__ void kd_search( const KdForest& f, uint32_t root, const int q[3], KdResult<K>& res ) {
  res.init();
  uint32_t d0 = 0, d1 = 0, d2 = 0;
  {
    const int16_t* rb = f.rootBox + (size_t)root * 6;  // computeInitialDistances against root_bbox (:1183-1201)
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
  uint32_t node = root;  // 0 (never a node) = nothing to visit, take the next entry off the stack
  for ( ;; ) {
    if ( node ) {
      const uint4 nv = __ldg( reinterpret_cast<const uint4*>( f.nodes + node ) );
      if ( nv.y & KD_LEAF ) {
        const uint32_t worst = res.worst();  // read once per leaf (:1213)
        const uint32_t lend  = nv.x + ( nv.y & 0xFFFFu );
        for ( uint32_t i = nv.x; i < lend; i++ ) {
          const uint64_t r  = f.rec[i];
          const int      dx = q[0] - kd_coord( r, 0 ), dy = q[1] - kd_coord( r, 1 ), dz = q[2] - kd_coord( r, 2 );
          const uint32_t d  = (uint32_t)( dx * dx ) + (uint32_t)( dy * dy ) + (uint32_t)( dz * dz );
          if ( d < worst ) { res.add( d, kd_index( r ) ); }
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
        const uint32_t m   = d0 + d1 + d2 + v - dst;  // mindistsq + cut_dist - dists[idx] (:1246)
        if ( m <= res.worst() ) {                     // mindistsq * epsError <= worstDist(), epsError = 1 (:1248)
          stN[sp] = e | 0x80000000u;                  // restore dists[axis] = dst once the far subtree is done (:1251)
          stV[sp] = dst;
          sp++;
          d0 = axis == 0 ? v : d0, d1 = axis == 1 ? v : d1, d2 = axis == 2 ? v : d2;
          node = e & ( KD_NODE_MAX - 1u );
        }
      }
    }
  }
}
