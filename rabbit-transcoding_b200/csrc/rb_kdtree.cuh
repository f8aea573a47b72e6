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

constexpr int      KD_STACK = 96;
constexpr uint32_t KD_INF   = 0xFFFFFFFFu;

// KNNResultSet::addPoint (nanoflann.hpp:110-133): insertion from the back, only entries with a LARGER distance shift,
// so ties keep first-seen order and a candidate equal to the current k-th distance is dropped when the set is full.
template <int K>
struct KdResult {
  uint32_t dist[K];
  uint32_t idx[K];
  int      count;
  __device__ __forceinline__ void init() {
    count       = 0;
    dist[K - 1] = KD_INF;  // (std::numeric_limits<DistanceType>::max)(), :96
  }
  __device__ __forceinline__ uint32_t worst() const { return dist[K - 1]; }
  __device__ __forceinline__ void     add( uint32_t d, uint32_t index ) {
    int i;
    for ( i = count; i > 0; --i ) {
      if ( dist[i - 1] > d ) {
        if ( i < K ) {
          dist[i] = dist[i - 1];
          idx[i]  = idx[i - 1];
        }
      } else {
        break;
      }
    }
    if ( i < K ) {
      dist[i] = d;
      idx[i]  = index;
    }
    if ( count < K ) { count++; }
  }
};

// findNeighbors + searchLevel (nanoflann.hpp:901-915, 1207-1253) for the tree rooted at `root`; the query is given in
// forest-relative coordinates (may lie outside [0, 4096): it is an int).  All distances are exact integers.
template <int K>
__device__ __forceinline__ void kd_search( const KdForest& f, uint32_t root, const int q[3], KdResult<K>& res ) {
  res.init();
  uint32_t dists[3] = {0, 0, 0};
  uint32_t mind     = 0;
  {
    const int16_t* rb = f.rootBox + (size_t)root * 6;  // computeInitialDistances against root_bbox (:1183-1201)
#pragma unroll
    for ( int i = 0; i < 3; i++ ) {
      const int tmin = rb[i], tmax = rb[3 + i];
      if ( q[i] < tmin ) {
        const int d = q[i] - tmin;
        dists[i]    = (uint32_t)( d * d );
        mind += dists[i];
      }
      if ( q[i] > tmax ) {
        const int d = q[i] - tmax;
        dists[i]    = (uint32_t)( d * d );
        mind += dists[i];
      }
    }
  }
  // explicit stack: kind 0 = "visit the other child" (pending check), kind 1 = "restore dists[axis]"
  uint32_t stNode[KD_STACK];
  uint32_t stA[KD_STACK];  // pending: cut_dist;        restore: saved dists[axis]
  uint32_t stB[KD_STACK];  // pending: mindistsq at the parent | axis << 30 ... kept separately below
  uint8_t  stAxis[KD_STACK];
  uint8_t  stKind[KD_STACK];
  int      sp   = 0;
  uint32_t node = root;
  uint32_t cur  = mind;
  for ( ;; ) {
    // ---- descend to a leaf ----
    for ( ;; ) {
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
        break;
      }
      const int axis   = (int)( nv.y & 3u );
      const int val    = axis == 0 ? q[0] : ( axis == 1 ? q[1] : q[2] );
      const int divlow = (int16_t)( nv.z & 0xFFFFu ), divhigh = (int16_t)( nv.z >> 16 );
      const int diff1  = val - divlow, diff2 = val - divhigh;
      uint32_t  best, other;
      int       cd;
      if ( diff1 + diff2 < 0 ) {
        best  = nv.x;
        other = nv.x + 1;
        cd    = val - divhigh;
      } else {
        best  = nv.x + 1;
        other = nv.x;
        cd    = val - divlow;
      }
      stNode[sp] = other;
      stA[sp]    = (uint32_t)( cd * cd );
      stB[sp]    = cur;
      stAxis[sp] = (uint8_t)axis;
      stKind[sp] = 0;
      sp++;
      node = best;
    }
    // ---- unwind ----
    bool descend = false;
    while ( sp > 0 ) {
      sp--;
      const int axis = stAxis[sp];
      if ( stKind[sp] == 1 ) {
        dists[axis] = stA[sp];
        continue;
      }
      const uint32_t cut = stA[sp], dst = dists[axis];
      const uint32_t m   = stB[sp] + cut - dst;  // mindistsq + cut_dist - dists[idx] (:1246)
      if ( m <= res.worst() ) {                  // mindistsq * epsError <= worstDist(), epsError = 1 (:1248)
        const uint32_t other = stNode[sp];
        stKind[sp]           = 1;  // restore dists[axis] = dst once the other subtree is done (:1251)
        stA[sp]              = dst;
        sp++;
        dists[axis] = cut;
        node        = other;
        cur         = m;
        descend     = true;
        break;
      }
    }
    if ( !descend ) { break; }
  }
}
