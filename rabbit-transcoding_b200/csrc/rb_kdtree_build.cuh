// rb_kdtree_build.cuh — host-side handle of a kd forest build (see rb_kdtree.cu)
#pragma once
#include <vector>

#include "rb_common.cuh"
#include "rb_kdtree.cuh"


struct RbKdBuild {
  RbBuf    rec, tmp, gnodes, nodes, rootBox, chunkNode, chunks, cls, wA, wB, wC, cA, cB, smallRoots, largeList, counters, stats;
  KdForest forest{};
  uint32_t nNodes = 0;
  int      nTrees = 0;
  int      levels = 0;
  void     release() {
    RbBuf* b[] = {&rec, &tmp, &gnodes, &nodes, &rootBox, &chunkNode, &chunks, &cls, &wA, &wB, &wC, &cA, &cB, &smallRoots, &largeList, &counters, &stats};
    for ( auto* x : b ) { x->release(); }
  }
};

// Builds one tree per cloud.  pos: concatenated points (device), dOff: [nTrees + 1] offsets on the device, hOff: the
// same on the host, (ox, oy, oz): origin subtracted from the coordinates (all relative coordinates must be < 4096).
int rb_kd_build( rb200_ctx* c, RbKdBuild& B, const short4* pos, const int64_t* dOff, const std::vector<int64_t>& hOff, int ox,
                 int oy, int oz );
