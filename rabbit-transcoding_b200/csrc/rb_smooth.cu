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
//  * the grid is sparse and two-level: a small per-frame hash table maps a 4x4x4 *block* of cells to a block of 64
//    accumulators taken from a pool (one 8-byte table entry per occupied block, a few thousand per frame), so every
//    cell lookup is "one table entry + one 32-byte accumulator", collisions are rare (table load < 25 %) and resolved
//    for all eight neighbours of a point together; the table and the touched accumulators of a frame stay in L2;
//  * every occupied cell is accumulated (the reference only accumulates cells marked by a boundary point, but its
//    filter never reads any other cell, so the result is the same and the marking pass disappears);
//  * accumulators are integer atomics.  The reference sums small integers in float in emission order; below 2^24
//    every partial sum is exact, so the order-free integer sum converts to the identical float;
//    a per-frame flag reports any cell that leaves that range;
//  * "doSmooth" (a cell holds two different partitions) is order-free: max and min partition, compared by the filter;
//  * the per-cell luma lists of the colour gate are filled by the accumulation itself (the atomic add on the count
//    hands out list positions), and the median is only computed for cells whose luma variance can exceed the gate
//    (|mean - median| <= standard deviation);
//  * the filter passes repeat the reference's double arithmetic operation by operation (compiled with
//    -fmad=false), including the integer-truncating abs() of the colour gates.
#include <algorithm>
#include <type_traits>

#include "rb_common.cuh"

namespace {

// accumulator of one cell; all-zero == empty.  32 bytes = one L2 sector; the filters read the first 16 bytes.
struct Cell {
  uint32_t s0, s1;    // {s0, s1} and {cw, s2} are each updated with ONE 64-bit atomic add
  uint32_t cw, s2;    // cw = point count | FC_GATE;  s* = coordinate sums (geometry) / colour sums (colour)
  uint32_t pmax;      // max of (partition + 1) over the points of the cell
  uint32_t pmin_inv;  // max of ~(partition + 1): the cell holds two partitions (the reference's doSmooth, :989-995)
                      // iff pmax != ~pmin_inv.  Order-free, and no atomic has to return a value
  unsigned long long q2;  // colour: sum of luma^2 (bounds |mean - median| by the standard deviation)
};
// cw: the count can never carry into the flag (a frame has < 2^28 points); FC_GATE is the mean/median gate of
// gridFilteringColor (:1228-1243)
constexpr uint32_t FC_GATE = 1u << 31, FC_MULTI = 1u << 30, FC_CNT = 0x0FFFFFFFu;
// Once a grid is complete a per-cell pass turns the accumulators, in place, into what the filters read — so the
// centre of a cell is divided once per cell instead of once per boundary point and neighbour:
//   geometry: {float centre x, y, cw', centre z}            centre = float sum / float count (:135-137)
//   colour  : {double mean c0, c1, c2, cw', -}              mean = (double)(float)sum / (double)count (:1225)
// with cw' = count | FC_MULTI | FC_GATE.
template <bool COLOUR>
struct CellView {
  typename std::conditional<COLOUR, double, float>::type m0, m1, m2;
  uint32_t cnt;
  bool     multi, gate;
};
constexpr int      MAX_PROBES = 256;
constexpr uint32_t NO_BLOCK   = 0xFFFFFFFFu;
enum { CTR_CURSOR = 0, CTR_FLAGS = 1, CTR_MAXCNT = 2, CTR_RANGE = 3, CTR_ORDERED = 4 };  // counters[]
constexpr uint32_t ORDERED_CAP = 4096;  // cells per stage whose float sums are re-done in emission order (see k_ordered_cells)
enum { OVF_BLOCKS = 1, OVF_LUM = 2, OVF_SPIN = 4 };     // CTR_FLAGS bits

struct GridArgs {
  int            F;
  int            g;          // cell size
  uint32_t       ginv;       // floor( 2^32 / g ) + 1: x / g == __umulhi( x, ginv ) for 0 <= x < 2^16, 2 <= g <= 64
  int            wmax;       // cells per axis
  int            by_bbox;    // geometry: th = g * ceil(maxCoord / g); colour: th = 2^bitdepth
  int            pcmax;      // 2^geometryBitDepth3D
  unsigned long long* table; // [F][tslots] 0 = empty, else (bx | by << 8 | bz << 16) + 1  |  (block id + 1) << 32
  uint32_t       tslots;     // per frame, power of two
  int            tshift;     // 32 - log2( tslots )
  Cell*          cells;      // [cap_blocks][64] pool; block-local cell index = cx&3 | (cy&3) << 2 | (cz&3) << 4
  uint32_t       cap_blocks;
  uint16_t*      lum;        // colour: [cap_blocks][64][lum_cap] luma lists
  uint32_t       lum_cap;
  uint2*         binfo;      // [cap_blocks] {frame, block key} of every pool block (written by the thread that created it)
  uint32_t*      ordered;    // [ORDERED_CAP] pool cells whose sums left the exact float range
  int32_t*       counters;
  uint32_t*      marks;      // [F][wmax][wmax][mwords] one bit per cell: some boundary point blends it (or null: all)
  int            mwords;
  const int64_t* frame_off;
  RbFrameInfo*   finfo;
  short4*        pos;
  ushort4*       col;
  const uint32_t* part;
  const uint32_t* blist;     // indices of the points classified as boundary (type 1) by the reconstruction
  const uint32_t* blist_n;
  uint32_t*      moved_bits; // geometry: one bit per point, set for the points the filter moves (type 3)
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

__device__ __forceinline__ int cell_of( const GridArgs& a, int x ) {  // x / g for a non-negative coordinate
  return a.g == 1 ? x : (int)__umulhi( (uint32_t)x, a.ginv );
}
__device__ __forceinline__ uint32_t block_key( int cx, int cy, int cz ) {
  return ( (uint32_t)( cx >> 2 ) | ( (uint32_t)( cy >> 2 ) << 8 ) | ( (uint32_t)( cz >> 2 ) << 16 ) ) + 1u;
}
__device__ __forceinline__ uint32_t cell_local( int cx, int cy, int cz ) {
  return (uint32_t)( ( cx & 3 ) | ( ( cy & 3 ) << 2 ) | ( ( cz & 3 ) << 4 ) );
}
__device__ __forceinline__ uint32_t block_home( const GridArgs& a, uint32_t key ) { return ( key * 0x9E3779B1u ) >> a.tshift; }

// the 2x2x2 cells a boundary point blends.  All eight table probes are issued before any is examined, unresolved
// probes (another block's key in the slot) advance together, then the eight 16-byte records are fetched together:
// two dependent memory round trips in the common case.
template <bool COLOUR>
__device__ __forceinline__ void cell_find8( const GridArgs& a, int f, const int S[3], CellView<COLOUR> fc[8] ) {
  const uint32_t            mask = a.tslots - 1;
  const unsigned long long* T    = a.table + (size_t)f * a.tslots;
  uint32_t                  h[8], key[8], id[8];
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {  // k = dz*4 + dy*2 + dx, the reference's loop order (:1019-1027)
    const int cx = S[0] + ( k & 1 ), cy = S[1] + ( ( k >> 1 ) & 1 ), cz = S[2] + ( k >> 2 );
    key[k]       = block_key( cx, cy, cz );
    h[k]         = block_home( a, key[k] );
    id[k]        = NO_BLOCK;
  }
  uint32_t open = 0xFFu;
  for ( int probe = 0; probe < MAX_PROBES && open; probe++ ) {
    unsigned long long e[8];
#pragma unroll
    for ( int k = 0; k < 8; k++ ) {
      if ( open >> k & 1u ) { e[k] = __ldg( T + h[k] ); }
    }
#pragma unroll
    for ( int k = 0; k < 8; k++ ) {
      if ( !( open >> k & 1u ) ) { continue; }
      const uint32_t v = (uint32_t)e[k];
      if ( v == key[k] ) {
        id[k] = (uint32_t)( e[k] >> 32 ) - 1u;
        open &= ~( 1u << k );
      } else if ( v == 0 ) {
        open &= ~( 1u << k );
      } else {
        h[k] = ( h[k] + 1 ) & mask;
      }
    }
  }
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {
    const int cx = S[0] + ( k & 1 ), cy = S[1] + ( ( k >> 1 ) & 1 ), cz = S[2] + ( k >> 2 );
    uint32_t cw = 0;  // the cell holds no point
    if ( id[k] < a.cap_blocks ) {
      const Cell* c = a.cells + (size_t)id[k] * 64 + cell_local( cx, cy, cz );
      if ( COLOUR ) {
        const double2 m01 = __ldg( reinterpret_cast<const double2*>( c ) );
        const double2 m2w = __ldg( reinterpret_cast<const double2*>( c ) + 1 );
        fc[k].m0 = m01.x, fc[k].m1 = m01.y, fc[k].m2 = m2w.x;
        cw = (uint32_t)__double_as_longlong( m2w.y );
      } else {
        const uint4 lo = __ldg( reinterpret_cast<const uint4*>( c ) );
        fc[k].m0 = __uint_as_float( lo.x ), fc[k].m1 = __uint_as_float( lo.y ), fc[k].m2 = __uint_as_float( lo.w );
        cw = lo.z;
      }
    }
    fc[k].cnt   = cw & FC_CNT;
    fc[k].gate  = ( cw & FC_GATE ) != 0;
    fc[k].multi = fc[k].cnt != 0 && ( cw & FC_MULTI );
  }
}

// find or create the block of a cell; NO_BLOCK on overflow (flagged; the stage is then repeated with larger tables)
__device__ __forceinline__ uint32_t block_claim( const GridArgs& a, int f, uint32_t key ) {
  const uint32_t      mask = a.tslots - 1;
  unsigned long long* T    = a.table + (size_t)f * a.tslots;
  uint32_t            h    = block_home( a, key );
  for ( int probe = 0; probe < MAX_PROBES; probe++ ) {
    unsigned long long e = T[h];  // may be a stale L1 line: only a complete entry with the right key is trusted
    if ( (uint32_t)e != key || ( e >> 32 ) == 0 ) { e = *reinterpret_cast<volatile unsigned long long*>( T + h ); }
    if ( e == 0 ) {
      e = atomicCAS( T + h, 0ull, (unsigned long long)key );
      if ( e == 0 ) {  // this thread created the block: take accumulators from the pool and publish their id
        uint32_t id = (uint32_t)atomicAdd( &a.counters[CTR_CURSOR], 1 );
        if ( id >= a.cap_blocks ) {
          atomicOr( &a.counters[CTR_FLAGS], OVF_BLOCKS );
          id = NO_BLOCK - 1u;
        }
        if ( id < a.cap_blocks ) { a.binfo[id] = make_uint2( (uint32_t)f, key ); }
        atomicExch( T + h, (unsigned long long)key | ( (unsigned long long)( id + 1u ) << 32 ) );
        return id < a.cap_blocks ? id : NO_BLOCK;
      }
    }
    if ( (uint32_t)e == key ) {
      for ( int spin = 0; ( e >> 32 ) == 0; spin++ ) {  // created by another thread a moment ago: wait for the id
        if ( spin > ( 1 << 20 ) ) {
          atomicOr( &a.counters[CTR_FLAGS], OVF_SPIN );
          return NO_BLOCK;
        }
        e = *reinterpret_cast<volatile unsigned long long*>( T + h );
      }
      const uint32_t id = (uint32_t)( e >> 32 ) - 1u;
      return id < a.cap_blocks ? id : NO_BLOCK;
    }
    h = ( h + 1 ) & mask;
  }
  atomicOr( &a.counters[CTR_FLAGS], OVF_BLOCKS );
  return NO_BLOCK;
}

// ---- marking (:89-112, :170-195): the 2x2x2 cells every in-margin boundary point blends.  The reference numbers the
// marked cells; here they become one bit each in a dense bitmap, and the accumulation skips unmarked cells (the
// filters never read those), which is most of the cloud. ----
__global__ void __launch_bounds__( 256 ) k_mark_cells( const GridArgs a ) {
  const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
  if ( li >= *a.blist_n ) { return; }
  const int64_t i = a.blist[li];
  const short4  p = a.pos[i];
  if ( p.w != 1 ) { return; }
  const int f     = frame_of( a.frame_off, a.F, i );
  const int g     = a.g, hg = g / 2;
  const int disth = max( hg, 1 ), th = grid_th( a, f );
  if ( !inside( p.x, p.y, p.z, disth, th ) ) { return; }
  const int P[3] = {p.x, p.y, p.z};
  int       S[3];
  for ( int k = 0; k < 3; k++ ) { S[k] = P[k] / g + ( ( P[k] - ( P[k] / g ) * g < hg ) ? -1 : 0 ); }
  const uint64_t bits = 3ull << ( S[0] & 31 );  // cells S[0], S[0]+1: one word or two neighbouring words
  const uint32_t b0 = (uint32_t)bits, b1 = (uint32_t)( bits >> 32 );
#pragma unroll
  for ( int k = 0; k < 4; k++ ) {
    const int cy = S[1] + ( k & 1 ), cz = S[2] + ( k >> 1 );
    uint32_t* w  = a.marks + ( ( (size_t)f * a.wmax + cz ) * a.wmax + cy ) * a.mwords + ( S[0] >> 5 );
    if ( ( w[0] & b0 ) != b0 ) { atomicOr( w, b0 ); }  // a stale cached word only costs a redundant atomic
    if ( b1 && ( w[1] & b1 ) != b1 ) { atomicOr( w + 1, b1 ); }
  }
}

constexpr int ACC_RUN = 8;  // consecutive points per thread

// Points arrive in emission order (patch -> 16x16 block -> pixel row -> layer): the 8 consecutive points of a thread
// (4 neighbouring pixels x 2 layers) fall into one to four cells.  Every thread first merges its own points per cell
// in registers (no memory traffic), then the warp flushes "the next cell of every lane" together: two to four rounds
// of one table probe and three or four atomics per lane, instead of a flush at every point where some lane's cell
// changes.  (Combining equal cells across lanes with per-group __reduce_*_sync masks was measured 4x slower: partial
// masks are executed group by group.)
// ONEFRAME: all 256 points of the warp lie in frame `fw` (all but a handful of warps): no per-point frame bookkeeping
template <bool COLOUR, bool ONEFRAME>
__device__ __forceinline__ void accumulate_warp( const GridArgs& a, int64_t n, int64_t i0, int lane, int fw ) {
  short4        p[ACC_RUN];
  ushort4       cv[ACC_RUN];
  uint32_t      pp[ACC_RUN];
  if ( i0 + ACC_RUN <= n ) {  // 16-byte loads: the arena is 16-byte aligned and i0 is a multiple of 8
    const uint4* vp = reinterpret_cast<const uint4*>( a.pos + i0 );
    const uint4* vq = reinterpret_cast<const uint4*>( a.part + i0 );
#pragma unroll
    for ( int k = 0; k < ACC_RUN / 2; k++ ) {
      const uint4 v = vp[k];
      p[2 * k]      = make_short4( (short)( v.x & 0xFFFF ), (short)( v.x >> 16 ), (short)( v.y & 0xFFFF ), (short)( v.y >> 16 ) );
      p[2 * k + 1]  = make_short4( (short)( v.z & 0xFFFF ), (short)( v.z >> 16 ), (short)( v.w & 0xFFFF ), (short)( v.w >> 16 ) );
    }
#pragma unroll
    for ( int k = 0; k < ACC_RUN / 4; k++ ) {
      const uint4 v = vq[k];
      pp[4 * k] = v.x + 1u, pp[4 * k + 1] = v.y + 1u, pp[4 * k + 2] = v.z + 1u, pp[4 * k + 3] = v.w + 1u;
    }
    if ( COLOUR ) {
      const uint4* vc = reinterpret_cast<const uint4*>( a.col + i0 );
#pragma unroll
      for ( int k = 0; k < ACC_RUN / 2; k++ ) {
        const uint4 v = vc[k];
        cv[2 * k]     = make_ushort4( v.x & 0xFFFF, v.x >> 16, v.y & 0xFFFF, v.y >> 16 );
        cv[2 * k + 1] = make_ushort4( v.z & 0xFFFF, v.z >> 16, v.w & 0xFFFF, v.w >> 16 );
      }
    }
  } else {
#pragma unroll
    for ( int k = 0; k < ACC_RUN; k++ ) {
      const bool ok = i0 + k < n;
      p[k]          = ok ? a.pos[i0 + k] : make_short4( -1, -1, -1, 0 );
      pp[k]         = ok ? a.part[i0 + k] + 1u : 0u;
      if ( COLOUR ) { cv[k] = ok ? a.col[i0 + k] : make_ushort4( 0, 0, 0, 0 ); }
    }
  }
  // ---- per point: frame and cell key (cx | cy << 10 | cz << 20), `rem` = points that take part ----
  uint32_t ck[ACC_RUN];
  int      fr[ACC_RUN];
  uint32_t rem = 0;
  if ( i0 < n ) {
    int       f     = ONEFRAME ? fw : frame_of( a.frame_off, a.F, i0 );
    int64_t   fend  = ONEFRAME ? 0 : a.frame_off[f + 1];
    const int disth = max( a.g / 2, 1 );
    int       th    = COLOUR ? 0 : grid_th( a, f );
#pragma unroll
    for ( int k = 0; k < ACC_RUN; k++ ) {
      const int64_t i = i0 + k;
      ck[k]           = 0;
      fr[k]           = f;
      if ( i >= n ) { continue; }
      if ( !ONEFRAME ) {
        while ( i >= fend ) {  // the run crosses into the next frame (empty frames are skipped)
          f++;
          fend = a.frame_off[f + 1];
          if ( !COLOUR ) { th = grid_th( a, f ); }
        }
        fr[k] = f;
      }
      const short4 q = p[k];
      bool         in;
      if ( COLOUR ) {  // no margin test, :208-224 with the :212 guard
        in = q.x >= 0 && q.y >= 0 && q.z >= 0;
      } else {  // :120-134
        in = inside( q.x, q.y, q.z, disth, th );
      }
      if ( !in ) { continue; }
      const int cx = cell_of( a, q.x ), cy = cell_of( a, q.y ), cz = cell_of( a, q.z );
      if ( COLOUR && ( cx >= a.wmax || cy >= a.wmax || cz >= a.wmax ) ) { continue; }
      ck[k] = (uint32_t)cx | ( (uint32_t)cy << 10 ) | ( (uint32_t)cz << 20 );
      rem |= 1u << k;
    }
  } else {
#pragma unroll
    for ( int k = 0; k < ACC_RUN; k++ ) { ck[k] = 0, fr[k] = 0; }
  }
  if ( COLOUR && a.marks ) {
    // cells no boundary point blends are never read: the mark words of the thread's 8 points are fetched together
    // and unmarked points leave before the flush rounds (half of the rounds disappear)
    uint32_t w[ACC_RUN];
#pragma unroll
    for ( int k = 0; k < ACC_RUN; k++ ) {
      w[k] = 0;
      if ( rem >> k & 1u ) {
        const int cx = ck[k] & 1023, cy = ( ck[k] >> 10 ) & 1023, cz = ck[k] >> 20;
        w[k]         = a.marks[( ( (size_t)( ONEFRAME ? fw : fr[k] ) * a.wmax + cz ) * a.wmax + cy ) * a.mwords + ( cx >> 5 )];
      }
    }
#pragma unroll
    for ( int k = 0; k < ACC_RUN; k++ ) {
      if ( !( ( w[k] >> ( ck[k] & 31 ) ) & 1u ) ) { rem &= ~( 1u << k ); }
    }
  }
  // ---- one cell per lane and iteration, lanes with the same cell combined ----
  while ( __any_sync( 0xFFFFFFFFu, rem != 0 ) ) {
    bool           act   = rem != 0;
    const int      first = __ffs( rem ) - 1;
    uint32_t       key = 0;
    int            f   = fw;
#pragma unroll
    for ( int k = 0; k < ACC_RUN; k++ ) {
      if ( k == first ) {
        key = ck[k];
        if ( !ONEFRAME ) { f = fr[k]; }
      }
    }
    uint32_t           mem = 0, cnt = 0, t0 = 0, t1 = 0, t2 = 0, mx = 0, mn = 0xFFFFFFFFu;
    unsigned long long q2 = 0;
#pragma unroll
    for ( int k = 0; k < ACC_RUN; k++ ) {
      if ( ( rem >> k & 1u ) && ck[k] == key && ( ONEFRAME || fr[k] == f ) ) {
        mem |= 1u << k;
        cnt++;
        if ( COLOUR ) {
          t0 += cv[k].x, t1 += cv[k].y, t2 += cv[k].z;
          q2 += (unsigned long long)( (uint32_t)cv[k].x * (uint32_t)cv[k].x );
        } else {
          t0 += (uint32_t)p[k].x, t1 += (uint32_t)p[k].y, t2 += (uint32_t)p[k].z;
        }
        mx = max( mx, pp[k] );
        mn = min( mn, pp[k] );
      }
    }
    rem &= ~mem;
    const int cx = key & 1023, cy = ( key >> 10 ) & 1023, cz = key >> 20;
    const uint32_t mask = __ballot_sync( 0xFFFFFFFFu, act );
    if ( !act ) { continue; }
    // one lane per distinct block looks it up (or creates it): no lane ever waits for another lane of its own warp.
    // (all warp primitives here use the same mask in every lane: per-group masks are executed group by group)
    const uint32_t bk = block_key( cx, cy, cz );
    const uint32_t bpeers  = ONEFRAME ? __match_any_sync( mask, bk ) : __match_any_sync( mask, ( (unsigned long long)(uint32_t)f << 32 ) | bk );
    const int      bleader = __ffs( bpeers ) - 1;
    uint32_t       bl      = NO_BLOCK;
    if ( lane == bleader ) { bl = block_claim( a, f, bk ); }
    bl = __shfl_sync( mask, bl, bleader );
    if ( bl == NO_BLOCK ) { continue; }
    const uint32_t           cell = bl * 64u + cell_local( cx, cy, cz );
    Cell*                    c    = a.cells + cell;
    unsigned long long old = 0;
    if ( COLOUR ) {  // the returned count hands out the list positions; everything else is fire-and-forget
      old = atomicAdd( (unsigned long long*)&c->cw, (unsigned long long)cnt | ( (unsigned long long)t2 << 32 ) );
    } else {
      atomicAdd( (unsigned long long*)&c->cw, (unsigned long long)cnt | ( (unsigned long long)t2 << 32 ) );
    }
    atomicAdd( (unsigned long long*)&c->s0, (unsigned long long)t0 | ( (unsigned long long)t1 << 32 ) );
    atomicMax( &c->pmax, mx );
    atomicMax( &c->pmin_inv, ~mn );
    if ( COLOUR ) {  // the atomic add on the count hands out the list positions oldc .. oldc + cnt - 1
      atomicAdd( &c->q2, q2 );
      const uint32_t oldc = (uint32_t)old & FC_CNT;
      if ( oldc + cnt <= a.lum_cap ) {
        uint16_t* L = a.lum + (size_t)cell * a.lum_cap + oldc;
#pragma unroll
        for ( int k = 0; k < ACC_RUN; k++ ) {
          if ( mem >> k & 1u ) { *L++ = cv[k].x; }
        }
      }  // else: the gate kernel sees count > lum_cap and asks for a retry with longer lists
    }
  }
}

template <bool COLOUR>
__global__ void __launch_bounds__( 256, 4 ) k_accumulate( const GridArgs a, int64_t n ) {
  const int     lane = threadIdx.x & 31;
  const int64_t i0   = ( (int64_t)blockIdx.x * 256 + threadIdx.x ) * ACC_RUN;
  const int64_t w0   = i0 - (int64_t)lane * ACC_RUN;  // first point of the warp
  if ( w0 >= n ) { return; }
  const int  fw  = frame_of( a.frame_off, a.F, w0 );
  const bool one = min( w0 + 32 * ACC_RUN, n ) <= a.frame_off[fw + 1];
  if ( one ) {
    accumulate_warp<COLOUR, true>( a, n, i0, lane, fw );
  } else {
    accumulate_warp<COLOUR, false>( a, n, i0, lane, fw );
  }
}

// ---- warp-staged accumulation (geometry) ----
// Points arrive in emission order: the 256 consecutive points of a warp are a few pixel rows of one 16x16 patch block
// (two layers), i.e. a handful of cells, and consecutive points mostly share their cell.  Instead of sending every
// thread's partial sums to L2 (one table probe + four atomics per thread and cell, ~80 groups per warp) the warp
//   1. walks its points 64 at a time, two consecutive points per lane, merged in registers when they share the cell;
//   2. adds the lanes of a run of equal cells with a segmented shuffle scan (full-mask shuffles, run boundaries from one
//      ballot), so only the last lane of a run holds something to deliver (a second point that lies in another cell
//      than the first is delivered by its lane; running those through a scan of their own measured slower);
//   3. delivers into a 32-slot table of cell accumulators in the warp's slice of SHARED memory: a plain read finds the
//      slot (a CAS only claims an empty one), two shared atomics add {count, sum dx} and {sum dy, sum dz} — sums of the
//      offsets inside the cell (< g <= 64), so two words hold what four would — and the partition extrema are only
//      tracked when the warp's points come from more than one patch;
//   4. after its 256 points flushes every used slot once: one table probe and four atomics per distinct cell.
// A warp whose points span two frames uses the per-thread path (k_accumulate_crossing); records that find the 32 slots
// taken by other cells go to the grid directly.
struct StageSlot {  // structure of arrays per warp: slot s of field f at f[s]
  uint32_t key[32];       // cx | cy << 10 | cz << 20 | 1 << 30, 0 = free
  uint32_t w1[32];        // count | sum dx << 9      (<= 256 points, dx < 64)
  uint32_t w2[32];        // sum dy | sum dz << 16
  uint32_t pmax[32], pmin_inv[32];
};

// one record per lane into the global grid (the tail of the per-thread path): `want` lanes take part
__device__ __forceinline__ void flush_global( const GridArgs& a, int f, bool want, uint32_t key, uint32_t w1, uint32_t w2, uint32_t mx,
                                              uint32_t mn_inv, int lane ) {
  const uint32_t mask = __ballot_sync( 0xFFFFFFFFu, want );
  if ( !want ) { return; }
  const int      cx = key & 1023, cy = ( key >> 10 ) & 1023, cz = ( key >> 20 ) & 1023;
  const uint32_t cnt = w1 & 511u;
  const uint32_t t0 = cnt * (uint32_t)( cx * a.g ) + ( w1 >> 9 ), t1 = cnt * (uint32_t)( cy * a.g ) + ( w2 & 0xFFFFu ),
                 t2 = cnt * (uint32_t)( cz * a.g ) + ( w2 >> 16 );
  const uint32_t bk      = block_key( cx, cy, cz );
  const uint32_t bpeers  = __match_any_sync( mask, bk );
  const int      bleader = __ffs( bpeers ) - 1;
  uint32_t       bl      = NO_BLOCK;
  if ( lane == bleader ) { bl = block_claim( a, f, bk ); }
  bl = __shfl_sync( mask, bl, bleader );
  if ( bl == NO_BLOCK ) { return; }
  Cell* c = a.cells + ( bl * 64u + cell_local( cx, cy, cz ) );
  atomicAdd( (unsigned long long*)&c->cw, (unsigned long long)cnt | ( (unsigned long long)t2 << 32 ) );
  atomicAdd( (unsigned long long*)&c->s0, (unsigned long long)t0 | ( (unsigned long long)t1 << 32 ) );
  atomicMax( &c->pmax, mx );
  atomicMax( &c->pmin_inv, mn_inv );
}

// one record per lane (`want`) into the warp's table
__device__ __forceinline__ void stage_insert( const GridArgs& a, StageSlot& S, int f, int lane, bool onePart, bool want, uint32_t K, uint32_t W1,
                                              uint32_t W2, uint32_t MX, uint32_t MN ) {
  bool spill = false;
  if ( want ) {
    uint32_t h   = ( K * 0x9E3779B1u ) >> 27;
    bool     hit = false;
    for ( int probe = 0; probe < 32 && !hit; probe++ ) {
      uint32_t cur = *reinterpret_cast<volatile uint32_t*>( &S.key[h] );
      if ( cur == 0u ) {
        cur = atomicCAS( &S.key[h], 0u, K );
        if ( cur == 0u ) { cur = K; }
      }
      if ( cur == K ) {
        hit = true;
      } else {
        h = ( h + 1 ) & 31u;
      }
    }
    if ( hit ) {
      atomicAdd( &S.w1[h], W1 );
      atomicAdd( &S.w2[h], W2 );
      if ( !onePart ) {
        atomicMax( &S.pmax[h], MX );
        atomicMax( &S.pmin_inv[h], MN );
      }
    } else {
      spill = true;  // more than 32 distinct cells in this warp's points
    }
  }
  if ( __any_sync( 0xFFFFFFFFu, spill ) ) { flush_global( a, f, spill, K, W1, W2, MX, MN, lane ); }
}

// runs of equal cells over the lanes (K = 0: the lane has nothing) -> their sums into the warp's table
__device__ __forceinline__ void stage_runs( const GridArgs& a, StageSlot& S, int f, int lane, bool onePart, uint32_t K, uint32_t W1, uint32_t W2,
                                            uint32_t MX, uint32_t MN ) {
  const uint32_t prevK    = __shfl_up_sync( 0xFFFFFFFFu, K, 1 );
  const bool     head     = lane == 0 || K != prevK;
  const uint32_t heads    = __ballot_sync( 0xFFFFFFFFu, head );
  const int      segstart = 31 - __clz( heads & ( 0xFFFFFFFFu >> ( 31 - lane ) ) );
  const bool     tail     = lane == 31 || ( ( heads >> ( lane + 1 ) ) & 1u );
#pragma unroll
  for ( int d = 1; d < 32; d <<= 1 ) {
    const uint32_t t1 = __shfl_up_sync( 0xFFFFFFFFu, W1, d ), t2 = __shfl_up_sync( 0xFFFFFFFFu, W2, d );
    if ( lane - d >= segstart ) { W1 += t1, W2 += t2; }
  }
  if ( !onePart ) {
#pragma unroll
    for ( int d = 1; d < 32; d <<= 1 ) {
      const uint32_t tX = __shfl_up_sync( 0xFFFFFFFFu, MX, d ), tN = __shfl_up_sync( 0xFFFFFFFFu, MN, d );
      if ( lane - d >= segstart ) { MX = max( MX, tX ), MN = max( MN, tN ); }
    }
  }
  stage_insert( a, S, f, lane, onePart, tail && K != 0u, K, W1, W2, MX, MN );
}

// the (at most F - 1) warps of the staged kernel whose 256 points span two frames, on the per-thread path: warp j looks
// at the window that holds the end of frame j
template <bool COLOUR>
__global__ void __launch_bounds__( 256 ) k_accumulate_crossing( const GridArgs a, int64_t n ) {
  const int lane = threadIdx.x & 31, j = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
  if ( j >= a.F ) { return; }
  const int64_t e = a.frame_off[j + 1];
  if ( e >= n ) { return; }
  const int64_t w0 = e & ~255ll;
  if ( frame_of( a.frame_off, a.F, w0 ) != j ) { return; }  // the window starts in another frame (or exactly at e)
  accumulate_warp<COLOUR, false>( a, n, w0 + (int64_t)lane * ACC_RUN, lane, j );
}

template <int MINB>
__global__ void __launch_bounds__( 256, MINB ) k_accumulate_geo_staged( const GridArgs a, int64_t n ) {
  __shared__ StageSlot stage[8];
  const int     lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int64_t w0   = ( (int64_t)blockIdx.x * 8 + wi ) * 256;  // first point of the warp
  if ( w0 >= n ) { return; }
  const int fw = frame_of( a.frame_off, a.F, w0 );
  if ( min( w0 + 256, n ) > a.frame_off[fw + 1] ) { return; }  // spans two frames: k_accumulate_crossing
  StageSlot& S = stage[wi];
  S.key[lane] = 0, S.w1[lane] = 0, S.w2[lane] = 0, S.pmax[lane] = 0, S.pmin_inv[lane] = 0;
  // ---- all loads first: iteration k covers points w0 + 64 k + 2 lane, + 1 ----
  uint4 vp[4];
  uint2 vq[4];
#pragma unroll
  for ( int k = 0; k < 4; k++ ) {
    const int64_t i = w0 + 64 * k + 2 * lane;
    if ( i + 1 < n ) {
      vp[k] = *reinterpret_cast<const uint4*>( a.pos + i );
      vq[k] = *reinterpret_cast<const uint2*>( a.part + i );
    } else {
      vp[k] = make_uint4( 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu );  // x = y = z = -1: outside
      vq[k] = make_uint2( 0, 0 );
      if ( i < n ) {
        const short4 p = a.pos[i];
        vp[k].x        = (uint32_t)(uint16_t)p.x | ( (uint32_t)(uint16_t)p.y << 16 );
        vp[k].y        = (uint32_t)(uint16_t)p.z | ( (uint32_t)(uint16_t)p.w << 16 );
        vq[k].x        = a.part[i];
      }
    }
  }
  // the points of a warp nearly always belong to one patch: then max / min partition need neither scans nor atomics
  const uint32_t pFirst = __shfl_sync( 0xFFFFFFFFu, vq[0].x, 0 ) + 1u;
  bool           uni    = true;
#pragma unroll
  for ( int k = 0; k < 4; k++ ) { uni = uni && vq[k].x + 1u == pFirst && vq[k].y + 1u == pFirst; }
  const bool onePart = __all_sync( 0xFFFFFFFFu, uni );
  __syncwarp();
  const int g = a.g, disth = max( g / 2, 1 ), th = grid_th( a, fw );
#pragma unroll
  for ( int k = 0; k < 4; k++ ) {
    // the lane's two points: cell and offsets inside the cell
    const int x0 = (short)( vp[k].x & 0xFFFF ), y0 = (short)( vp[k].x >> 16 ), z0 = (short)( vp[k].y & 0xFFFF );
    const int x1 = (short)( vp[k].z & 0xFFFF ), y1 = (short)( vp[k].z >> 16 ), z1 = (short)( vp[k].w & 0xFFFF );
    const bool in0 = inside( x0, y0, z0, disth, th ), in1 = inside( x1, y1, z1, disth, th );
    const int  cx0 = cell_of( a, x0 & 0xFFFF ), cy0 = cell_of( a, y0 & 0xFFFF ), cz0 = cell_of( a, z0 & 0xFFFF );
    const int  cx1 = cell_of( a, x1 & 0xFFFF ), cy1 = cell_of( a, y1 & 0xFFFF ), cz1 = cell_of( a, z1 & 0xFFFF );
    const uint32_t k0 = in0 ? ( (uint32_t)cx0 | ( (uint32_t)cy0 << 10 ) | ( (uint32_t)cz0 << 20 ) | 0x40000000u ) : 0u;
    const uint32_t k1 = in1 ? ( (uint32_t)cx1 | ( (uint32_t)cy1 << 10 ) | ( (uint32_t)cz1 << 20 ) | 0x40000000u ) : 0u;
    const uint32_t u1 = in0 ? ( 1u | ( (uint32_t)( x0 - cx0 * g ) << 9 ) ) : 0u;
    const uint32_t u2 = in0 ? ( (uint32_t)( y0 - cy0 * g ) | ( (uint32_t)( z0 - cz0 * g ) << 16 ) ) : 0u;
    const uint32_t v1 = in1 ? ( 1u | ( (uint32_t)( x1 - cx1 * g ) << 9 ) ) : 0u;
    const uint32_t v2 = in1 ? ( (uint32_t)( y1 - cy1 * g ) | ( (uint32_t)( z1 - cz1 * g ) << 16 ) ) : 0u;
    const uint32_t p0 = vq[k].x + 1u, p1 = vq[k].y + 1u;
    // stream A: point 0 (with point 1 when it shares the cell), or point 1 alone; stream B: point 1 in another cell
    const bool     same = k0 == k1;                 // (also when both are outside: nothing)
    const bool     bB   = in0 && in1 && !same;
    const bool     add1 = in1 && ( same || !in0 );
    const uint32_t KA   = in0 ? k0 : k1;
    uint32_t       W1 = u1, W2 = u2, MX = in0 ? p0 : 0u, MN = in0 ? ~p0 : 0u;
    if ( add1 ) { W1 += v1, W2 += v2, MX = max( MX, p1 ), MN = max( MN, ~p1 ); }
    stage_runs( a, S, fw, lane, onePart, KA, W1, W2, MX, MN );
    if ( __any_sync( 0xFFFFFFFFu, bB ) ) { stage_insert( a, S, fw, lane, onePart, bB, k1, v1, v2, p1, ~p1 ); }
  }
  __syncwarp();
  // ---- one flush per distinct cell ----
  const uint32_t key = S.key[lane];
  flush_global( a, fw, key != 0u, key, S.w1[lane], S.w2[lane], onePart ? pFirst : S.pmax[lane], onePart ? ~pFirst : S.pmin_inv[lane], lane );
}

// The reference sums the cell members in float, in emission order (:996, :1177).  Integer sums convert to the same float
// while every partial sum stays below 2^24; a cell beyond that is listed and k_ordered_cells repeats
// its float accumulation in emission order.  More than 65535 members wrap the reference's uint16 counter (not reproduced).
__device__ __forceinline__ void note_inexact_cell( const GridArgs& a, uint32_t cell, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t n ) {
  if ( n > 65535u ) {
    a.counters[CTR_RANGE] = 1;
  } else if ( s0 >= ( 1u << 24 ) || s1 >= ( 1u << 24 ) || s2 >= ( 1u << 24 ) ) {
    const uint32_t k = (uint32_t)atomicAdd( &a.counters[CTR_ORDERED], 1 );
    if ( k < ORDERED_CAP ) {
      a.ordered[k] = cell;
    } else {
      a.counters[CTR_RANGE] = 2;
    }
  }
}

// ---- colour: per-cell mean/median gate (:1228-1236, :1239-1243) ----
// One warp per 32 pool cells.  |mean - median| <= standard deviation, so only cells whose luma variance can exceed the
// threshold need the median; those are sorted four at a time with a 32-lane bitonic network in registers.
__device__ __forceinline__ bool gate_of( int vhi, int vlo, int m, uint32_t s0, double mmThresh ) {
  // median (PCCCodec.h:271-278) and mean (:280-285); abs() on the double difference is int abs(int) (App. A.9)
  const double med  = ( m % 2 == 0 ) ? ( (double)vhi + (double)vlo ) / 2.0 : (double)vhi;
  const double mean = (double)s0 / (double)m;
  const int    diff = (int)( mean - med );
  return (double)( diff < 0 ? -diff : diff ) > mmThresh;
}

__global__ void __launch_bounds__( 256 ) k_cell_median_gate( const GridArgs a, double mmThresh ) {
  const int      lane   = threadIdx.x & 31;
  const uint32_t nCells = (uint32_t)min( (unsigned)a.counters[CTR_CURSOR], a.cap_blocks ) * 64u;
  const uint32_t nWarps = ( gridDim.x * blockDim.x ) >> 5;
  const unsigned long long tf = mmThresh >= 1.0 ? (unsigned long long)mmThresh : 0ull;  // floor
  for ( uint32_t c0 = ( ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5 ) * 32u; c0 < nCells; c0 += nWarps * 32u ) {
    Cell*          c    = a.cells + c0 + lane;
    const uint4    lo   = *reinterpret_cast<const uint4*>( c );
    const uint32_t n    = lo.z & FC_CNT;
    bool           want = false;
    if ( n > 1 ) {
      // n * sum(x^2) - sum(x)^2 < floor(T)^2 n^2  =>  sigma < T  =>  |mean - median| < T: the gate is off
      const unsigned __int128 var = (unsigned __int128)n * c->q2 - (unsigned __int128)lo.x * lo.x;
      want                        = var >= (unsigned __int128)( tf * tf ) * ( (unsigned long long)n * n );
      if ( want && n > a.lum_cap ) {  // list truncated: repeat the stage with longer lists
        atomicOr( &a.counters[CTR_FLAGS], OVF_LUM );
        atomicMax( &a.counters[CTR_MAXCNT], (int)n );
        want = false;
      }
    }
    bool     gate  = false;
    uint32_t small = __ballot_sync( 0xFFFFFFFFu, want && n <= 32 );
    uint32_t large = __ballot_sync( 0xFFFFFFFFu, want && n > 32 );
    while ( small ) {  // up to four cells per round: their loads and their sorting networks overlap
      int src[4], m[4], v[4];
#pragma unroll
      for ( int q = 0; q < 4; q++ ) {
        src[q] = small ? __ffs( small ) - 1 : -1;
        small &= small - 1;
        m[q] = src[q] >= 0 ? (int)__shfl_sync( 0xFFFFFFFFu, n, src[q] & 31 ) : 0;
        v[q] = lane < m[q] ? (int)a.lum[(size_t)( c0 + src[q] ) * a.lum_cap + lane] : 0x10000;
      }
#pragma unroll
      for ( int k = 2; k <= 32; k <<= 1 ) {
#pragma unroll
        for ( int j = k >> 1; j > 0; j >>= 1 ) {
          const bool keep_min = ( ( lane & j ) == 0 ) == ( ( lane & k ) == 0 );
#pragma unroll
          for ( int q = 0; q < 4; q++ ) {
            const int u = __shfl_xor_sync( 0xFFFFFFFFu, v[q], j );
            v[q]        = keep_min ? min( v[q], u ) : max( v[q], u );
          }
        }
      }
#pragma unroll
      for ( int q = 0; q < 4; q++ ) {
        const int vhi = __shfl_sync( 0xFFFFFFFFu, v[q], ( m[q] / 2 ) & 31 ), vlo = __shfl_sync( 0xFFFFFFFFu, v[q], ( m[q] / 2 - 1 ) & 31 );
        if ( lane == src[q] ) { gate = gate_of( vhi, vlo, m[q], lo.x, mmThresh ); }
      }
    }
    for ( ; large; large &= large - 1 ) {  // rare: more than 32 points in a cell, rank selection
      const int       src = __ffs( large ) - 1;
      const int       m   = (int)__shfl_sync( 0xFFFFFFFFu, n, src );
      const uint16_t* L   = a.lum + (size_t)( c0 + src ) * a.lum_cap;
      const int       hiR = m / 2, loR = m / 2 - 1;
      int             vhi = -1, vlo = -1;
      for ( int i = lane; i < m; i += 32 ) {
        const int v    = L[i];
        int       rank = 0;
        for ( int j = 0; j < m; j++ ) {
          const int u = L[j];
          rank += ( u < v ) || ( u == v && j < i );
        }
        if ( rank == hiR ) { vhi = v; }
        if ( rank == loR ) { vlo = v; }
      }
#pragma unroll
      for ( int d = 16; d > 0; d >>= 1 ) {
        vhi = max( vhi, __shfl_xor_sync( 0xFFFFFFFFu, vhi, d ) );
        vlo = max( vlo, __shfl_xor_sync( 0xFFFFFFFFu, vlo, d ) );
      }
      if ( lane == src ) { gate = gate_of( vhi, vlo, m, lo.x, mmThresh ); }
    }
    if ( n > 0 ) {  // finalise in place: mean colour (:1225: float accumulator read back as double / count), flags
      note_inexact_cell( a, c0 + lane, lo.x, lo.y, lo.w, n );
      const uint2    pm = *reinterpret_cast<const uint2*>( &c->pmax );
      const uint32_t cw = n | ( pm.x != ~pm.y ? FC_MULTI : 0u ) | ( gate ? FC_GATE : 0u );
      const double   dn = (double)n;
      double2*       o  = reinterpret_cast<double2*>( c );
      o[0]              = make_double2( (double)(float)lo.x / dn, (double)(float)lo.y / dn );
      o[1]              = make_double2( (double)(float)lo.w / dn, __longlong_as_double( (long long)cw ) );
    }
  }
}

// geometry: finalise the cells in place (centre = float sum / float count, :135-137; doSmooth flag)
__global__ void __launch_bounds__( 256 ) k_finalize_geo( const GridArgs a ) {
  const uint32_t nCells = (uint32_t)min( (unsigned)a.counters[CTR_CURSOR], a.cap_blocks ) * 64u;
  for ( uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nCells; s += gridDim.x * blockDim.x ) {
    Cell*          c  = a.cells + s;
    const uint4    lo = *reinterpret_cast<const uint4*>( c );
    const uint32_t n  = lo.z & FC_CNT;
    if ( n == 0 ) { continue; }
    note_inexact_cell( a, s, lo.x, lo.y, lo.w, n );
    const uint2 pm  = *reinterpret_cast<const uint2*>( &c->pmax );
    const float fcn = (float)n;
    *reinterpret_cast<uint4*>( c ) =
        make_uint4( __float_as_uint( __fdiv_rn( (float)lo.x, fcn ) ), __float_as_uint( __fdiv_rn( (float)lo.y, fcn ) ),
                    n | ( pm.x != ~pm.y ? FC_MULTI : 0u ), __float_as_uint( __fdiv_rn( (float)lo.w, fcn ) ) );
  }
}

// ordered float accumulation of the listed cells (one warp per cell): the frame's points are walked in emission order,
// 32 at a time; the members of the cell are added one after the other in float, exactly as addGridCentroid /
// addGridColorCentroid do (:996, :1177), and the cell's finalised record is rewritten from those sums
template <bool COLOUR>
__global__ void __launch_bounds__( 256 ) k_ordered_cells( const GridArgs a ) {
  const int      lane = threadIdx.x & 31;
  const uint32_t k    = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
  if ( k >= min( (uint32_t)a.counters[CTR_ORDERED], ORDERED_CAP ) ) { return; }
  const uint32_t cell = a.ordered[k], bl = cell >> 6, loc = cell & 63u;
  const uint2    bi   = a.binfo[bl];
  const int      f    = (int)bi.x;
  const uint32_t bk   = bi.y - 1u;
  const int      cx = (int)( ( bk & 0xFFu ) << 2 | ( loc & 3u ) ), cy = (int)( ( ( bk >> 8 ) & 0xFFu ) << 2 | ( ( loc >> 2 ) & 3u ) ),
                 cz = (int)( ( ( bk >> 16 ) & 0xFFu ) << 2 | ( loc >> 4 ) );
  const int      disth = max( a.g / 2, 1 ), th = COLOUR ? 0 : grid_th( a, f );
  float          s0 = 0.f, s1 = 0.f, s2 = 0.f;
  uint32_t       n  = 0;
  for ( int64_t i0 = a.frame_off[f]; i0 < a.frame_off[f + 1]; i0 += 32 ) {
    const int64_t i = i0 + lane;
    bool          in = false;
    float         v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if ( i < a.frame_off[f + 1] ) {
      const short4 q = a.pos[i];
      in = COLOUR ? ( q.x >= 0 && q.y >= 0 && q.z >= 0 ) : inside( q.x, q.y, q.z, disth, th );
      in = in && cell_of( a, q.x ) == cx && cell_of( a, q.y ) == cy && cell_of( a, q.z ) == cz;
      if ( in ) {
        if ( COLOUR ) {
          const ushort4 cv = a.col[i];
          v0 = (float)cv.x, v1 = (float)cv.y, v2 = (float)cv.z;
        } else {
          v0 = (float)q.x, v1 = (float)q.y, v2 = (float)q.z;
        }
      }
    }
    for ( uint32_t m = __ballot_sync( 0xFFFFFFFFu, in ); m; m &= m - 1 ) {  // members of this chunk, in order
      const int src = __ffs( m ) - 1;
      s0 = __fadd_rn( s0, __shfl_sync( 0xFFFFFFFFu, v0, src ) );
      s1 = __fadd_rn( s1, __shfl_sync( 0xFFFFFFFFu, v1, src ) );
      s2 = __fadd_rn( s2, __shfl_sync( 0xFFFFFFFFu, v2, src ) );
      n++;
    }
  }
  if ( lane != 0 ) { return; }
  Cell* c = a.cells + cell;
  if ( COLOUR ) {  // {double mean c0, c1, c2, cw'}: the flags of k_cell_median_gate stay
    double2*     o  = reinterpret_cast<double2*>( c );
    const double dn = (double)n;
    const double cw = o[1].y;
    o[0]            = make_double2( (double)s0 / dn, (double)s1 / dn );
    o[1]            = make_double2( (double)s2 / dn, cw );
  } else {  // {float centre x, y, cw', centre z}
    uint4*      o   = reinterpret_cast<uint4*>( c );
    const float fcn = (float)n;
    o->x = __float_as_uint( __fdiv_rn( s0, fcn ) ), o->y = __float_as_uint( __fdiv_rn( s1, fcn ) ), o->w = __float_as_uint( __fdiv_rn( s2, fcn ) );
  }
}

// ---- geometry filter: smoothPointCloudGrid + gridFiltering (:1000-1104) ----
// per-frame statistics without same-address atomic storms: one atomic per warp and frame
__device__ __forceinline__ void count_hits( int32_t* field0, int f, bool hit ) {  // field0 = &finfo[0].<field>
  const uint32_t m = __ballot_sync( 0xFFFFFFFFu, hit );
  if ( hit ) {
    const uint32_t peers = __match_any_sync( m, f );
    if ( ( threadIdx.x & 31 ) == __ffs( peers ) - 1 ) {
      atomicAdd( field0 + (size_t)f * ( sizeof( RbFrameInfo ) / sizeof( int32_t ) ), __popc( peers ) );
    }
  }
}

// G: the cell size when it is known at compile time (8: CTC gridSize), 0: a.g
template <int G>
__device__ __forceinline__ bool filter_geo_point( const GridArgs& a, uint32_t li, double threshold, int& f ) {
  if ( li >= *a.blist_n || a.counters[CTR_FLAGS] ) { return false; }  // tables overflowed: the stage is repeated
  const int64_t i = a.blist[li];
  const short4  p = a.pos[i];
  if ( p.w != 1 ) { return false; }  // :1087
  f               = frame_of( a.frame_off, a.F, i );
  const int g     = G ? G : a.g, hg = g / 2;
  const int disth = max( hg, 1 ), th = grid_th( a, f );
  if ( !inside( p.x, p.y, p.z, disth, th ) ) { return false; }  // :1078-1081
  const int      P[3] = {p.x, p.y, p.z};
  int            S[3];
  for ( int k = 0; k < 3; k++ ) { S[k] = P[k] / g + ( ( P[k] - ( P[k] / g ) * g < hg ) ? -1 : 0 ); }  // :1014-1017
  CellView<false> fc[8];
  bool     other = false;
  uint32_t cnt[8];
  cell_find8<false>( a, f, S, fc );
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {
    cnt[k] = fc[k].cnt;
    if ( fc[k].multi ) { other = true; }  // doSmooth && count (:1024)
  }
  if ( !other ) { return false; }  // :1028
  const int    g2 = 2 * g;
  int          Wt[3], Q[3];
  for ( int k = 0; k < 3; k++ ) {
    Wt[k] = ( P[k] - S[k] * g - hg ) * 2 + 1;  // :1034
    Q[k]  = g2 - Wt[k];                         // :1047
  }
  double c4[3] = {0.0, 0.0, 0.0};
  int    count = 0;
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {  // :1050-1058
    const int    dx = k & 1, dy = ( k >> 1 ) & 1, dz = k >> 2;
    const int    wgt = ( dx ? Wt[0] : Q[0] ) * ( dy ? Wt[1] : Q[1] ) * ( dz ? Wt[2] : Q[2] );
    double       v[3];
    if ( cnt[k] > 0 ) {  // :1040: the cell centre (k_finalize_geo)
      v[0] = (double)fc[k].m0, v[1] = (double)fc[k].m1, v[2] = (double)fc[k].m2;
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
  if ( count == 0 ) { return false; }  // 0.0/0.0 = NaN, NaN >= x is false: the reference leaves the point alone
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
    atomicOr( &a.moved_bits[i >> 5], 1u << ( i & 31 ) );  // the list of the moved points for the attribute re-transfer
    return true;
  }
  return false;
}
template <int G>
__global__ void __launch_bounds__( 128, 8 ) k_filter_geo( const GridArgs a, double threshold ) {
  int        f   = 0;
  const bool hit = filter_geo_point<G>( a, blockIdx.x * blockDim.x + threadIdx.x, threshold, f );
  count_hits( &a.finfo[0].smoothed, f, hit );
}

// ---- colour filter: smoothPointCloudColorLC + gridFilteringColor (:1182-1306) ----
template <int G>
__device__ __forceinline__ bool filter_col_point( const GridArgs& a, uint32_t li, double thrSmoothing, double yThresh, int& f ) {
  if ( li >= *a.blist_n || a.counters[CTR_FLAGS] ) { return false; }  // tables overflowed: the stage is repeated
  const int64_t i = a.blist[li];
  const short4  p = a.pos[i];
  if ( p.w != 1 ) { return false; }  // :1288
  const int g = G ? G : a.g, hg = g / 2, disth = max( hg, 1 );
  if ( !inside( p.x, p.y, p.z, disth, a.pcmax ) ) { return false; }  // :1280-1283
  f = frame_of( a.frame_off, a.F, i );
  const int      P[3] = {p.x, p.y, p.z};
  int            S[3];
  for ( int k = 0; k < 3; k++ ) { S[k] = P[k] / g + ( ( ( P[k] % g ) < hg ) ? -1 : 0 ); }  // :1197-1199
  const ushort4 cv = a.col[i];  // issued with the probes: independent of them
  CellView<true> fc[8];
  bool          other = false;
  cell_find8<true>( a, f, S, fc );
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {
    if ( fc[k].multi ) { other = true; }  // :1204
  }
  if ( !other ) { return false; }  // :1210
  const double  cur[3] = {(double)cv.x, (double)cv.y, (double)cv.z};
  int           Wt[3], Q[3];
  const int     g2 = 2 * g;
  for ( int k = 0; k < 3; k++ ) {
    Wt[k] = ( P[k] - S[k] * g - hg ) * 2 + 1;  // :1212
    Q[k]  = g2 - Wt[k];                         // :1252
  }
  double c4[3]    = {0.0, 0.0, 0.0};
  double Y0       = 0.0;
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {  // :1218-1261, loop order dz, dy, dx; the blend is accumulated in the same order
    const uint32_t cn   = fc[k].cnt;
    const bool     gate = cn > 1 && fc[k].gate;
    double         d[3];
    bool           own = cn == 0;
    if ( cn > 0 ) {
      d[0] = fc[k].m0, d[1] = fc[k].m1, d[2] = fc[k].m2;  // :1225, divided once per cell by k_cell_median_gate
      if ( k == 0 ) {
        if ( gate ) { return false; }  // :1228-1235: centroid = own colour -> |dY| = 0 < threshold unless threshold <= 0
      } else {
        const int dy0 = (int)( Y0 - d[0] );  // abs() truncates (App. A.9), :1238
        own           = (double)( dy0 < 0 ? -dy0 : dy0 ) > yThresh || gate;  // :1238-1243
      }
    }
    if ( own ) { d[0] = cur[0], d[1] = cur[1], d[2] = cur[2]; }
    if ( k == 0 ) { Y0 = d[0]; }  // :1248
    const int    dx = k & 1, dy = ( k >> 1 ) & 1, dz = k >> 2;
    const double dw = (double)( ( dx ? Wt[0] : Q[0] ) * ( dy ? Wt[1] : Q[1] ) * ( dz ? Wt[2] : Q[2] ) );
    c4[0]           = __dadd_rn( c4[0], __dmul_rn( d[0], dw ) );
    c4[1]           = __dadd_rn( c4[1], __dmul_rn( d[1], dw ) );
    c4[2]           = __dadd_rn( c4[2], __dmul_rn( d[2], dw ) );
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
    a.col[i] = q;
    return q.x != cv.x || q.y != cv.y || q.z != cv.z;
  }
  return false;
}
template <int G>
__global__ void __launch_bounds__( 128, 8 ) k_filter_col( const GridArgs a, double thrSmoothing, double yThresh ) {
  int        f   = 0;
  const bool hit = filter_col_point<G>( a, blockIdx.x * blockDim.x + threadIdx.x, thrSmoothing, yThresh, f );
  count_hits( &a.finfo[0].recolored, f, hit );
}

// ---- cleanup: reset the used part of the pool ----
__global__ void __launch_bounds__( 256 ) k_cleanup_cells( const GridArgs a ) {
  const uint32_t nCells = (uint32_t)min( (unsigned)a.counters[CTR_CURSOR], a.cap_blocks ) * 64u;
  for ( uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nCells; s += gridDim.x * blockDim.x ) {
    uint4*      c  = reinterpret_cast<uint4*>( a.cells + s );
    const uint4 lo = c[0], hi = c[1];
    if ( ( lo.x | lo.y | lo.z | lo.w | hi.x | hi.y | hi.z | hi.w ) == 0 ) { continue; }
    c[0] = c[1] = make_uint4( 0, 0, 0, 0 );
  }
}

// ---- convertYUV16ToRGB8 (PCCPointSet.h:133-166) / copyRGB16ToRGB8 (:121-127) ----
// The reference's double arithmetic, operation by operation (the fallback of the short path below).
__device__ __forceinline__ uchar4 yuv16_to_rgb8_f64( ushort4 c ) {
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
  return make_uchar4( (unsigned char)r, (unsigned char)g, (unsigned char)b, 0 );
}

// The same function with a third of the double operations.  In exact arithmetic (with the clamps: only chroma 0 is
// clamped, to -0.5, i.e. 2 (c - 32768) = -65535) a channel is  255 (Y + k_u U2 / 2 + k_v V2 / 2) / 65535
// = (Y + k_u/2 U2 + k_v/2 V2) / 257 = num / 257000000 with integer num, so two different values are >= 3.9e-9
// apart and a value that is not EXACTLY a tie k + 1/2 is at least that far from one.  Both the reference's sequence
// of operations and the short sequence below are within 1e-12 of the exact value; hence wherever the short result
// is further than 1e-9 from a tie, round-half-away of the reference sequence is the nearest integer of the short
// result.  Only exact ties (a few colours per million million) take the reference sequence.
__device__ __forceinline__ int round_channel( double x, bool& near_tie ) {
  const double M  = 4503599627370496.0;                 // 2^52: adding it rounds to the nearest integer
  const double xc = fmin( fmax( x, 0.0 ), 255.25 );     // PCCClip after round; ties 0.5 .. 254.5 are kept
  const double z  = __dadd_rn( xc, M );
  const double e  = __dsub_rn( xc, __dsub_rn( z, M ) );  // exact: distance to the nearest integer
  near_tie |= fabs( e ) > 0.499999999;
  return __double2loint( z );
}
__device__ __forceinline__ uchar4 yuv16_to_rgb8( ushort4 c ) {
  const double M   = 4503599627370496.0;
  const int    u2i = max( 2 * ( (int)c.y - 32768 ), -65535 ), v2i = max( 2 * ( (int)c.z - 32768 ), -65535 );
  const double Y   = __dsub_rn( __hiloint2double( 0x43300000, (int)c.x ), M );  // exact int -> double without I2F
  const double U2  = __dsub_rn( __hiloint2double( 0x43300000, u2i + 65536 ), M + 65536.0 );
  const double V2  = __dsub_rn( __hiloint2double( 0x43300000, v2i + 65536 ), M + 65536.0 );
  const double inv = 1.0 / 257.0;
  bool         near_tie = false;
  const int    r = round_channel( __dmul_rn( __fma_rn( 1.57480 * 0.5, V2, Y ), inv ), near_tie );
  const int    g = round_channel( __dmul_rn( __fma_rn( -0.46813 * 0.5, V2, __fma_rn( -0.18733 * 0.5, U2, Y ) ), inv ), near_tie );
  const int    b = round_channel( __dmul_rn( __fma_rn( 1.85563 * 0.5, U2, Y ), inv ), near_tie );
  if ( near_tie ) { return yuv16_to_rgb8_f64( c ); }
  return make_uchar4( (unsigned char)r, (unsigned char)g, (unsigned char)b, 0 );
}

__device__ __forceinline__ uchar4 to_rgb8_one( ushort4 c, int rgb444, int attr_count, int force_f64 ) {
  if ( attr_count == 0 ) { return make_uchar4( 127, 127, 127, 0 ); }  // PCCCodec.cpp:1327-1330
  if ( rgb444 ) { return make_uchar4( (unsigned char)c.x, (unsigned char)c.y, (unsigned char)c.z, 0 ); }
  return force_f64 ? yuv16_to_rgb8_f64( c ) : yuv16_to_rgb8( c );
}
// four points per thread: two 16-byte loads, one 16-byte store
__global__ void __launch_bounds__( 256 ) k_to_rgb8( const ushort4* __restrict__ col, uchar4* __restrict__ rgb, int64_t n,
                                                     int rgb444, int attr_count, int force_f64 ) {
  const int64_t i0 = ( blockIdx.x * (int64_t)blockDim.x + threadIdx.x ) * 4;
  if ( i0 >= n ) { return; }
  if ( i0 + 4 <= n ) {
    const uint4 a = *reinterpret_cast<const uint4*>( col + i0 ), b = *reinterpret_cast<const uint4*>( col + i0 + 2 );
    const ushort4 c[4] = {make_ushort4( a.x & 0xFFFF, a.x >> 16, a.y & 0xFFFF, a.y >> 16 ), make_ushort4( a.z & 0xFFFF, a.z >> 16, a.w & 0xFFFF, a.w >> 16 ),
                          make_ushort4( b.x & 0xFFFF, b.x >> 16, b.y & 0xFFFF, b.y >> 16 ), make_ushort4( b.z & 0xFFFF, b.z >> 16, b.w & 0xFFFF, b.w >> 16 )};
    uint32_t      o[4];
#pragma unroll
    for ( int k = 0; k < 4; k++ ) {
      const uchar4 q = to_rgb8_one( c[k], rgb444, attr_count, force_f64 );
      o[k]           = (uint32_t)q.x | ( (uint32_t)q.y << 8 ) | ( (uint32_t)q.z << 16 );
    }
    *reinterpret_cast<uint4*>( rgb + i0 ) = make_uint4( o[0], o[1], o[2], o[3] );
  } else {
    for ( int64_t i = i0; i < n; i++ ) { rgb[i] = to_rgb8_one( col[i], rgb444, attr_count, force_f64 ); }
  }
}

// table geometry + (re)allocation; table and pool are all-zero between calls (the cleanup pass resets what was used)
struct GridBufs {
  RbBuf &table, &cells, &counters, &lum, &marks, &binfo;
};
int setup_grid( rb200_ctx* c, GridArgs& a, GridBufs b, int g, int wmax, bool colour, int grow ) {
  if ( wmax > 1024 ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "smoothing grid wider than 1024 cells per axis" ); }
  const int64_t n        = c->h_frame_off[c->F];
  int64_t       maxFrame = 1;
  for ( int f = 0; f < c->F; f++ ) { maxFrame = std::max<int64_t>( maxFrame, c->h_frame_off[f + 1] - c->h_frame_off[f] ); }
  // a 4x4x4 block of cells spans (4g)^3 voxels; a surface crossing it leaves ~ (4g)^2 .. 3 (4g)^2 points in it, so
  // n / (6 g^2) blocks is about twice what a V-PCC cloud needs (`grow` quadruples it after an overflow)
  const int64_t perBlock = std::max<int64_t>( 1, 6ll * g * g >> ( 2 * grow ) );
  int64_t       cap      = std::min<int64_t>( n, n / perBlock + 1024ll * c->F ) + 64;
  int64_t       wantT    = 4 * std::min<int64_t>( maxFrame, maxFrame / perBlock + 1024 );
  // test hook (rb200_debug_set_grid_shrink): start with 2^k times smaller tables and 4-entry luma lists, so that the
  // overflow -> regrow -> repeat path runs on small inputs (tests/test_gpu_parity.py)
  const int shrink = c->test_grid_shrink;
  if ( shrink ) {
    cap   = std::max<int64_t>( 8, ( cap >> shrink ) << ( 2 * grow ) );
    wantT = std::max<int64_t>( 16, ( wantT >> shrink ) << ( 2 * grow ) );
  }
  uint32_t tslots = shrink ? 16 : 1024;
  int      tlog   = shrink ? 4 : 10;
  while ( (int64_t)tslots < wantT ) { tslots <<= 1, tlog++; }
  const size_t tBytes = (size_t)c->F * tslots * 8, cBytes = (size_t)cap * 64 * sizeof( Cell );
  if ( cap >= ( 1ll << 25 ) ) { return rb_fail( c, RB200_ERR_NOMEM, "smoothing cell pool too large" ); }
  if ( b.table.cap < tBytes ) {
    RB_CUDA( b.table.ensure( tBytes ) );
    RB_CUDA( cudaMemsetAsync( b.table.p, 0, b.table.cap, c->stream ) );
  }
  if ( b.cells.cap < cBytes ) {
    RB_CUDA( b.cells.ensure( cBytes ) );
    RB_CUDA( cudaMemsetAsync( b.cells.p, 0, b.cells.cap, c->stream ) );
  }
  a.lum_cap = 0;
  if ( colour ) {  // a cell of g^3 voxels seen by two layers (+ duplicates across patches)
    const int64_t base = shrink ? 4 : std::min<int64_t>( 64, std::max<int64_t>( 8, (int64_t)g * g * g ) );
    a.lum_cap          = (uint32_t)std::max<int64_t>( base, c->col_lum_want );
    RB_CUDA( b.lum.ensure( (size_t)cap * 64 * a.lum_cap * 2 + 64 ) );
  }
  RB_CUDA( b.counters.ensure( 64 + ORDERED_CAP * 4 ) );
  RB_CUDA( cudaMemsetAsync( b.counters.p, 0, 64, c->stream ) );
  RB_CUDA( b.binfo.ensure( (size_t)cap * 8 ) );
  // mark bitmap: dense, one bit per cell; beyond 1 GiB per GOF (cells of 1 or 2 voxels at 11+ bits) every cell is
  // accumulated instead
  a.ginv   = g > 1 ? (uint32_t)( ( 1ull << 32 ) / (unsigned)g ) + 1u : 0u;
  a.mwords = ( wmax + 1 + 31 ) / 32 + 1;
  a.marks  = nullptr;
  const size_t mBytes = (size_t)c->F * wmax * wmax * a.mwords * 4;
  if ( colour && c->blist_cap > 0 && mBytes <= ( 1ull << 30 ) ) {  // (geometry, 8-voxel cells: measured no gain)
    RB_CUDA( b.marks.ensure( mBytes ) );
    RB_CUDA( cudaMemsetAsync( b.marks.p, 0, mBytes, c->stream ) );
    a.marks = b.marks.as<uint32_t>();
  }
  a.table      = b.table.as<unsigned long long>();
  a.tslots     = tslots;
  a.tshift     = 32 - tlog;
  a.cells      = b.cells.as<Cell>();
  a.cap_blocks = (uint32_t)cap;
  a.lum        = b.lum.as<uint16_t>();
  a.counters   = b.counters.as<int32_t>();
  a.ordered    = b.counters.as<uint32_t>() + 16;
  a.binfo      = b.binfo.as<uint2>();
  return RB200_OK;
}

// the counters, read after the stage has been enqueued (one small read-back, also the stage's error check):
// 0 = done, 1 = tables overflowed (nothing was modified: repeat with larger tables), < 0 = error
int stage_result( rb200_ctx* c, const GridArgs& a, GridBufs b, const char* what ) {
  int32_t* h = (int32_t*)rb_pinned( c, 64 );
  if ( !h ) { return -rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  RB_CUDA( cudaMemcpyAsync( h, a.counters, 32, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  c->stats.d2h_bytes += 32;
  if ( h[CTR_FLAGS] & OVF_SPIN ) { return -rb_fail( c, RB200_ERR_CUDA, "%s: block table wait timed out", what ); }
  if ( h[CTR_FLAGS] ) {
    if ( h[CTR_FLAGS] & OVF_LUM ) { c->col_lum_want = std::max<int64_t>( c->col_lum_want, ( (int64_t)h[CTR_MAXCNT] + 15 ) & ~15ll ); }
    return ( h[CTR_FLAGS] & OVF_BLOCKS ) ? 1 : 2;
  }
  if ( h[CTR_RANGE] ) {
    return -rb_fail( c, RB200_ERR_UNSUPPORTED, h[CTR_RANGE] == 1 ? "%s: a cell holds more than 65535 points (the reference's uint16 counter wraps)"
                                                                  : "%s: too many cells whose float sums need ordered accumulation", what );
  }
  return 0;
}

constexpr int WALK_CTAS = 148 * 8;  // grid-stride passes over the used part of the pool

}  // namespace

int rb_smooth_geometry_impl( rb200_ctx* c ) {
  const rb200_params& P = c->P;
  const int64_t       n = c->h_frame_off[c->F];
  if ( n == 0 || !P.flag_geometry_smoothing ) { return RB200_OK; }  // :64-65
  if ( !P.grid_smoothing ) {  // :140-142 (never reached from the decoder, PCCDecoder.cpp:436: the encoder's reconstruction)
    return P.pbf_enable ? RB200_OK : rb_smooth_radius_impl( c );
  }
  const int g = P.grid_size;
  if ( g < 1 || g > 64 ) { return rb_fail( c, RB200_ERR_INVALID, "grid_size %d out of range", g ); }
  const int pcmax = 1 << P.geometry_bitdepth_3d;
  const int wmax  = ( pcmax + g - 1 ) / g + 1;
  if ( P.attribute_count > 0 && P.attr_transfer_filter_type == 1 ) {
    // tempFrameBuffer = reconstruct (PCCDecoder.cpp:435): the colour transfer needs the pre-smoothing cloud
    RB_CUDA( c->d_pos_pre.ensure( (size_t)n * 8 ) );
    RB_CUDA( cudaMemcpyAsync( c->d_pos_pre.p, c->d_pos.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream ) );
    c->pos_pre_valid = true;
  }
  GridArgs a{};
  a.F         = c->F;
  a.g         = g;
  a.wmax      = wmax;
  a.by_bbox   = 1;
  a.pcmax     = pcmax;
  a.frame_off = c->d_frame_off.as<int64_t>();
  a.finfo     = c->d_frame_info.as<RbFrameInfo>();
  a.pos       = c->d_pos.as<short4>();
  a.col       = c->d_col.as<ushort4>();
  a.part      = c->d_part.as<uint32_t>();
  a.blist     = c->d_blist.as<uint32_t>();
  a.blist_n   = c->d_blist_n.as<uint32_t>();
  a.moved_bits = c->d_moved_bits.as<uint32_t>();
  GridBufs b{c->d_geo_grid, c->d_geo_cells, c->d_scratch[2], c->d_col_lum, c->d_geo_cell_ids, c->d_scratch[5]};
  for ( ;; ) {
    int r = setup_grid( c, a, b, g, wmax, false, c->geo_grow );
    if ( r ) { return r; }
    if ( a.marks ) { RB_LAUNCH( "geo_mark", k_mark_cells, rb_div_up( c->blist_cap, 256 ), 256, 0, a ); }
    // per-warp staging in shared memory; 5 CTAs per SM, 48 registers: 0.195 ms on the 32-frame vox10 GOF (4 CTAs, 60
    // registers: 0.199; the per-thread path k_accumulate<false> for every point: 0.235)
    RB_LAUNCH( "geo_accumulate", k_accumulate_geo_staged<5>, rb_div_up( n, 2048 ), 256, 0, a, n );
    if ( c->F > 1 ) { RB_LAUNCH( "geo_accumulate_x", k_accumulate_crossing<false>, rb_div_up( c->F, 8 ), 256, 0, a, n ); }
    RB_LAUNCH( "geo_finalize", k_finalize_geo, WALK_CTAS, 256, 0, a );
    RB_LAUNCH( "geo_ordered", k_ordered_cells<false>, ORDERED_CAP / 8, 256, 0, a );
    if ( c->blist_cap > 0 ) {
      auto kf = g == 8 ? k_filter_geo<8> : k_filter_geo<0>;
      RB_LAUNCH( "geo_filter", kf, rb_div_up( c->blist_cap, 128 ), 128, 0, a, P.threshold_smoothing );
    }
    RB_LAUNCH( "geo_cleanup", k_cleanup_cells, WALK_CTAS, 256, 0, a );
    RB_CUDA( cudaMemsetAsync( a.table, 0, (size_t)a.F * a.tslots * 8, c->stream ) );
    r = stage_result( c, a, b, "geometry smoothing" );
    if ( r < 0 ) { return -r; }
    if ( r == 0 ) { return RB200_OK; }
    if ( ++c->geo_grow > 6 ) { return rb_fail( c, RB200_ERR_NOMEM, "geometry smoothing: cell table overflow" ); }
  }
}

int rb_smooth_color_impl( rb200_ctx* c ) {
  const rb200_params& P = c->P;
  const int64_t       n = c->h_frame_off[c->F];
  if ( n == 0 || P.attribute_count == 0 ) { return RB200_OK; }
  const int g     = P.occupancy_precision;  // the colour grid uses occupancyPrecision, not cgridSize (:152)
  const int pcmax = 1 << P.geometry_bitdepth_3d;
  const int wmax  = pcmax / g;  // :154
  if ( wmax < 2 || g < 1 || g > 64 ) { return rb_fail( c, RB200_ERR_INVALID, "colour grid degenerate (occupancy precision %d)", g ); }
  GridArgs a{};
  a.F         = c->F;
  a.g         = g;
  a.wmax      = wmax;
  a.by_bbox   = 0;
  a.pcmax     = pcmax;
  a.frame_off = c->d_frame_off.as<int64_t>();
  a.finfo     = c->d_frame_info.as<RbFrameInfo>();
  a.pos       = c->d_pos.as<short4>();
  a.col       = c->d_col.as<ushort4>();
  a.part      = c->d_part.as<uint32_t>();
  a.blist     = c->d_blist.as<uint32_t>();
  a.blist_n   = c->d_blist_n.as<uint32_t>();
  GridBufs b{c->d_col_grid, c->d_col_cells, c->d_scratch[3], c->d_col_lum, c->d_col_cell_ids, c->d_scratch[6]};
  for ( int attempt = 0;; attempt++ ) {
    int r = setup_grid( c, a, b, g, wmax, true, c->col_grow );
    if ( r ) { return r; }
    if ( a.marks ) { RB_LAUNCH( "col_mark", k_mark_cells, rb_div_up( c->blist_cap, 256 ), 256, 0, a ); }
    RB_LAUNCH( "col_accumulate", k_accumulate<true>, rb_div_up( n, 256 * ACC_RUN ), 256, 0, a, n );
    RB_LAUNCH( "col_median_gate", k_cell_median_gate, WALK_CTAS, 256, 0, a, P.threshold_color_variation * 256.0 );
    RB_LAUNCH( "col_ordered", k_ordered_cells<true>, ORDERED_CAP / 8, 256, 0, a );
    if ( c->blist_cap > 0 ) {
      auto kf = g == 4 ? k_filter_col<4> : ( g == 2 ? k_filter_col<2> : k_filter_col<0> );
      RB_LAUNCH( "col_filter", kf, rb_div_up( c->blist_cap, 128 ), 128, 0, a, P.threshold_color_smoothing,
                 P.threshold_color_difference * 256.0 );
    }
    RB_LAUNCH( "col_cleanup", k_cleanup_cells, WALK_CTAS, 256, 0, a );
    RB_CUDA( cudaMemsetAsync( a.table, 0, (size_t)a.F * a.tslots * 8, c->stream ) );
    r = stage_result( c, a, b, "colour smoothing" );
    if ( r < 0 ) { return -r; }
    if ( r == 0 ) { return RB200_OK; }
    if ( r == 1 && ++c->col_grow > 6 ) { return rb_fail( c, RB200_ERR_NOMEM, "colour smoothing: cell table overflow" ); }
    if ( attempt > 8 ) { return rb_fail( c, RB200_ERR_NOMEM, "colour smoothing: luma lists do not fit" ); }
  }
}

int rb_convert_rgb8_impl( rb200_ctx* c ) {
  const int64_t n = c->h_frame_off[c->F];
  if ( n == 0 ) { return RB200_OK; }
  RB_LAUNCH( "to_rgb8", k_to_rgb8, rb_div_up( n, 1024 ), 256, 0, c->d_col.as<ushort4>(), c->d_rgb.as<uchar4>(), n,
             c->P.attribute_rgb444, c->P.attribute_count, 0 );
  return RB200_OK;
}

// test hook (rb200_debug_yuv16_to_rgb8): n colour triples through the production conversion kernel (integer path with
// double fallback) or, force_f64 != 0, through the double path alone
static int debug_rgb8_run( rb200_ctx* c, RbBuf& in, RbBuf& out, const uint16_t* yuv, int64_t n, uint8_t* rgb, int force_f64 ) {
  RB_CUDA( in.ensure( (size_t)n * 8 + 16 ) );
  RB_CUDA( out.ensure( (size_t)n * 4 + 16 ) );
  std::vector<uint16_t> h4( (size_t)n * 4 );
  for ( int64_t i = 0; i < n; i++ ) { h4[4 * i] = yuv[3 * i], h4[4 * i + 1] = yuv[3 * i + 1], h4[4 * i + 2] = yuv[3 * i + 2], h4[4 * i + 3] = 0; }
  RB_CUDA( cudaMemcpyAsync( in.p, h4.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream ) );
  RB_LAUNCH( "to_rgb8", k_to_rgb8, rb_div_up( n, 1024 ), 256, 0, in.as<ushort4>(), out.as<uchar4>(), n, 0, 1, force_f64 );
  std::vector<uint8_t> o4( (size_t)n * 4 );
  RB_CUDA( cudaMemcpyAsync( o4.data(), out.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  for ( int64_t i = 0; i < n; i++ ) { rgb[3 * i] = o4[4 * i], rgb[3 * i + 1] = o4[4 * i + 1], rgb[3 * i + 2] = o4[4 * i + 2]; }
  return RB200_OK;
}
int rb_debug_rgb8_impl( rb200_ctx* c, const uint16_t* yuv, int64_t n, uint8_t* rgb, int force_f64 ) {
  RbBuf     in, out;
  const int r = debug_rgb8_run( c, in, out, yuv, n, rgb, force_f64 );
  in.release();  // on every path
  out.release();
  return r;
}
