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

struct Cell {  // accumulator of one occupied cell; all-zero == empty.  32 bytes, 8-byte aligned pairs
  uint32_t cnt, s0;      // {cnt, s0} and {s1, s2} are each updated with ONE 64-bit atomic add
  uint32_t s1, s2;       // coordinate sums (geometry) / colour sums (colour)
  uint32_t pfirst;       // partition + 1 of the first point that reached the cell (0: none yet)
  uint32_t multi;        // 1 once a second, different partition was seen: the reference's doSmooth (:989-995)
  uint32_t lum_off;      // colour: start of the cell's luma list;   geometry: unused
  uint32_t aux;          // colour: scatter cursor, then the mean/median gate flag
};
__device__ __forceinline__ bool cell_do_smooth( const Cell& c ) { return c.cnt != 0 && c.multi != 0; }
__device__ __forceinline__ Cell cell_load( const Cell* p ) {  // read-only path (filters): two 16-byte non-coherent loads
  const uint4 lo = __ldg( reinterpret_cast<const uint4*>( p ) ), hi = __ldg( reinterpret_cast<const uint4*>( p ) + 1 );
  Cell        c;
  c.cnt = lo.x, c.s0 = lo.y, c.s1 = lo.z, c.s2 = lo.w, c.pfirst = hi.x, c.multi = hi.y, c.lum_off = hi.z, c.aux = hi.w;
  return c;
}

// The reference walks a dense int grid (w^3 ints per frame, memset every frame).  Here every frame owns an
// open-addressing hash table keyed by the cell coordinates: 4x4x4 neighbourhoods of cells share 64 consecutive slots,
// so the 2x2x2 cells a boundary point reads sit in one or two cache lines and the whole table of a frame (a few MB)
// stays in L2 while that frame is being processed.  Every occupied cell is accumulated (the reference only
// accumulates cells marked by a boundary point, but its filter never reads any other cell, so the result is the
// same and the marking pass disappears).  Claimed slots are recorded and reset by the cleanup pass.
struct GridArgs {
  int            F;
  int            g;          // cell size
  int            wmax;       // cells per axis
  int            by_bbox;    // geometry: th = g * ceil(maxCoord / g); colour: th = 2^bitdepth
  int            pcmax;      // 2^geometryBitDepth3D
  uint32_t*      keys;       // [F][slots] 0 = empty, else (cx | cy << 10 | cz << 20) + 1
  Cell*          cells;      // [F][slots]
  uint32_t       slots;      // per frame, power of two
  int32_t*       counters;   // [1] overflow flag, [2] luma cursor
  const int64_t* frame_off;
  RbFrameInfo*   finfo;
  short4*        pos;
  ushort4*       col;
  const uint32_t* part;
  const uint32_t* blist;     // indices of the points classified as boundary (type 1) by the reconstruction
  const uint32_t* blist_n;
  uint4*          fcell;     // [F][slots] what the filters read: {float sum0, sum1, sum2, cnt | multi << 30 | gate << 31}
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

__device__ __forceinline__ uint32_t cell_key( int cx, int cy, int cz ) {
  return ( (uint32_t)cx | ( (uint32_t)cy << 10 ) | ( (uint32_t)cz << 20 ) ) + 1u;
}
__device__ __forceinline__ uint32_t cell_home( int cx, int cy, int cz, uint32_t mask ) {
  const uint32_t h = ( (uint32_t)( cx >> 2 ) * 73856093u ) ^ ( (uint32_t)( cy >> 2 ) * 19349663u ) ^ ( (uint32_t)( cz >> 2 ) * 83492791u );
  return ( ( h << 6 ) | (uint32_t)( ( cx & 3 ) | ( ( cy & 3 ) << 2 ) | ( ( cz & 3 ) << 4 ) ) ) & mask;
}
constexpr int MAX_PROBES = 64;

// slot of the cell (global index), or 0xFFFFFFFF when the cell holds no point
__device__ __forceinline__ uint32_t cell_find( const GridArgs& a, int f, int cx, int cy, int cz ) {
  const uint32_t mask = a.slots - 1, key = cell_key( cx, cy, cz );
  uint32_t       s    = cell_home( cx, cy, cz, mask );
  const uint32_t base = (uint32_t)f * a.slots;
  for ( int k = 0; k < MAX_PROBES; k++ ) {
    const uint32_t v = __ldg( a.keys + base + s );  // the keys are final once the accumulate kernel has finished
    if ( v == key ) { return base + s; }
    if ( v == 0 ) { return 0xFFFFFFFFu; }
    s = ( s + 64 ) & mask;
  }
  return 0xFFFFFFFFu;
}
// What the filters need of a cell, in ONE 16-byte record written by k_finalize_cells / k_cell_median_gate: the three
// sums as floats (exact: < 2^24, App. A.3), the count, doSmooth and the mean/median gate.
constexpr uint32_t FC_MULTI = 1u << 30, FC_GATE = 1u << 31, FC_CNT = 0xFFFFFu;

// the 2x2x2 cells a boundary point blends: all eight first probes are issued before any of them is examined, then the
// eight 16-byte records are fetched together — two dependent memory round trips instead of dozens
__device__ __forceinline__ void cell_find8( const GridArgs& a, int f, const int S[3], uint4 fc[8] ) {
  const uint32_t mask = a.slots - 1, base = (uint32_t)f * a.slots;
  uint32_t       s[8], key[8], v[8];
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {  // k = dz*4 + dy*2 + dx, the reference's loop order (:1019-1027)
    const int cx = S[0] + ( k & 1 ), cy = S[1] + ( ( k >> 1 ) & 1 ), cz = S[2] + ( k >> 2 );
    key[k]       = cell_key( cx, cy, cz );
    s[k]         = cell_home( cx, cy, cz, mask );
  }
#pragma unroll
  for ( int k = 0; k < 8; k++ ) { v[k] = __ldg( a.keys + base + s[k] ); }
  uint32_t miss = 0;  // cells whose first probe hit another key: continue their probe sequence (rare)
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {
    if ( v[k] != key[k] && v[k] != 0 ) { miss |= 1u << k; }
  }
  if ( miss ) {
#pragma unroll
    for ( int k = 0; k < 8; k++ ) {
      if ( !( miss >> k & 1u ) ) { continue; }
      uint32_t t = s[k];
      v[k]       = 0;
      for ( int j = 1; j < MAX_PROBES; j++ ) {
        t                = ( t + 64 ) & mask;
        const uint32_t w = __ldg( a.keys + base + t );
        if ( w == key[k] ) {
          s[k] = t;
          v[k] = w;
          break;
        }
        if ( w == 0 ) { break; }
      }
    }
  }
#pragma unroll
  for ( int k = 0; k < 8; k++ ) { fc[k] = __ldg( a.fcell + base + s[k] ); }  // always a valid address
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {
    if ( v[k] != key[k] ) { fc[k] = make_uint4( 0, 0, 0, 0 ); }  // the cell holds no point
  }
}

// find or claim; 0xFFFFFFFF on table overflow (flagged)
__device__ __forceinline__ uint32_t cell_claim( const GridArgs& a, int f, int cx, int cy, int cz ) {
  const uint32_t mask = a.slots - 1, key = cell_key( cx, cy, cz );
  uint32_t       s    = cell_home( cx, cy, cz, mask );
  const uint32_t base = (uint32_t)f * a.slots;
  for ( int k = 0; k < MAX_PROBES; k++ ) {
    uint32_t v = a.keys[base + s];
    if ( v == 0 ) {
      v = atomicCAS( &a.keys[base + s], 0u, key );
      if ( v == 0 ) { return base + s; }  // claimed (no global list: the per-cell passes walk the slot array)
    }
    if ( v == key ) { return base + s; }
    s = ( s + 64 ) & mask;
  }
  a.counters[1] = 1;
  return 0xFFFFFFFFu;
}

// warp-aggregated accumulation: lanes that hit the same cell are reduced with one hardware warp reduction per field
// and their leader issues the atomics (points arrive in patch-block order, so a warp touches one to four cells)
__device__ __forceinline__ void cell_accumulate( const GridArgs& a, bool valid, uint32_t slot, uint32_t v0, uint32_t v1,
                                                 uint32_t v2, uint32_t pp ) {
  const uint32_t act = __ballot_sync( 0xFFFFFFFFu, valid );
  if ( !valid ) { return; }
  const uint32_t peers = __match_any_sync( act, slot );
  const uint32_t n     = __popc( peers );
  const uint32_t t0 = __reduce_add_sync( peers, v0 ), t1 = __reduce_add_sync( peers, v1 ), t2 = __reduce_add_sync( peers, v2 );
  const uint32_t mx = __reduce_max_sync( peers, pp ), mn = __reduce_min_sync( peers, pp );
  if ( ( threadIdx.x & 31 ) == __ffs( peers ) - 1 ) {
    Cell* c = a.cells + slot;
    atomicAdd( (unsigned long long*)&c->cnt, (unsigned long long)n | ( (unsigned long long)t0 << 32 ) );
    atomicAdd( (unsigned long long*)&c->s1, (unsigned long long)t1 | ( (unsigned long long)t2 << 32 ) );
    if ( mx != mn ) {
      c->multi = 1;  // two partitions inside this very group
      atomicCAS( &c->pfirst, 0u, mx );
    } else {
      const uint32_t old = atomicCAS( &c->pfirst, 0u, mx );
      if ( old != 0 && old != mx ) { c->multi = 1; }
    }
  }
}

constexpr int ACC_RUN = 8;  // consecutive points per thread

// Points arrive in emission order (patch -> 16x16 block -> pixel row -> layer), so the 8 consecutive points of a
// thread — 4 neighbouring pixels x 2 layers — fall into one or two cells: the thread merges them in registers and
// issues one hash probe and three atomics per distinct cell instead of per point.
struct RunAcc {
  uint32_t slot, n, t0, t1, t2, mx, mn;
};
__device__ __forceinline__ void run_flush( const GridArgs& a, RunAcc& r ) {
  if ( r.n == 0 || r.slot == 0xFFFFFFFFu ) {
    r.n = 0;
    return;
  }
  Cell* c = a.cells + r.slot;
  atomicAdd( (unsigned long long*)&c->cnt, (unsigned long long)r.n | ( (unsigned long long)r.t0 << 32 ) );
  atomicAdd( (unsigned long long*)&c->s1, (unsigned long long)r.t1 | ( (unsigned long long)r.t2 << 32 ) );
  if ( r.mx != r.mn ) {
    c->multi = 1;  // two partitions inside this very run
    atomicCAS( &c->pfirst, 0u, r.mx );
  } else {
    const uint32_t old = atomicCAS( &c->pfirst, 0u, r.mx );
    if ( old != 0 && old != r.mx ) { c->multi = 1; }
  }
  r.n = 0;
}

template <bool COLOUR>
__global__ void __launch_bounds__( 256 ) k_accumulate( const GridArgs a, int64_t n ) {
  const int64_t i0 = ( (int64_t)blockIdx.x * 256 + threadIdx.x ) * ACC_RUN;
  if ( i0 >= n ) { return; }
  short4   p[ACC_RUN];
  ushort4  cv[ACC_RUN];
  uint32_t pp[ACC_RUN];
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
  int       f      = frame_of( a.frame_off, a.F, i0 );
  int64_t   fend   = a.frame_off[f + 1];
  const int disth  = max( a.g / 2, 1 );
  int       th     = COLOUR ? 0 : grid_th( a, f );
  int       lx = -1, ly = -1, lz = -1, lf = -1;
  RunAcc    r{0xFFFFFFFFu, 0, 0, 0, 0, 0, 0xFFFFFFFFu};
#pragma unroll
  for ( int k = 0; k < ACC_RUN; k++ ) {
    const int64_t i = i0 + k;
    if ( i >= n ) { break; }
    while ( i >= fend ) {  // the run crosses into the next frame (empty frames are skipped)
      f++;
      fend = a.frame_off[f + 1];
      if ( !COLOUR ) { th = grid_th( a, f ); }
    }
    const short4 q = p[k];
    bool         in;
    if ( COLOUR ) {  // no margin test, :208-224 with the :212 guard
      in = q.x >= 0 && q.y >= 0 && q.z >= 0 && q.x / a.g < a.wmax && q.y / a.g < a.wmax && q.z / a.g < a.wmax;
    } else {  // :120-134
      in = inside( q.x, q.y, q.z, disth, th );
    }
    if ( !in ) { continue; }
    const int cx = q.x / a.g, cy = q.y / a.g, cz = q.z / a.g;
    if ( cx != lx || cy != ly || cz != lz || f != lf ) {
      run_flush( a, r );
      r.slot = cell_claim( a, f, cx, cy, cz );
      r.t0 = r.t1 = r.t2 = r.mx = 0;
      r.mn               = 0xFFFFFFFFu;
      lx = cx, ly = cy, lz = cz, lf = f;
    }
    r.n++;
    if ( COLOUR ) {
      r.t0 += cv[k].x, r.t1 += cv[k].y, r.t2 += cv[k].z;
    } else {
      r.t0 += (uint32_t)q.x, r.t1 += (uint32_t)q.y, r.t2 += (uint32_t)q.z;
    }
    r.mx = max( r.mx, pp[k] );
    r.mn = min( r.mn, pp[k] );
  }
  run_flush( a, r );
}

// exactness guard of App. A.3 + luma list allocation (colour), over the claimed slots
__global__ void k_finalize_cells( const GridArgs a, uint32_t nSlots, int colour ) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  const bool     live = slot < nSlots && a.keys[slot] != 0;
  if ( !__any_sync( 0xFFFFFFFFu, live ) ) { return; }
  uint32_t cnt = 0, s0 = 0, s1 = 0, s2 = 0;
  if ( live ) {
    const Cell* c = a.cells + slot;
    cnt = c->cnt, s0 = c->s0, s1 = c->s1, s2 = c->s2;
    if ( s0 >= ( 1u << 24 ) || s1 >= ( 1u << 24 ) || s2 >= ( 1u << 24 ) || cnt > 65535u ) {
      a.finfo[slot / a.slots].sum_overflow = 1;
    }
    const Cell* cc = a.cells + slot;
    a.fcell[slot]  = make_uint4( __float_as_uint( (float)s0 ), __float_as_uint( (float)s1 ), __float_as_uint( (float)s2 ),
                                 min( cnt, FC_CNT ) | ( cc->multi ? FC_MULTI : 0u ) );
  }
  if ( colour ) {  // luma list offsets: one atomic per warp, shuffle prefix inside
    const uint32_t want = ( live && cnt > 1 ) ? cnt : 0u;
    const int      lane = threadIdx.x & 31;
    uint32_t       incl = want;
#pragma unroll
    for ( int d = 1; d < 32; d <<= 1 ) {
      const uint32_t t = __shfl_up_sync( 0xFFFFFFFFu, incl, d );
      if ( lane >= d ) { incl += t; }
    }
    uint32_t base = 0;
    if ( lane == 31 && incl ) { base = (uint32_t)atomicAdd( &a.counters[2], (int)incl ); }
    base = __shfl_sync( 0xFFFFFFFFu, base, 31 );
    if ( live ) {
      a.cells[slot].lum_off = base + incl - want;
      a.cells[slot].aux     = 0;
    }
  }
}

// ---- geometry filter: smoothPointCloudGrid + gridFiltering (:1000-1104) ----
__global__ void __launch_bounds__( 128, 12 ) k_filter_geo( const GridArgs a, double threshold ) {
  const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
  if ( li >= *a.blist_n ) { return; }
  const int64_t i = a.blist[li];
  const short4  p = a.pos[i];
  if ( p.w != 1 ) { return; }  // :1087
  const int f     = frame_of( a.frame_off, a.F, i );
  const int g     = a.g, hg = g / 2;
  const int disth = max( hg, 1 ), th = grid_th( a, f );
  if ( !inside( p.x, p.y, p.z, disth, th ) ) { return; }  // :1078-1081
  const int      P[3] = {p.x, p.y, p.z};
  int            S[3];
  for ( int k = 0; k < 3; k++ ) { S[k] = P[k] / g + ( ( P[k] - ( P[k] / g ) * g < hg ) ? -1 : 0 ); }  // :1014-1017
  uint4    fc[8];
  bool     other = false;
  uint32_t cnt[8];
  cell_find8( a, f, S, fc );
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {
    cnt[k] = fc[k].w & FC_CNT;
    if ( cnt[k] != 0 && ( fc[k].w & FC_MULTI ) ) { other = true; }  // doSmooth && count (:1024)
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
      const float fcn = (float)cnt[k];
      v[0]            = (double)__fdiv_rn( __uint_as_float( fc[k].x ), fcn );
      v[1]            = (double)__fdiv_rn( __uint_as_float( fc[k].y ), fcn );
      v[2]            = (double)__fdiv_rn( __uint_as_float( fc[k].z ), fcn );
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
__global__ void __launch_bounds__( 256 ) k_scatter_lum( const GridArgs a, int64_t n, uint16_t* __restrict__ lum ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  const short4 p = a.pos[i];
  const int    f = frame_of( a.frame_off, a.F, i );
  if ( p.x < 0 || p.y < 0 || p.z < 0 || p.x / a.g >= a.wmax || p.y / a.g >= a.wmax || p.z / a.g >= a.wmax ) { return; }
  const uint32_t slot = cell_find( a, f, p.x / a.g, p.y / a.g, p.z / a.g );
  if ( slot == 0xFFFFFFFFu ) { return; }
  Cell* c = a.cells + slot;
  if ( c->cnt < 2 ) { return; }  // the median is only consulted for count > 1 (:1228, :1239)
  const uint32_t k       = atomicAdd( &c->aux, 1u );
  lum[c->lum_off + k]    = a.col[i].x;
}

__device__ __forceinline__ void k_median_one( const GridArgs& a, Cell* c, int lane, const uint16_t* __restrict__ lum, double mmThresh ) {
  const int n = (int)c->cnt;
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
    const bool gate = (double)( diff < 0 ? -diff : diff ) > mmThresh;
    c->aux          = gate ? 1u : 0u;
    if ( gate ) { a.fcell[c - a.cells].w |= FC_GATE; }
  }
}

// ---- colour: per-cell mean/median gate (:1228-1236, :1239-1243); one warp per cell, rank selection ----
__global__ void k_cell_median_gate( const GridArgs a, uint32_t nSlots, const uint16_t* __restrict__ lum, double mmThresh ) {
  const uint32_t s0   = ( blockIdx.x * blockDim.x + threadIdx.x ) & ~31u;  // this warp owns slots s0 .. s0+31
  const int      lane = threadIdx.x & 31;
  const bool     want = s0 + lane < nSlots && a.keys[s0 + lane] != 0 && a.cells[s0 + lane].cnt > 1;
  uint32_t       todo = __ballot_sync( 0xFFFFFFFFu, want );
  if ( s0 + lane < nSlots && a.keys[s0 + lane] != 0 && !want ) { a.cells[s0 + lane].aux = 0; }
  for ( ; todo; todo &= todo - 1 ) {
    k_median_one( a, a.cells + s0 + ( __ffs( todo ) - 1 ), lane, lum, mmThresh );
  }
}
// ---- colour filter: smoothPointCloudColorLC + gridFilteringColor (:1182-1306) ----
__global__ void __launch_bounds__( 128, 10 ) k_filter_col( const GridArgs a, double thrSmoothing, double yThresh ) {
  const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
  if ( li >= *a.blist_n ) { return; }
  const int64_t i = a.blist[li];
  const short4  p = a.pos[i];
  if ( p.w != 1 ) { return; }  // :1288
  const int g = a.g, hg = g / 2, disth = max( hg, 1 );
  if ( !inside( p.x, p.y, p.z, disth, a.pcmax ) ) { return; }  // :1280-1283
  const int      f = frame_of( a.frame_off, a.F, i );
  const int      P[3] = {p.x, p.y, p.z};
  int            S[3];
  for ( int k = 0; k < 3; k++ ) { S[k] = P[k] / g + ( ( ( P[k] % g ) < hg ) ? -1 : 0 ); }  // :1197-1199
  const ushort4 cv = a.col[i];  // issued with the probes: independent of them
  uint4         fc[8];
  bool          other = false;
  cell_find8( a, f, S, fc );
#pragma unroll
  for ( int k = 0; k < 8; k++ ) {
    if ( ( fc[k].w & FC_CNT ) != 0 && ( fc[k].w & FC_MULTI ) ) { other = true; }  // :1204
  }
  if ( !other ) { return; }  // :1210
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
    const uint32_t cn   = fc[k].w & FC_CNT;
    const bool     gate = cn > 1 && ( fc[k].w & FC_GATE );
    double*        d    = c3[k];
    if ( cn > 0 ) {
      const double dn = (double)cn;  // :1225: float accumulator read back as double, divided by the count in double
      d[0]            = (double)__uint_as_float( fc[k].x ) / dn;
      d[1]            = (double)__uint_as_float( fc[k].y ) / dn;
      d[2]            = (double)__uint_as_float( fc[k].z ) / dn;
      if ( k == 0 ) {
        if ( gate ) {  // :1228-1235: result = own colour
          keep_own = true;
          break;
        }
      } else {
        const int dy0 = (int)( Y0 - d[0] );  // abs() truncates (App. A.9), :1238
        bool      own = (double)( dy0 < 0 ? -dy0 : dy0 ) > yThresh;
        if ( gate ) { own = true; }  // :1239-1243
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

// ---- cleanup: reset exactly the claimed slots ----
__global__ void k_cleanup_cells( const GridArgs a, uint32_t nSlots ) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if ( slot >= nSlots || a.keys[slot] == 0 ) { return; }
  a.keys[slot] = 0;
  uint4* c     = reinterpret_cast<uint4*>( a.cells + slot );
  c[0] = c[1] = make_uint4( 0, 0, 0, 0 );
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

// table geometry + (re)allocation; the tables are all-zero between calls (the cleanup pass resets what was claimed)
int setup_grid( rb200_ctx* c, GridArgs& a, RbBuf& keys, RbBuf& cells, RbBuf& counters, RbBuf& fcell, int g, int wmax,
                uint32_t* nSlotsOut ) {
  if ( wmax > 1024 ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "smoothing grid wider than 1024 cells per axis" ); }
  int64_t maxFrame = 1;
  for ( int f = 0; f < c->F; f++ ) { maxFrame = std::max<int64_t>( maxFrame, c->h_frame_off[f + 1] - c->h_frame_off[f] ); }
  int64_t want = std::min<int64_t>( 2 * maxFrame, std::max<int64_t>( 4096, 4 * maxFrame / ( (int64_t)g * g ) ) );
  uint32_t slots = 4096;
  while ( (int64_t)slots < want ) { slots <<= 1; }
  const size_t nSlots = (size_t)c->F * slots;
  if ( nSlots >= ( 1ull << 32 ) ) { return rb_fail( c, RB200_ERR_UNSUPPORTED, "smoothing hash table too large" ); }
  if ( keys.cap < nSlots * 4 || cells.cap < nSlots * sizeof( Cell ) ) {
    RB_CUDA( keys.ensure( nSlots * 4 ) );
    RB_CUDA( cells.ensure( nSlots * sizeof( Cell ) ) );
    RB_CUDA( cudaMemsetAsync( keys.p, 0, keys.cap, c->stream ) );
    RB_CUDA( cudaMemsetAsync( cells.p, 0, cells.cap, c->stream ) );
  }
  RB_CUDA( fcell.ensure( nSlots * 16 ) );
  RB_CUDA( counters.ensure( 64 ) );
  RB_CUDA( cudaMemsetAsync( counters.p, 0, 64, c->stream ) );
  a.keys     = keys.as<uint32_t>();
  a.cells    = cells.as<Cell>();
  a.slots    = slots;
  a.fcell    = fcell.as<uint4>();
  a.counters = counters.as<int32_t>();
  *nSlotsOut = (uint32_t)nSlots;
  return RB200_OK;
}

// the table-overflow flag, read after the stage has been enqueued (one small read-back, also the stage's error check)
int check_overflow( rb200_ctx* c, const GridArgs& a, RbBuf& keys, RbBuf& cells, const char* what ) {
  int32_t* h = (int32_t*)rb_pinned( c, 64 );
  if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  RB_CUDA( cudaMemcpyAsync( h, a.counters, 16, cudaMemcpyDeviceToHost, c->stream ) );
  RB_CUDA( cudaStreamSynchronize( c->stream ) );
  c->stats.d2h_bytes += 16;
  if ( h[1] ) {
    RB_CUDA( cudaMemsetAsync( keys.p, 0, keys.cap, c->stream ) );
    RB_CUDA( cudaMemsetAsync( cells.p, 0, cells.cap, c->stream ) );
    return rb_fail( c, RB200_ERR_NOMEM, "%s: cell table overflow", what );
  }
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
  uint32_t nSlots = 0;
  int      r      = setup_grid( c, a, c->d_geo_grid, c->d_geo_cells, c->d_scratch[2], c->d_scratch[5], g, wmax, &nSlots );
  if ( r ) { return r; }
  RB_LAUNCH( "geo_accumulate", k_accumulate<false>, rb_div_up( n, 256 * ACC_RUN ), 256, 0, a, n );
  RB_LAUNCH( "geo_finalize", k_finalize_cells, rb_div_up( nSlots, 256 ), 256, 0, a, nSlots, 0 );
  if ( c->blist_cap > 0 ) { RB_LAUNCH( "geo_filter", k_filter_geo, rb_div_up( c->blist_cap, 128 ), 128, 0, a, P.threshold_smoothing ); }
  RB_LAUNCH( "geo_cleanup", k_cleanup_cells, rb_div_up( nSlots, 256 ), 256, 0, a, nSlots );
  return check_overflow( c, a, c->d_geo_grid, c->d_geo_cells, "geometry smoothing" );
}

int rb_smooth_color_impl( rb200_ctx* c ) {
  const rb200_params& P = c->P;
  const int64_t       n = c->h_frame_off[c->F];
  if ( n == 0 || P.attribute_count == 0 ) { return RB200_OK; }
  const int g     = P.occupancy_precision;  // the colour grid uses occupancyPrecision, not cgridSize (:152)
  const int pcmax = 1 << P.geometry_bitdepth_3d;
  const int wmax  = pcmax / g;  // :154
  if ( wmax < 2 ) { return rb_fail( c, RB200_ERR_INVALID, "colour grid degenerate" ); }
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
  uint32_t nSlots = 0;
  int      r      = setup_grid( c, a, c->d_col_grid, c->d_col_cells, c->d_scratch[3], c->d_scratch[6], g, wmax, &nSlots );
  if ( r ) { return r; }
  RB_CUDA( c->d_col_lum.ensure( (size_t)n * 2 + 64 ) );
  RB_LAUNCH( "col_accumulate", k_accumulate<true>, rb_div_up( n, 256 * ACC_RUN ), 256, 0, a, n );
  RB_LAUNCH( "col_finalize", k_finalize_cells, rb_div_up( nSlots, 256 ), 256, 0, a, nSlots, 1 );
  RB_LAUNCH( "col_scatter_lum", k_scatter_lum, rb_div_up( n, 256 ), 256, 0, a, n, c->d_col_lum.as<uint16_t>() );
  RB_LAUNCH( "col_median_gate", k_cell_median_gate, rb_div_up( nSlots, 256 ), 256, 0, a, nSlots, c->d_col_lum.as<uint16_t>(),
             P.threshold_color_variation * 256.0 );
  if ( c->blist_cap > 0 ) {
    RB_LAUNCH( "col_filter", k_filter_col, rb_div_up( c->blist_cap, 128 ), 128, 0, a, P.threshold_color_smoothing,
               P.threshold_color_difference * 256.0 );
  }
  RB_LAUNCH( "col_cleanup", k_cleanup_cells, rb_div_up( nSlots, 256 ), 256, 0, a, nSlots );
  return check_overflow( c, a, c->d_col_grid, c->d_col_cells, "colour smoothing" );
}

int rb_convert_rgb8_impl( rb200_ctx* c ) {
  const int64_t n = c->h_frame_off[c->F];
  if ( n == 0 ) { return RB200_OK; }
  RB_LAUNCH( "to_rgb8", k_to_rgb8, rb_div_up( n, 256 ), 256, 0, c->d_col.as<ushort4>(), c->d_rgb.as<uchar4>(), n,
             c->P.attribute_rgb444, c->P.attribute_count );
  return RB200_OK;
}
