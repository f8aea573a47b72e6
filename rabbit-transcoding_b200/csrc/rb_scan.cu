// rb_scan.cu — exclusive prefix sum over uint32 (in place allowed) used by the metrics CSR build and the kd-tree
// level passes: per-tile sums -> scan of the sums (one CTA) -> per-tile scan + offset.  16 items per thread, uint4 I/O.
#include "rb_common.cuh"

namespace {
constexpr int TPB = 256;
// ------------------------------------------------------------------------------------------------
// generic exclusive scan over uint32 (in place allowed): block sums -> scan of sums -> apply
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE  = TPB * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan( uint32_t v, uint32_t* smem, uint32_t& total ) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t  incl = v;
#pragma unroll
  for ( int d = 1; d < 32; d <<= 1 ) {
    const uint32_t t = __shfl_up_sync( 0xFFFFFFFFu, incl, d );
    if ( lane >= d ) { incl += t; }
  }
  if ( lane == 31 ) { smem[w] = incl; }
  __syncthreads();
  if ( w == 0 ) {
    uint32_t x = lane < ( TPB / 32 ) ? smem[lane] : 0u, y = x;
#pragma unroll
    for ( int d = 1; d < 32; d <<= 1 ) {
      const uint32_t t = __shfl_up_sync( 0xFFFFFFFFu, y, d );
      if ( lane >= d ) { y += t; }
    }
    if ( lane < ( TPB / 32 ) ) { smem[lane] = y - x; }
    if ( lane == ( TPB / 32 ) - 1 ) { smem[32] = y; }
  }
  __syncthreads();
  total = smem[32];
  return smem[w] + incl - v;
}

// `in` / `out` are the 16-byte aligned addresses at or below the caller's, `skip` (0..3) the elements in front of the
// caller's first one: they read as 0 and are never written (n counts them)
__global__ void __launch_bounds__( TPB ) k_scan_sums( const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ sums, int skip ) {
  __shared__ uint32_t sm[33];
  const int64_t       base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t            s    = 0;
  if ( base + SCAN_ITEMS <= n && base >= skip ) {
    const uint4* p = reinterpret_cast<const uint4*>( in + base );
#pragma unroll
    for ( int k = 0; k < SCAN_ITEMS / 4; k++ ) {
      const uint4 v = p[k];
      s += v.x + v.y + v.z + v.w;
    }
  } else {
    for ( int k = 0; k < SCAN_ITEMS; k++ ) {
      if ( base + k < n && base + k >= skip ) { s += in[base + k]; }
    }
  }
  uint32_t total;
  block_exclusive_scan( s, sm, total );
  if ( threadIdx.x == 0 ) { sums[blockIdx.x] = total; }
}

__global__ void __launch_bounds__( 1024 ) k_scan_top( uint32_t* __restrict__ sums, int64_t nb, uint32_t* __restrict__ total_out ) {
  __shared__ uint32_t warpSum[32];
  __shared__ uint32_t carry;
  if ( threadIdx.x == 0 ) { carry = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for ( int64_t b = 0; b < nb; b += 1024 ) {
    const int64_t  i = b + threadIdx.x;
    const uint32_t v = i < nb ? sums[i] : 0u;
    uint32_t       incl = v;
#pragma unroll
    for ( int d = 1; d < 32; d <<= 1 ) {
      const uint32_t t = __shfl_up_sync( 0xFFFFFFFFu, incl, d );
      if ( lane >= d ) { incl += t; }
    }
    if ( lane == 31 ) { warpSum[w] = incl; }
    __syncthreads();
    if ( w == 0 ) {
      uint32_t x = warpSum[lane], y = x;
#pragma unroll
      for ( int d = 1; d < 32; d <<= 1 ) {
        const uint32_t t = __shfl_up_sync( 0xFFFFFFFFu, y, d );
        if ( lane >= d ) { y += t; }
      }
      warpSum[lane] = y - x;
    }
    __syncthreads();
    const uint32_t excl = carry + warpSum[w] + incl - v;
    if ( i < nb ) { sums[i] = excl; }
    __syncthreads();
    if ( threadIdx.x == 1023 ) { carry = excl + v; }
    __syncthreads();
  }
  if ( threadIdx.x == 0 && total_out ) { *total_out = carry; }
}

__global__ void __launch_bounds__( TPB ) k_scan_apply( const uint32_t* __restrict__ in, int64_t n, const uint32_t* __restrict__ sums,
                                                       uint32_t* __restrict__ out, int skip ) {
  __shared__ uint32_t sm[33];
  const int64_t       base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t            v[SCAN_ITEMS];
  uint32_t            s = 0;
  const bool          whole = base + SCAN_ITEMS <= n && base >= skip;
  if ( whole ) {
    const uint4* p = reinterpret_cast<const uint4*>( in + base );
#pragma unroll
    for ( int k = 0; k < SCAN_ITEMS / 4; k++ ) {
      const uint4 q = p[k];
      v[4 * k] = q.x, v[4 * k + 1] = q.y, v[4 * k + 2] = q.z, v[4 * k + 3] = q.w;
    }
  } else {
#pragma unroll
    for ( int k = 0; k < SCAN_ITEMS; k++ ) { v[k] = ( base + k < n && base + k >= skip ) ? in[base + k] : 0u; }
  }
#pragma unroll
  for ( int k = 0; k < SCAN_ITEMS; k++ ) { s += v[k]; }
  uint32_t total;
  uint32_t run = block_exclusive_scan( s, sm, total ) + sums[blockIdx.x];
  if ( whole ) {
    uint4* p = reinterpret_cast<uint4*>( out + base );
#pragma unroll
    for ( int k = 0; k < SCAN_ITEMS / 4; k++ ) {
      uint4 q;
      q.x = run, run += v[4 * k];
      q.y = run, run += v[4 * k + 1];
      q.z = run, run += v[4 * k + 2];
      q.w = run, run += v[4 * k + 3];
      p[k] = q;
    }
  } else {
#pragma unroll
    for ( int k = 0; k < SCAN_ITEMS; k++ ) {
      if ( base + k < n && base + k >= skip ) { out[base + k] = run; }
      run += v[k];
    }
  }
}

}  // namespace

size_t rb_scan_scratch_bytes( int64_t n ) { return (size_t)( n / SCAN_TILE + 2 ) * 4; }

int rb_scan_u32( rb200_ctx* c, const uint32_t* in, uint32_t* out, int64_t n, uint32_t* sums ) {
  uint32_t* total_out = nullptr;
  if ( n <= 0 ) { return RB200_OK; }
  // a range that does not start on a 16-byte boundary (the rebuilt tail of a cached metrics batch) is scanned from the
  // boundary below it, the elements in front masked out
  const int skip = (int)( ( (uintptr_t)in >> 2 ) & 3 );
  if ( skip != (int)( ( (uintptr_t)out >> 2 ) & 3 ) ) { return rb_fail( c, RB200_ERR_INVALID, "scan: in and out must be equally aligned" ); }
  in -= skip, out -= skip, n += skip;
  const int64_t nb = ( n + SCAN_TILE - 1 ) / SCAN_TILE;
  RB_LAUNCH( "scan_sums", k_scan_sums, (unsigned)nb, TPB, 0, in, n, sums, skip );
  RB_LAUNCH( "scan_top", k_scan_top, 1, 1024, 0, sums, nb, total_out );
  RB_LAUNCH( "scan_apply", k_scan_apply, (unsigned)nb, TPB, 0, in, n, sums, out, skip );
  return RB200_OK;
}

