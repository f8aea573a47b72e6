// rb_ingest.cu — decoder-native planes -> the planes the reconstruction reads (sm_100a).
//
// Restates what PCCVideoDecoder does to the attribute video after the codec returns:
//   PCCInternalColorConverter<T>::convertYUV420ToYUV444   PccLibColorConverter/source/PCCInternalColorConverter.cpp:456-486
//     YUVtoFloatYUV (:596-611), upsampling (:669-695) with upsamplingVertical0/1 + upsamplingHorizontal0/1
//     (include/PCCInternalColorConverter.h:187-249), floatYUVToYUV (:582-594), filters g_filter420to444 (:297-337)
//   as invoked by PccLibDecoder/source/PCCVideoDecoder.cpp:125-146, :365 ("YUV420ToYUV444_<bits>_<filter>").
// so that the 8/10-bit 4:2:0 planes a video decoder leaves behind go straight to HBM (a quarter of the bytes of the
// 16-bit 4:4:4 frames) and the conversion — a dense separable single-precision filter — runs on the GPU.
//
// Arithmetic: every sample goes through the reference's float / double steps (compiled with -fmad=false), the filter
// sums are accumulated in single precision in tap order.  The per-sample scalings are pure functions of an 8/10-bit
// value and are tabulated per CTA.  One CTA converts one 64x32 tile of one chroma plane: the chroma tile and its
// 6-sample halo are staged as floats in shared memory, filtered vertically into shared memory, then horizontally
// into 16-bit output rows (16-byte stores); there is no intermediate plane in HBM.
#include "rb_common.cuh"

namespace {

constexpr int MAX_TAPS = 12;
struct UpFilter {
  float h0[MAX_TAPS], v0[MAX_TAPS], h1[MAX_TAPS], v1[MAX_TAPS];
  int   nh0, nv0, nh1, nv1;
};
#define TAPS( ... ) { __VA_ARGS__ }
// g_filter420to444 (PCCInternalColorConverter.cpp:297-337), field order horizontal0, vertical0, horizontal1, vertical1
// (include/PCCInternalColorConverter.h:50-55); every filter has shift 8, the offsets are not used by upsampling
__constant__ UpFilter c_up[8] = {
    {TAPS( 0, 256 ), TAPS( -8, 64, 216, -16 ), TAPS( -16, 144, 144, -16 ), TAPS( -16, 216, 64, -8 ), 2, 4, 4, 4},
    {TAPS( 0, 256 ), TAPS( 0, -16, 56, 240, -32, 8 ), TAPS( -16, 144, 144, -16 ), TAPS( 8, -32, 240, 56, -16, 0 ), 2, 6, 4, 6},
    {TAPS( 0, 256 ), TAPS( -6, 58, 222, -18 ), TAPS( -16, 144, 144, -16 ), TAPS( -18, 222, 58, -6 ), 2, 4, 4, 4},
    {TAPS( 0, 256 ), TAPS( 2, -18, 70, 228, -34, 8 ), TAPS( 6, -34, 156, 156, -34, 6 ), TAPS( 8, -34, 228, 70, -18, 2 ), 2, 6, 6, 6},
    {TAPS( 0, 256 ), TAPS( -1, 8, -23, 72, 229, -39, 14, -4 ), TAPS( -3, 15, -43, 159, 159, -43, 15, -3 ),
     TAPS( -4, 14, -39, 229, 72, -23, 8, -1 ), 2, 8, 8, 8},
    {TAPS( 0, 256 ), TAPS( 3, -16, 67, 227, -32, 7 ), TAPS( 21, -52, 159, 159, -52, 21 ), TAPS( 7, -32, 227, 67, -16, 3 ), 2, 6, 6, 6},
    {TAPS( 0, 256 ), TAPS( 1, -5, 12, -27, 74, 230, -41, 18, -8, 2 ), TAPS( 2, -8, 21, -47, 160, 160, -47, 21, -8, 2 ),
     TAPS( 2, -8, 18, -41, 230, 74, -27, 12, -5, 1 ), 2, 10, 10, 10},
    {TAPS( 0, 256 ), TAPS( 0, 3, -7, 14, -29, 75, 230, -43, 20, -10, 5, -2 ), TAPS( -1, 5, -12, 24, -49, 161, 161, -49, 24, -12, 5, -1 ),
     TAPS( -2, 5, -10, 20, -43, 230, 75, -29, 14, -7, 3, 0 ), 2, 12, 12, 12},
};

// PCCImage::set (PccLibCommon/include/PCCImage.h:97-138): the decoder's sample -> the stored sample
__device__ __forceinline__ int image_set( int s, int shift ) {
  if ( shift <= 0 ) { return s; }
  const int v = ( s + ( 1 << ( shift - 1 ) ) ) >> shift;  // (T)( ( src + rounding ) >> shiftbits )
  return min( max( v & 0xFFFF, 0 ), ( 1 << ( 10 - shift ) ) - 1 );
}

// YUVtoFloatYUV (:596-611): clamp( (float)( weight * (double)( sample - offset ) ), min, max )
__device__ __forceinline__ float sample_to_float( int s, bool chroma, int nbyte ) {
  const int    offset = chroma ? ( nbyte == 1 ? 128 : 512 ) : 0;
  const double weight = 1.0 / ( nbyte == 1 ? 255. : 1023. );
  const float  f      = (float)__dmul_rn( weight, (double)( s - offset ) );
  return fminf( fmaxf( f, chroma ? -0.5f : 0.f ), chroma ? 0.5f : 1.f );
}
// floatYUVToYUV (:582-594) with nbyte = 2: (T) fClip( round( (float)( 65535. * (double)f + offset ) ), 0, 65535 )
// For |f| >= 2^-12 the double expression is EXACT (65535 f is a 40-bit integer times ulp( f ) >= 2^-35, the offset keeps the
// sum inside 53 bits), so its conversion to float is the single rounding of 65535 f + offset: one float fused
// multiply-add.  Below 2^-12 the double sum rounds first; all 1.93e9 such floats were enumerated on the host: the two
// float results differ for 4 of them and the rounded integer for none.
__device__ __forceinline__ uint16_t float_to_u16( float f, bool chroma ) {
  const float x = __fmaf_rn( 65535.0f, f, chroma ? 32768.0f : 0.0f );
  return (uint16_t)fminf( fmaxf( roundf( x ), 0.f ), 65535.f );
}

// ---- luma: a pure function of the sample, tabulated per CTA ----
template <typename T>
__global__ void __launch_bounds__( 256 ) k_luma_to_16( const T* __restrict__ src, uint16_t* __restrict__ dst, int W, int H,
                                                        int nbyte, int shift, size_t srcFrameStride, size_t dstFrameStride ) {
  __shared__ uint16_t lut[1024];
  const int           levels = 1024;  // decoder samples below 1024 go through the table, anything else is computed
  for ( int s = threadIdx.x; s < levels; s += blockDim.x ) { lut[s] = float_to_u16( sample_to_float( image_set( s, shift ), false, nbyte ), false ); }
  __syncthreads();
  const T*      in  = src + (size_t)blockIdx.y * srcFrameStride;
  uint16_t*     out = dst + (size_t)blockIdx.y * dstFrameStride;
  const int64_t n   = (int64_t)W * H;
  for ( int64_t i = ( (int64_t)blockIdx.x * blockDim.x + threadIdx.x ) * 8; i < n; i += (int64_t)gridDim.x * blockDim.x * 8 ) {
    if ( i + 8 <= n ) {  // W is a multiple of 16 and the planes are 16-byte aligned
      uint16_t o[8];
#pragma unroll
      for ( int k = 0; k < 8; k++ ) {
        const int s = (int)in[i + k];
        o[k]        = s < levels ? lut[s] : float_to_u16( sample_to_float( image_set( s, shift ), false, nbyte ), false );
      }
      *reinterpret_cast<uint4*>( out + i ) = *reinterpret_cast<const uint4*>( o );
    } else {
      for ( int64_t j = i; j < n; j++ ) { out[j] = float_to_u16( sample_to_float( image_set( (int)in[j], shift ), false, nbyte ), false ); }
    }
  }
}

// ---- chroma: 64 x 32 output tile per CTA ----
constexpr int TW = 64, TH = 32, HALO = 6;
constexpr int IW = TW / 2 + 2 * HALO, IH = TH / 2 + 2 * HALO;  // 44 x 28 staged chroma samples

// NV / NH: taps of the two vertical filters / of horizontal1 (4 .. 12; horizontal0 is {0, 256} in every entry of the
// table): the tap loops unroll and the coefficients sit in registers
template <typename T, int NV, int NH>
__global__ void __launch_bounds__( 256 ) k_chroma_420_to_444( const T* __restrict__ src, uint16_t* __restrict__ dst, int W, int H,
                                                               int nbyte, int shift, int filter, size_t srcFrameStride,
                                                               size_t dstFrameStride ) {
  constexpr int    levels = sizeof( T ) == 1 ? 256 : 1024;  // samples below go through the table, others are computed
  __shared__ float lut[levels];
  __shared__ float in[IH][IW];
  __shared__ float tmp[TH][IW + 1];
  const int        cw = W / 2, ch = H / 2;
  const int        frame = blockIdx.z >> 1, plane = 1 + ( blockIdx.z & 1 );
  // frame layout of the source: Y [H][W], U [H/2][W/2], V [H/2][W/2]
  const T*  cin  = src + (size_t)frame * srcFrameStride + (size_t)W * H + (size_t)( plane - 1 ) * cw * ch;
  uint16_t* cout = dst + (size_t)frame * dstFrameStride + (size_t)plane * W * H;
  for ( int s = threadIdx.x; s < levels; s += blockDim.x ) { lut[s] = sample_to_float( image_set( s, shift ), true, nbyte ); }
  const UpFilter& F = c_up[filter];
  float           v0[NV], v1[NV], h1[NH];
#pragma unroll
  for ( int t = 0; t < NV; t++ ) { v0[t] = F.v0[t], v1[t] = F.v1[t]; }
#pragma unroll
  for ( int t = 0; t < NH; t++ ) { h1[t] = F.h1[t]; }
  __syncthreads();
  const int ox = blockIdx.x * TW, oy = blockIdx.y * TH;  // output tile origin
  const int jx = ox / 2 - HALO, iy = oy / 2 - HALO;      // staged chroma origin
  for ( int k = threadIdx.x; k < IW * IH; k += blockDim.x ) {
    const int r = k / IW, c = k - r * IW;
    const int gy = min( max( iy + r, 0 ), ch - 1 ), gx = min( max( jx + c, 0 ), cw - 1 );  // clamp( ., 0, size - 1 ) of every tap
    const int s  = (int)cin[(size_t)gy * cw + gx];
    in[r][c]     = s < levels ? lut[s] : sample_to_float( image_set( s, shift ), true, nbyte );
  }
  __syncthreads();
  // vertical: row 2i from vertical0 at i0 = i, row 2i+1 from vertical1 at i0 = i + 1 (:678-686)
  const float   scale = 1.0f / 256.0f;
  constexpr int pv = ( NV + 1 ) >> 1, ph = ( NH + 1 ) >> 1;  // `position` of the reference
  for ( int k = threadIdx.x; k < TH * IW; k += blockDim.x ) {
    const int y = k / IW, c = k - y * IW;
    const int odd = y & 1, li = ( y >> 1 ) + HALO + odd - pv;
    float     value = 0.f;
#pragma unroll
    for ( int t = 0; t < NV; t++ ) { value = __fadd_rn( value, __fmul_rn( odd ? v1[t] : v0[t], in[li + t][c] ) ); }
    tmp[y][c] = __fmul_rn( __fadd_rn( value, 0.f ), scale );
  }
  __syncthreads();
  // horizontal: column 2j from horizontal0 at j0 = j, column 2j+1 from horizontal1 at j0 = j + 1 (:687-694);
  // every thread produces 8 consecutive outputs of one row.  horizontal0 = {0, 256} >> 8 in every filter of the table:
  // ( 0 * a + 256 * b ) / 256 is b exactly (the sign of a zero is lost in the reference too: 0 + -0 = +0)
  {
    const int y = threadIdx.x >> 3, x0 = ( threadIdx.x & 7 ) * 8;
    if ( oy + y < H && ox + x0 < W ) {
      uint16_t o[8];
#pragma unroll
      for ( int k = 0; k < 8; k += 2 ) {
        const int j = ( ( x0 + k ) >> 1 ) + HALO;
        o[k]        = float_to_u16( __fadd_rn( tmp[y][j], 0.f ), true );
        float value = 0.f;
#pragma unroll
        for ( int t = 0; t < NH; t++ ) { value = __fadd_rn( value, __fmul_rn( h1[t], tmp[y][j + 1 + t - ph] ) ); }
        o[k + 1] = float_to_u16( __fmul_rn( __fadd_rn( value, 0.f ), scale ), true );
      }
      *reinterpret_cast<uint4*>( cout + (size_t)( oy + y ) * W + ox + x0 ) = *reinterpret_cast<const uint4*>( o );
    }
  }
}

// geometry luma samples of one byte -> the uint16 plane the reprojection reads
__global__ void __launch_bounds__( 256 ) k_widen_u8( const uint8_t* __restrict__ src, uint16_t* __restrict__ dst, int64_t n ) {
  const int64_t i = ( (int64_t)blockIdx.x * blockDim.x + threadIdx.x ) * 16;
  if ( i + 16 <= n ) {
    const uint4 v = *reinterpret_cast<const uint4*>( src + i );
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t       o[8];
#pragma unroll
    for ( int k = 0; k < 4; k++ ) {
      o[2 * k]     = ( w[k] & 0xFFu ) | ( ( w[k] & 0xFF00u ) << 8 );
      o[2 * k + 1] = ( ( w[k] >> 16 ) & 0xFFu ) | ( ( w[k] >> 8 ) & 0xFF0000u );
    }
    uint4* d = reinterpret_cast<uint4*>( dst + i );
    d[0]     = make_uint4( o[0], o[1], o[2], o[3] );
    d[1]     = make_uint4( o[4], o[5], o[6], o[7] );
  } else {
    for ( int64_t j = i; j < n; j++ ) { dst[j] = src[j]; }
  }
}

// PCCImage<T, N>::convertBitdepth (PccLibCommon/source/PCCImage.cpp:258-299) of one sample held in a TOut:
// mode 0 none, 1: >> amount (msb aligned, input wider), 2: min( v, amount ) (input wider), 3: << amount (msb aligned,
// input narrower; the result is stored back into the sample type, i.e. truncated to it)
template <typename TOut>
__device__ __forceinline__ TOut convert_bitdepth( int v, int mode, int amount ) {
  if ( mode == 1 ) { return (TOut)( v >> amount ); }
  if ( mode == 2 ) { return (TOut)min( v, amount ); }
  if ( mode == 3 ) { return (TOut)( v << amount ); }
  return (TOut)v;
}

// geometry luma samples with PCCImage::set's shift, then convertBitdepth -> the uint16 plane (in == out allowed for
// 2-byte samples)
template <typename T>
__global__ void __launch_bounds__( 256 ) k_geometry_set( const T* src, uint16_t* dst, int64_t n, int shift, int mode, int amount ) {
  for ( int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x ) {
    dst[i] = convert_bitdepth<uint16_t>( (uint16_t)image_set( (int)src[i], shift ), mode, amount );
  }
}
// the occupancy video in place (PCCDecoder.cpp:119)
__global__ void __launch_bounds__( 256 ) k_occupancy_convert( uint8_t* v, int64_t n, int mode, int amount ) {
  for ( int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x ) {
    v[i] = convert_bitdepth<uint8_t>( v[i], mode, amount );
  }
}

// pitched decoder surfaces -> the dense planar layout of rb200_frames_yuv420 (luma rows packed; interleaved chroma
// split into a U and a V plane).  grid (x tiles, rows, surfaces); T = sample type, `shift` = right shift of 2-byte samples
struct SurfDev {
  const uint8_t* luma;
  const uint8_t* chroma;
  int32_t        pitch_luma, pitch_chroma;
};
template <typename T>
__global__ void __launch_bounds__( 256 ) k_gather_surfaces( const SurfDev* __restrict__ surf, T* __restrict__ dst, int W, int H,
                                                            size_t frameStride /* samples */, int withChroma, int shift ) {
  const SurfDev s = surf[blockIdx.z];
  const int     x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if ( x >= W ) { return; }
  T* out = dst + (size_t)blockIdx.z * frameStride;
  out[(size_t)y * W + x] = (T)( reinterpret_cast<const T*>( s.luma + (size_t)y * s.pitch_luma )[x] >> shift );
  if ( withChroma && y < H / 2 ) {  // row y of the interleaved chroma plane holds W samples: U0 V0 U1 V1 ...
    const T v = (T)( reinterpret_cast<const T*>( s.chroma + (size_t)y * s.pitch_chroma )[x] >> shift );
    T*      c = out + (size_t)W * H + ( ( x & 1 ) ? (size_t)( W / 2 ) * ( H / 2 ) : 0 );
    c[(size_t)y * ( W / 2 ) + ( x >> 1 )] = v;
  }
}

}  // namespace

// pitched NV12 / P010 surfaces (device or pinned memory) -> c->d_occ_video, c->d_raw_geo (or c->d_geometry for 2-byte
// samples) and c->d_raw_attr in the planar layout rb_ingest_yuv420_impl reads
int rb_gather_nv12_impl( rb200_ctx* c, const rb200_frames_nv12* fr ) {
  const int    F = c->F, M = c->M, nGA = F * M, bytes = fr->sample_bytes;
  const bool   attr = c->P.attribute_count > 0;
  const size_t plane = (size_t)c->W * c->H;
  const int    nSurf = F + nGA + ( attr ? nGA : 0 );
  SurfDev*     h = (SurfDev*)rb_pinned( c, (size_t)nSurf * sizeof( SurfDev ) );
  if ( !h ) { return rb_fail( c, RB200_ERR_NOMEM, "pinned allocation failed" ); }
  RB_CUDA( cudaStreamSynchronize( c->stream ) );  // the pinned block may still be read by an earlier copy
  auto put = [&]( int i, const rb200_surface& s ) {
    h[i] = SurfDev{(const uint8_t*)s.luma, (const uint8_t*)s.chroma, s.pitch_luma, s.pitch_chroma};
  };
  for ( int f = 0; f < F; f++ ) { put( f, fr->occupancy[f] ); }
  for ( int i = 0; i < nGA; i++ ) { put( F + i, fr->geometry[i] ); }
  for ( int i = 0; attr && i < nGA; i++ ) { put( F + nGA + i, fr->attribute[i] ); }
  RB_CUDA( c->d_scratch[0].ensure( (size_t)nSurf * sizeof( SurfDev ) + 256 ) );
  SurfDev* d = c->d_scratch[0].as<SurfDev>();
  RB_CUDA( cudaMemcpyAsync( d, h, (size_t)nSurf * sizeof( SurfDev ), cudaMemcpyHostToDevice, c->stream ) );
  c->stats.h2d_bytes += (int64_t)nSurf * sizeof( SurfDev );
  RB_LAUNCH( "gather_occupancy", k_gather_surfaces<uint8_t>, dim3( rb_div_up( c->oW, 256 ), c->oH, F ), 256, 0, d,
             c->d_occ_video.as<uint8_t>(), c->oW, c->oH, (size_t)c->oW * c->oH, 0, 0 );
  if ( bytes == 1 ) {
    RB_CUDA( c->d_raw_geo.ensure( (size_t)nGA * plane + 64 ) );
    RB_LAUNCH( "gather_geometry", k_gather_surfaces<uint8_t>, dim3( rb_div_up( c->W, 256 ), c->H, nGA ), 256, 0, d + F,
               c->d_raw_geo.as<uint8_t>(), c->W, c->H, plane, 0, 0 );
  } else {
    RB_LAUNCH( "gather_geometry", k_gather_surfaces<uint16_t>, dim3( rb_div_up( c->W, 256 ), c->H, nGA ), 256, 0, d + F,
               c->d_geometry.as<uint16_t>(), c->W, c->H, plane, 0, fr->sample_lsb_shift );
  }
  if ( attr ) {
    RB_CUDA( c->d_raw_attr.ensure( (size_t)nGA * ( plane + plane / 2 ) * bytes + 64 ) );
    if ( bytes == 1 ) {
      RB_LAUNCH( "gather_attribute", k_gather_surfaces<uint8_t>, dim3( rb_div_up( c->W, 256 ), c->H, nGA ), 256, 0, d + F + nGA,
                 c->d_raw_attr.as<uint8_t>(), c->W, c->H, plane + plane / 2, 1, 0 );
    } else {
      RB_LAUNCH( "gather_attribute", k_gather_surfaces<uint16_t>, dim3( rb_div_up( c->W, 256 ), c->H, nGA ), 256, 0, d + F + nGA,
                 c->d_raw_attr.as<uint16_t>(), c->W, c->H, plane + plane / 2, 1, fr->sample_lsb_shift );
    }
  }
  return RB200_OK;
}

// raw decoder planes (already in c->d_raw_geo / c->d_raw_attr) -> c->d_geometry / c->d_attribute
// (mode, amount) of convert_bitdepth for convertBitdepth( in, out, msbAlign ) on samples of `bits` bits
static void bitdepth_mode( int in, int out, int msb, int& mode, int& amount ) {
  mode = amount = 0;
  if ( out <= 0 ) { return; }
  const int diff = in - out;
  if ( diff >= 0 ) {
    mode   = msb ? 1 : 2;
    amount = msb ? diff : ( 1 << out ) - 1;
  } else if ( msb ) {
    mode   = 3;
    amount = -diff;
  }
  if ( mode == 1 && amount == 0 ) { mode = 0; }
}

int rb_ingest_yuv420_impl( rb200_ctx* c, int geo_bytes, int attr_bytes, int attr_bitdepth, int filter, int geo_shift, int attr_shift,
                           const int* geo_bd, const int* occ_bd ) {
  const size_t  plane = (size_t)c->W * c->H;
  const int64_t nGeo  = (int64_t)c->F * c->M * plane;
  int           gm, ga, om, oa;
  bitdepth_mode( geo_bd[0], geo_bd[1], geo_bd[2], gm, ga );
  bitdepth_mode( 8, occ_bd[0], occ_bd[1], om, oa );
  if ( gm == 2 && ga >= 65535 ) { gm = 0; }  // min( v, 2^16 - 1 ) on 16-bit samples
  if ( om == 2 && oa >= 255 ) { om = 0; }
  if ( om ) {
    RB_LAUNCH( "occupancy_convert", k_occupancy_convert, 148 * 4, 256, 0, c->d_occ_video.as<uint8_t>(), (int64_t)c->F * c->oW * c->oH, om, oa );
  }
  if ( geo_bytes == 1 && geo_shift == 0 && gm == 0 ) {
    RB_LAUNCH( "geometry_widen", k_widen_u8, rb_div_up( nGeo, 256 * 16 ), 256, 0, c->d_raw_geo.as<uint8_t>(), c->d_geometry.as<uint16_t>(), nGeo );
  } else if ( geo_bytes == 1 ) {
    RB_LAUNCH( "geometry_set", k_geometry_set<uint8_t>, 148 * 8, 256, 0, c->d_raw_geo.as<uint8_t>(), c->d_geometry.as<uint16_t>(), nGeo, geo_shift, gm, ga );
  } else if ( geo_shift > 0 || gm ) {  // 2-byte samples were copied straight into d_geometry
    RB_LAUNCH( "geometry_set", k_geometry_set<uint16_t>, 148 * 8, 256, 0, c->d_geometry.as<uint16_t>(), c->d_geometry.as<uint16_t>(), nGeo, geo_shift, gm, ga );
  }
  if ( c->P.attribute_count > 0 ) {
    const int    nbyte  = attr_bitdepth == 8 ? 1 : 2;
    const int    frames = c->F * c->M;
    const size_t sfs = plane + plane / 2, dfs = 3 * plane;
    const dim3   gl( 64, frames ), gc( rb_div_up( c->W, TW ), rb_div_up( c->H, TH ), frames * 2 );
    // tap counts of g_filter420to444[filter]: vertical0/1, horizontal1 (the table above)
    static const int NVt[8] = {4, 6, 4, 6, 8, 6, 10, 12}, NHt[8] = {4, 4, 4, 6, 8, 6, 10, 12};
    const int        nv = NVt[filter], nh = NHt[filter];
#define RB_CHROMA_CASE( TYPE, NV, NH )                                                                                     \
  if ( nv == NV && nh == NH ) {                                                                                            \
    RB_LAUNCH( "attribute_420_to_444", ( k_chroma_420_to_444<TYPE, NV, NH> ), gc, 256, 0, c->d_raw_attr.as<TYPE>(),        \
               c->d_attribute.as<uint16_t>(), c->W, c->H, nbyte, attr_shift, filter, sfs, dfs );                           \
  }
#define RB_CHROMA( TYPE )          \
  RB_CHROMA_CASE( TYPE, 4, 4 )     \
  RB_CHROMA_CASE( TYPE, 6, 4 )     \
  RB_CHROMA_CASE( TYPE, 6, 6 )     \
  RB_CHROMA_CASE( TYPE, 8, 8 )     \
  RB_CHROMA_CASE( TYPE, 10, 10 )   \
  RB_CHROMA_CASE( TYPE, 12, 12 )
    if ( attr_bytes == 1 ) {
      RB_LAUNCH( "attribute_luma_to_16", k_luma_to_16<uint8_t>, gl, 256, 0, c->d_raw_attr.as<uint8_t>(), c->d_attribute.as<uint16_t>(), c->W, c->H, nbyte, attr_shift, sfs, dfs );
      RB_CHROMA( uint8_t );
    } else {
      RB_LAUNCH( "attribute_luma_to_16", k_luma_to_16<uint16_t>, gl, 256, 0, c->d_raw_attr.as<uint16_t>(), c->d_attribute.as<uint16_t>(), c->W, c->H, nbyte, attr_shift, sfs, dfs );
      RB_CHROMA( uint16_t );
    }
  }
  return RB200_OK;
}
