// rb_io.cu — the wire format between PccAppDecoder and PccAppMetrics in transcode.sh, and the parity checksum.
//
// Restates
//   PCCPointSet3::computeChecksum( false ) / computeMd5     PccLibCommon/source/PCCPointSet.cpp:222-245
//   PCCPointSet3::write( fileName, asAscii = false )        :359-457   (binary little-endian PLY: float xyz + uchar rgb)
//   PCCPointSet3::read( fileName )                          :459-757   (ascii / binary little-endian, any scalar types)
// The 15-byte records are packed on the device, so a decoded frame goes to disk with one D2H copy and one write.
// MD5 (RFC 1321) is implemented here; the reference uses dependencies/libmd5 for the same digest.
#include <math.h>

#include <string>
#include <vector>

#include "rb_common.cuh"

namespace {

// ---- MD5, RFC 1321 ----
struct Md5 {
  uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
  uint64_t len = 0;
  uint8_t  buf[64];
  size_t   fill = 0;
  static uint32_t rol( uint32_t x, int s ) { return ( x << s ) | ( x >> ( 32 - s ) ); }
  void block( const uint8_t* p ) {
    static const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af,
        0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa,
        0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8,
        0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
        0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97,
        0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
        0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
    static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                              4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
    uint32_t M[16];
    for ( int i = 0; i < 16; i++ ) { M[i] = p[4 * i] | ( p[4 * i + 1] << 8 ) | ( p[4 * i + 2] << 16 ) | ( (uint32_t)p[4 * i + 3] << 24 ); }
    uint32_t A = a, B = b, C = c, D = d;
    for ( int i = 0; i < 64; i++ ) {
      uint32_t F;
      int      g;
      if ( i < 16 ) {
        F = ( B & C ) | ( ~B & D ), g = i;
      } else if ( i < 32 ) {
        F = ( D & B ) | ( ~D & C ), g = ( 5 * i + 1 ) & 15;
      } else if ( i < 48 ) {
        F = B ^ C ^ D, g = ( 3 * i + 5 ) & 15;
      } else {
        F = C ^ ( B | ~D ), g = ( 7 * i ) & 15;
      }
      const uint32_t t = D;
      D = C, C = B;
      B = B + rol( A + F + K[i] + M[g], S[i] );
      A = t;
    }
    a += A, b += B, c += C, d += D;
  }
  void update( const uint8_t* p, size_t n ) {
    len += n;
    while ( n ) {
      const size_t k = std::min( n, 64 - fill );
      memcpy( buf + fill, p, k );
      fill += k, p += k, n -= k;
      if ( fill == 64 ) {
        block( buf );
        fill = 0;
      }
    }
  }
  void finalize( uint8_t out[16] ) {
    const uint64_t bits = len * 8;
    const uint8_t  one  = 0x80, zero = 0;
    update( &one, 1 );
    while ( fill != 56 ) { update( &zero, 1 ); }
    uint8_t l[8];
    for ( int i = 0; i < 8; i++ ) { l[i] = (uint8_t)( bits >> ( 8 * i ) ); }
    update( l, 8 );
    const uint32_t v[4] = {a, b, c, d};
    for ( int i = 0; i < 4; i++ ) {
      for ( int k = 0; k < 4; k++ ) { out[4 * i + k] = (uint8_t)( v[i] >> ( 8 * k ) ); }
    }
  }
};

// 15-byte PLY vertex records: float x, y, z (the int16 coordinate converted to float) + uchar r, g, b
__global__ void k_pack_ply( const short4* __restrict__ pos, const uchar4* __restrict__ rgb, int64_t n, uint8_t* __restrict__ out ) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ( i >= n ) { return; }
  const short4 p = pos[i];
  const uchar4 c = rgb[i];
  const float  v[3] = {(float)p.x, (float)p.y, (float)p.z};
  uint8_t*     o = out + i * 15;
  const uint8_t* b = reinterpret_cast<const uint8_t*>( v );
#pragma unroll
  for ( int k = 0; k < 12; k++ ) { o[k] = b[k]; }
  o[12] = c.x, o[13] = c.y, o[14] = c.z;
}

}  // namespace

extern "C" {

int rb200_frame_md5( rb200_ctx* c, int f, uint8_t* out16 ) {
  if ( !c || !out16 ) { return RB200_ERR_INVALID; }
  if ( !c->reconstructed ) { return rb_fail( c, RB200_ERR_STATE, "frame_md5 before reconstruct" ); }
  if ( f < 0 || f >= c->F ) { return rb_fail( c, RB200_ERR_INVALID, "frame index out of range" ); }
  const int64_t        n = c->h_frame_off[f + 1] - c->h_frame_off[f];
  std::vector<int16_t> pos( (size_t)n * 3 );
  std::vector<uint8_t> col( (size_t)n * 3 );
  rb200_cloud_host     h{};
  h.positions = pos.data();
  h.colors    = col.data();
  const int r = n ? rb200_download_frame( c, f, &h ) : RB200_OK;
  if ( r ) { return r; }
  Md5 md5;
  md5.update( reinterpret_cast<const uint8_t*>( pos.data() ), pos.size() * 2 );
  if ( c->P.attribute_count > 0 || c->rgb_done ) { md5.update( col.data(), col.size() ); }  // withColors_
  md5.finalize( out16 );
  return RB200_OK;
}

// PCCPointSet3::computeChecksum( true ) (PCCPointSet.cpp:221-229): reorder( out, dropDuplicates = true ) (:258-296) — the
// points in lexicographic (x, y, z) order, one per position, colour = integer mean of the duplicates — then computeMd5.
// That is removeDuplicate( out, 2 )'s result (:169-218), so the device de-duplication of the metrics path produces it.
int rb200_frame_md5_canonical( rb200_ctx* c, int f, uint8_t* out16 ) {
  if ( !c || !out16 ) { return RB200_ERR_INVALID; }
  if ( !c->reconstructed || !c->rgb_done ) { return rb_fail( c, RB200_ERR_STATE, "frame_md5_canonical needs a decoded GOF (RGB8 done)" ); }
  if ( f < 0 || f >= c->F ) { return rb_fail( c, RB200_ERR_INVALID, "frame index out of range" ); }
  if ( c->P.attribute_count == 0 ) {
    return rb_fail( c, RB200_ERR_UNSUPPORTED, "frame_md5_canonical: clouds without colours keep their duplicates (PCCPointSet.cpp:287-295); not implemented" );
  }
  const int64_t        n = c->h_frame_off[f + 1] - c->h_frame_off[f];
  std::vector<int16_t> pos( (size_t)n * 3 );
  std::vector<uint8_t> col( (size_t)n * 3 );
  int64_t              m = 0;
  if ( n ) {
    rb200_cloud_host h{};
    h.positions = pos.data();
    h.colors    = col.data();
    int r       = rb200_download_frame( c, f, &h );
    if ( r ) { return r; }
    rb200_cloud_view v{pos.data(), col.data(), nullptr, n};
    std::vector<int16_t> upos( (size_t)n * 3 );
    std::vector<uint8_t> ucol( (size_t)n * 3 );
    r = rb200_remove_duplicates( c, &v, 2, upos.data(), ucol.data(), &m );
    if ( r ) { return r; }
    pos.swap( upos );
    col.swap( ucol );
  }
  Md5 md5;
  md5.update( reinterpret_cast<const uint8_t*>( pos.data() ), (size_t)m * 6 );
  md5.update( col.data(), (size_t)m * 3 );
  md5.finalize( out16 );
  return RB200_OK;
}

int rb200_write_ply( rb200_ctx* c, int f, const char* path ) {
  if ( !c || !path ) { return RB200_ERR_INVALID; }
  if ( !c->reconstructed || !c->rgb_done ) { return rb_fail( c, RB200_ERR_STATE, "write_ply needs a decoded GOF (RGB8 done)" ); }
  if ( f < 0 || f >= c->F ) { return rb_fail( c, RB200_ERR_INVALID, "frame index out of range" ); }
  cudaSetDevice( c->device );
  const int64_t b = c->h_frame_off[f], n = c->h_frame_off[f + 1] - b;
  std::vector<uint8_t> rec( (size_t)n * 15 );
  if ( n ) {
    RB_CUDA( c->d_pack.ensure( (size_t)n * 15 ) );
    RB_LAUNCH( "pack_ply", k_pack_ply, rb_div_up( n, 256 ), 256, 0, c->d_pos.as<short4>() + b, c->d_rgb.as<uchar4>() + b, n,
               c->d_pack.as<uint8_t>() );
    RB_CUDA( cudaMemcpyAsync( rec.data(), c->d_pack.p, (size_t)n * 15, cudaMemcpyDeviceToHost, c->stream ) );
    RB_CUDA( cudaStreamSynchronize( c->stream ) );
    c->stats.d2h_bytes += n * 15;
  }
  FILE* fp = fopen( path, "wb" );
  if ( !fp ) { return rb_fail( c, RB200_ERR_INVALID, "write_ply: cannot open %s", path ); }
  // header exactly as PCCPointSet3::write emits it for a coloured cloud without normals (:363-409)
  fprintf( fp, "ply\nformat binary_little_endian 1.0\nelement vertex %lld\nproperty float x\nproperty float y\nproperty float z\n"
               "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face 0\n"
               "property list uint8 int32 vertex_index\nend_header\n", (long long)n );
  const size_t w = n ? fwrite( rec.data(), 15, (size_t)n, fp ) : 0;
  fclose( fp );
  if ( (int64_t)w != n ) { return rb_fail( c, RB200_ERR_INVALID, "write_ply: short write to %s", path ); }
  return RB200_OK;
}

// PCCPointSet3::read: header parsing (:489-612) and the vertex payload (:613-757) for the scalar properties x, y, z
// (float32 / float64 / intN) and red, green, blue (uchar).  Other properties are skipped by their byte size.
int rb200_read_ply( const char* path, int16_t* outPos, uint8_t* outCol, int64_t capacity, int64_t* outCount, int* hasColors ) {
  if ( !path || !outCount ) { return RB200_ERR_INVALID; }
  FILE* fp = fopen( path, "rb" );
  if ( !fp ) { return RB200_ERR_INVALID; }
  struct Prop {
    std::string name;
    int         kind;  // 0 float32, 1 float64, 2 uint, 3 int
    int         bytes;
  };
  std::vector<Prop> props;
  char              line[4096];
  bool              ascii = false, vertexProps = true, ok = false;
  int64_t           n = 0;
  if ( !fgets( line, sizeof( line ), fp ) || strncmp( line, "ply", 3 ) != 0 ) {
    fclose( fp );
    return RB200_ERR_INVALID;
  }
  while ( fgets( line, sizeof( line ), fp ) ) {
    char a[64] = {0}, b[64] = {0}, d[64] = {0};
    const int k = sscanf( line, "%63s %63s %63s", a, b, d );
    if ( k <= 0 || !strcmp( a, "comment" ) ) { continue; }
    if ( !strcmp( a, "format" ) ) {
      ascii = !strcmp( b, "ascii" );
      if ( !ascii && strcmp( b, "binary_little_endian" ) != 0 ) {
        fclose( fp );
        return RB200_ERR_UNSUPPORTED;
      }
    } else if ( !strcmp( a, "element" ) ) {
      if ( !strcmp( b, "vertex" ) ) {
        n = atoll( d );
      } else {
        vertexProps = false;
      }
    } else if ( !strcmp( a, "property" ) && vertexProps && k == 3 ) {
      Prop p;
      p.name = d;
      const std::string t = b;
      if ( t == "float" || t == "float32" ) {
        p.kind = 0, p.bytes = 4;
      } else if ( t == "float64" || t == "double" ) {
        p.kind = 1, p.bytes = 8;
      } else if ( t == "uchar" || t == "uint8" ) {
        p.kind = 2, p.bytes = 1;
      } else if ( t == "uint16" || t == "ushort" ) {
        p.kind = 2, p.bytes = 2;
      } else if ( t == "uint32" || t == "uint" ) {
        p.kind = 2, p.bytes = 4;
      } else if ( t == "uint64" ) {
        p.kind = 2, p.bytes = 8;
      } else if ( t == "int8" || t == "char" ) {
        p.kind = 3, p.bytes = 1;
      } else if ( t == "int16" || t == "short" ) {
        p.kind = 3, p.bytes = 2;
      } else if ( t == "int32" || t == "int" ) {
        p.kind = 3, p.bytes = 4;
      } else if ( t == "int64" ) {
        p.kind = 3, p.bytes = 8;
      } else {
        fclose( fp );
        return RB200_ERR_UNSUPPORTED;
      }
      props.push_back( p );
    } else if ( !strcmp( a, "end_header" ) ) {
      ok = true;
      break;
    }
  }
  if ( !ok ) {
    fclose( fp );
    return RB200_ERR_INVALID;
  }
  int ix[3] = {-1, -1, -1}, ic[3] = {-1, -1, -1};
  for ( size_t i = 0; i < props.size(); i++ ) {
    const std::string& s = props[i].name;
    if ( s == "x" ) { ix[0] = (int)i; }
    if ( s == "y" ) { ix[1] = (int)i; }
    if ( s == "z" ) { ix[2] = (int)i; }
    if ( s == "red" || s == "r" ) { ic[0] = (int)i; }
    if ( s == "green" || s == "g" ) { ic[1] = (int)i; }
    if ( s == "blue" || s == "b" ) { ic[2] = (int)i; }
  }
  const bool colours = ic[0] >= 0 && ic[1] >= 0 && ic[2] >= 0;
  if ( hasColors ) { *hasColors = colours ? 1 : 0; }
  *outCount = n;
  if ( ix[0] < 0 || ix[1] < 0 || ix[2] < 0 ) {
    fclose( fp );
    return RB200_ERR_INVALID;
  }
  if ( !outPos || n > capacity ) {  // count only
    fclose( fp );
    return n > capacity && outPos ? RB200_ERR_NOMEM : RB200_OK;
  }
  size_t stride = 0;
  for ( auto& p : props ) { stride += p.bytes; }
  std::vector<uint8_t> row( stride );
  std::vector<double>  val( props.size() );
  for ( int64_t i = 0; i < n; i++ ) {
    if ( ascii ) {
      for ( size_t k = 0; k < props.size(); k++ ) {
        if ( fscanf( fp, "%lf", &val[k] ) != 1 ) {
          fclose( fp );
          return RB200_ERR_INVALID;
        }
      }
    } else {
      if ( fread( row.data(), 1, stride, fp ) != stride ) {
        fclose( fp );
        return RB200_ERR_INVALID;
      }
      size_t o = 0;
      for ( size_t k = 0; k < props.size(); k++ ) {
        const Prop& p = props[k];
        double      v = 0;
        if ( p.kind == 0 ) {
          float x;
          memcpy( &x, &row[o], 4 );
          v = x;
        } else if ( p.kind == 1 ) {
          memcpy( &v, &row[o], 8 );
        } else {
          uint64_t u = 0;
          memcpy( &u, &row[o], p.bytes );
          if ( p.kind == 3 && p.bytes < 8 && ( u >> ( 8 * p.bytes - 1 ) ) ) { u |= ~0ull << ( 8 * p.bytes ); }
          v = p.kind == 3 ? (double)(int64_t)u : (double)u;
        }
        val[k] = v;
        o += p.bytes;
      }
    }
    for ( int k = 0; k < 3; k++ ) { outPos[3 * i + k] = (int16_t)val[ix[k]]; }  // PCCType is int16_t: the cast truncates
    if ( colours && outCol ) {
      for ( int k = 0; k < 3; k++ ) { outCol[3 * i + k] = (uint8_t)val[ic[k]]; }
    }
  }
  fclose( fp );
  return RB200_OK;
}

}  // extern "C"
