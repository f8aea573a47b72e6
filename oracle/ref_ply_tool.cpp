// ref_ply_tool.cpp — tiny command-line front end of the harness' PLY helpers (PCCPointSet3::write / read of the
// unmodified reference).  TEST INFRASTRUCTURE ONLY.  A separate process because the reference formats numbers through
// iostreams, which the statically linked libstdc++ of oracle/_ref cannot do inside a Python process.
//   ref_ply_tool write <n> <positions.i16> <colors.u8> <out.ply>
//   ref_ply_tool read  <in.ply> <positions.i16> <colors.u8>          (prints the point count)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rabbit_b200.h"

extern "C" int     ref_write_ply( const rb200_cloud_view* in, const char* path );
extern "C" int64_t ref_read_ply( const char* path, int16_t* outPos, uint8_t* outCol, int64_t cap );

static std::vector<uint8_t> slurp( const char* p ) {
  std::vector<uint8_t> v;
  FILE*                f = fopen( p, "rb" );
  if ( !f ) { return v; }
  fseek( f, 0, SEEK_END );
  v.resize( ftell( f ) );
  fseek( f, 0, SEEK_SET );
  if ( !v.empty() && fread( v.data(), 1, v.size(), f ) != v.size() ) { v.clear(); }
  fclose( f );
  return v;
}

int main( int argc, char** argv ) {
  if ( argc == 6 && !strcmp( argv[1], "write" ) ) {
    const int64_t n = atoll( argv[2] );
    auto          p = slurp( argv[3] ), c = slurp( argv[4] );
    if ( (int64_t)p.size() != n * 6 || (int64_t)c.size() != n * 3 ) { return 2; }
    rb200_cloud_view v{reinterpret_cast<const int16_t*>( p.data() ), c.data(), nullptr, n};
    return ref_write_ply( &v, argv[5] );
  }
  if ( argc == 5 && !strcmp( argv[1], "read" ) ) {
    const int64_t n = ref_read_ply( argv[2], nullptr, nullptr, 0 );
    if ( n < 0 ) { return 2; }
    std::vector<int16_t> p( n * 3 );
    std::vector<uint8_t> c( n * 3 );
    ref_read_ply( argv[2], p.data(), c.data(), n );
    FILE* f = fopen( argv[3], "wb" );
    fwrite( p.data(), 2, p.size(), f );
    fclose( f );
    f = fopen( argv[4], "wb" );
    fwrite( c.data(), 1, c.size(), f );
    fclose( f );
    printf( "%lld\n", (long long)n );
    return 0;
  }
  return 1;
}
