/* Serial stand-in for the two TBB constructs the reference's hot-path translation units name
 * (tbb::task_arena::execute and tbb::parallel_for; PCCCodec.cpp:1114-1155, PCCGroupOfFrames.cpp:53-78,
 * PCCMetrics.cpp:37 only includes the header).  None of those loops is reached by the decoder path the
 * oracle runs (grid smoothing is on, PLY I/O is not used), so building the vendored TBB 2019 (which
 * needs cmake) is unnecessary; this header lets the unmodified reference sources compile with g++ alone.
 * Test infrastructure only. */
#ifndef RABBIT_B200_ORACLE_TBB_STUB_H
#define RABBIT_B200_ORACLE_TBB_STUB_H
#include <cstddef>
namespace tbb {
class task_arena {
 public:
  explicit task_arena( int = 1 ) {}
  template <typename F>
  void execute( F&& f ) {
    f();
  }
};
template <typename Index, typename F>
void parallel_for( Index first, Index last, F&& f ) {
  for ( Index i = first; i < last; ++i ) { f( i ); }
}
}  // namespace tbb
#endif
