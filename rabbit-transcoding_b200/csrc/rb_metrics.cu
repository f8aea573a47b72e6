// rb_metrics.cu — D1 / D2 / colour PSNR (PccLibMetrics) on the GPU.  Placeholder entry points: they fail loudly
// until the grid-hashed exact kNN lands (no CPU fallback).
#include "rb_common.cuh"

extern "C" {

int rb200_metrics( rb200_ctx* c, const rb200_metrics_params*, int, const rb200_cloud_view*, const rb200_cloud_view*,
                   rb200_metrics_result* ) {
  return rb_fail( c, RB200_ERR_UNSUPPORTED, "rb200_metrics is not implemented yet in this build" );
}

int rb200_remove_duplicates( rb200_ctx* c, const rb200_cloud_view*, int, int16_t*, uint8_t*, int64_t* ) {
  return rb_fail( c, RB200_ERR_UNSUPPORTED, "rb200_remove_duplicates is not implemented yet in this build" );
}
}
