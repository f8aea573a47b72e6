// PCCCodecB200.cpp — reference-side shim: the reference's own member functions, same signatures, bodies replaced by
// calls into the C ABI of include/rabbit_b200.h.
//
// Compiles against the UNMODIFIED reference headers (/root/reference/source/lib/*/include); it is host glue, not part
// of the CUDA library.  Replaced bodies (reference file:line of the originals):
//   PCCCodec::generateOccupancyMap                         PccLibCommon/source/PCCCodec.cpp:1584-1606
//   PCCCodec::generateBlockToPatchFromOccupancyMapVideo    :1725-1763
//   PCCCodec::generatePointCloud                           :517-978
//   PCCCodec::colorPointCloud                              :1308-1449
//   PCCCodec::smoothPointCloudPostprocess                  :52-147
//   PCCCodec::colorSmoothing                               :149-236
//   PCCPointSet3::transferColors16bitBP                    PccLibCommon/source/PCCPointSet.cpp:1126-1485
//   PCCMetrics::compute( sources, reconstructs, normals )  PccLibMetrics/source/PCCMetrics.cpp:334-369
//   PCCMetrics::compute( source, reconstruct, normals )    :371-385
// There is NO fallback into the reference's bodies: the shim never calls them (in the test build their definitions are
// weakened in the copied objects, oracle/Makefile, so the strong definitions below are the ones every caller binds to;
// nothing named rb200_orig_* exists), every result comes from the CUDA library, and a mode the CUDA path does not
// implement (multiple tiles,
// auxiliary video, PBF, other transfer-filter arguments ...) ends the way the reference ends on an error: a message and
// exit( -1 ) (PCCMetrics.cpp:342-346, PCCPatch.cpp:237-245).
//
// Per-frame calls, per-GOF execution: the reference calls these functions frame by frame; the CUDA path processes the
// whole GOF in one batched launch sequence.  The first call of a stage for a GOF runs that stage for every frame on
// the GPU and brings the stage's result of the whole GOF back with ONE packed copy per field into pinned staging;
// every call then fills the caller's containers for the frame it was asked for from that staging.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <fstream>
#include <iostream>
#include <list>
#include <memory>
#include <mutex>
#include <queue>
#include <sstream>
#include <unordered_map>
#include <algorithm>
#include <array>
#include <cmath>
#include <chrono>
#include <limits>
#include <numeric>
#include <set>
#include <stack>
#include <thread>
#include <iomanip>
#include <cassert>
#include <sys/time.h>
#include <sys/resource.h>

// QualityMetrics keeps its result floats private and has no setter; the shim fills them like the original body does
#define private public
#define protected public
#include "PCCCommon.h"
#include "PCCImage.h"
#include "PCCVideo.h"
#include "PCCPointSet.h"
#include "PCCPatch.h"
#include "PCCContext.h"
#include "PCCFrameContext.h"
#include "PCCGroupOfFrames.h"
#include "PCCCodec.h"
#include "PCCMetricsParameters.h"
#include "PCCMetrics.h"
#undef private
#undef protected

#include "rabbit_b200.h"
#include "rb200_atlas_export.h"

using namespace pcc;

namespace {

struct Pinned {  // grow-only pinned staging (rb200_host_alloc): transfers to / from it are DMA, and it is reused GOF after GOF
  void*  p   = nullptr;
  size_t cap = 0;
  template <typename T>
  T* get( size_t count ) {
    const size_t bytes = count * sizeof( T );
    if ( bytes > cap ) {
      rb200_host_free( p );
      cap = bytes + bytes / 8 + 4096;
      p   = rb200_host_alloc( cap );
      if ( !p ) {
        std::fprintf( stderr, "rabbit_b200: pinned host allocation of %zu bytes failed\n", cap );
        std::exit( -1 );
      }
    }
    return static_cast<T*>( p );
  }
};

struct Gof {  // the GOF currently resident on the GPU
  rb200_ctx*                      ctx     = nullptr;
  PCCContext*                     context = nullptr;
  size_t                          frames  = 0;
  rb200_params                    P{};
  GeneratePointCloudParameters    gp;
  bool                            bDecoder      = false;
  bool                            reconstructed = false, geo = false, transfer = false, color = false;
  std::vector<rb200_frame_counts> counts;
  std::vector<size_t>             off;  // [frames + 1] first point of every frame in the staging arrays
  std::map<size_t, std::vector<uint8_t>> occOriginal;  // occupancy video of the frames seen by generateOccupancyMap
  bool                            occFresh = false;    // generateOccupancyMap ran for the frame block-to-patch is called for
  std::set<size_t>                rawSeen;             // frames of this GOF that skipped it (occupancy synthesis)
  size_t                          thresholdLossyOM = 0;
  size_t                          currentFrame     = 0;
  // pinned staging: the planes going up, and per stage the fields coming down (whole GOF, frame after frame)
  Pinned inOcc, inGeo, inAtt, inAuxGeo, inAuxAtt;
  Pinned pos[2], typ[2];  // after reconstruction / after geometry smoothing
  Pinned col[3];          // colours16 after reconstruction / attribute re-transfer / colour smoothing
  Pinned part, p2p;
};
Gof g;  // the decoder's frame loop is single-threaded (PCCDecoder.cpp:330) and PCCCodec is not re-entrant per instance

[[noreturn]] void die( int status ) {
  // the reference prints and exits (PCCPatch.cpp:237-245 exits 180; everything else -1)
  std::fprintf( stderr, "rabbit_b200: %s\n", rb200_error_string( g.ctx ) );
  std::exit( status == RB200_ERR_PATCH_OUT_OF_CANVAS ? 180 : -1 );
}
[[noreturn]] void unsupported( const char* what ) {
  std::fprintf( stderr, "rabbit_b200: %s is not implemented on the CUDA path (there is no CPU fallback)\n", what );
  std::exit( -1 );
}
#define RB( call )                             \
  do {                                         \
    int st__ = ( call );                       \
    if ( st__ != RB200_OK ) { die( st__ ); }   \
  } while ( 0 )

void ensureContext() {
  if ( g.ctx ) { return; }
  const char* dev = std::getenv( "RB200_DEVICE" );  // the GPU of this process (one process per GPU)
  const int   st  = rb200_create( dev ? std::atoi( dev ) : 0, &g.ctx );
  if ( st != RB200_OK ) {
    std::fprintf( stderr, "rabbit_b200: no usable CUDA device %s (status %d); there is no CPU fallback\n", dev ? dev : "0", st );
    std::exit( -1 );
  }
}

void checkSupported( PCCContext& context, const GeneratePointCloudParameters& p, size_t tileIndex ) {
  if ( p.pbfEnableFlag_ && ( p.enhancedOccupancyMapCode_ || p.singleMapPixelInterleaving_ || p.pointLocalReconstruction_ ||
                             p.enableSizeQuantization_ ) ) {
    unsupported( "occupancy synthesis (PCCCodec.cpp:541-554) together with EOM, pixel interleaving, point local reconstruction or "
                 "patch size quantisation" );
  }
  if ( p.mapCountMinus1_ > 1 ) { unsupported( "more than two maps" ); }
  if ( p.occupancyResolution_ != 16 ) { unsupported( "an occupancy resolution other than 16" ); }
  if ( ( p.pointLocalReconstruction_ || p.singleMapPixelInterleaving_ ) &&
       ( p.mapCountMinus1_ != 0 || p.enhancedOccupancyMapCode_ || p.multipleStreams_ ) ) {
    unsupported( "pixel interleaving / point local reconstruction together with two maps, EOM or multiple streams" );
  }
  if ( tileIndex != 0 ) { unsupported( "an atlas frame with several tiles (PCCDecoder.cpp:356-381)" ); }
  for ( size_t f = 0; f < context.size(); f++ ) {
    if ( context[f].getNumTilesInAtlasFrame() != 1 ) { unsupported( "an atlas frame with several tiles (PCCDecoder.cpp:356-381)" ); }
    auto& tile = context[f].getTile( 0 );
    if ( tile.getLeftTopXInFrame() != 0 || tile.getLeftTopYInFrame() != 0 ) { unsupported( "a tile that does not start at (0, 0)" ); }
    if ( tile.getUseRawPointsSeparateVideo() != p.useAuxSeperateVideo_ ) {
      unsupported( "a tile whose auxiliary-video flag differs from the sequence's (PCCCodec.cpp:606 vs :880)" );
    }
  }
}

// PCCContext -> flat structs of the C ABI (planes through pinned staging), upload, reconstruction of every frame,
// one packed download of everything generatePointCloud / colorPointCloud hand back
void reconstructGof( PCCContext& context, const GeneratePointCloudParameters& gp, bool bDecoder, bool relativeT1 ) {
  const size_t F = context.size();
  const size_t W = context[0].getAtlasFrameWidth(), H = context[0].getAtlasFrameHeight();
  const size_t M = gp.mapCountMinus1_ + 1, P = gp.occupancyPrecision_;
  const size_t oW = W / P, oH = H / P;
  auto&        ai       = context.getVps().getAttributeInformation( 0 );
  const bool   hasAttr  = ai.getAttributeCount() > 0;
  const bool   streams  = gp.multipleStreams_;  // map m of frame f = frame f of video m (PCCCodec.cpp:609-618)
  auto&        geoVideos = context.getVideoGeometryMultiple();
  auto&        attVideos = context.getVideoAttributesMultiple();
  rb200_params& p = g.P;
  std::memset( &p, 0, sizeof( p ) );
  p.width = (int)W, p.height = (int)H;
  p.occupancy_resolution        = (int)gp.occupancyResolution_;
  p.occupancy_precision         = (int)P;
  // with occupancy synthesis the decoder never calls generateOccupancyMap (PCCDecoder.cpp:362-366): the threshold
  // travels in the parameters (PCCCodec.cpp:551)
  p.threshold_lossy_om          = gp.pbfEnableFlag_ ? (int)gp.thresholdLossyOM_ : (int)g.thresholdLossyOM;
  p.pbf_enable                  = gp.pbfEnableFlag_;
  p.pbf_passes_count            = gp.pbfPassesCount_;
  p.pbf_filter_size             = gp.pbfFilterSize_;
  p.pbf_log2_threshold          = gp.pbfLog2Threshold_;
  p.map_count_minus1            = (int)gp.mapCountMinus1_;
  p.absolute_d1                 = gp.absoluteD1_;
  p.remove_duplicate_points     = gp.removeDuplicatePoints_;
  p.enhanced_occupancy_map_code = gp.enhancedOccupancyMapCode_;
  p.eom_fix_bit_count           = (int)gp.EOMFixBitCount_;
  p.enable_size_quantization    = gp.enableSizeQuantization_;
  p.log2_quantizer_x            = (int)context[0].getTile( 0 ).getLog2PatchQuantizerSizeX();
  p.log2_quantizer_y            = (int)context[0].getTile( 0 ).getLog2PatchQuantizerSizeY();
  auto& asps                    = context.getAtlasSequenceParameterSet( 0 );
  p.patch_precedence_reverse    = bDecoder && asps.getPatchPrecedenceOrderFlag();
  p.use_additional_points_patch = gp.useAdditionalPointsPatch_;
  p.single_map_pixel_interleaving = gp.singleMapPixelInterleaving_;
  p.point_local_reconstruction    = gp.pointLocalReconstruction_;
  p.surface_thickness             = (int)gp.surfaceThickness_;
  p.multiple_streams              = streams ? 1 : 0;
  p.relative_t1                   = ( streams && M > 1 && relativeT1 ) ? 1 : 0;
  p.attribute_count             = hasAttr ? 1 : 0;
  p.attribute_rgb444 = hasAttr && attVideos[0].getFrameCount() > 0 && attVideos[0].getColorFormat() == PCCCOLORFORMAT::RGB444;
  p.geometry_bitdepth_3d       = (int)gp.geometryBitDepth3D_;
  p.flag_geometry_smoothing    = gp.flagGeometrySmoothing_;
  p.grid_smoothing             = gp.gridSmoothing_;
  p.grid_size                  = (int)gp.gridSize_;
  p.apply_geo_smoothing        = 1;  // the caller decides which stage functions it calls
  p.attr_transfer_filter_type  = 1;  // keep the pre-smoothing cloud: transferColors16bitBP may follow
  p.flag_color_smoothing       = gp.flagColorSmoothing_;
  p.apply_attr_smoothing       = 1;
  p.threshold_smoothing        = gp.thresholdSmoothing_;
  p.neighbor_count_smoothing   = gp.gridSmoothing_ ? 0 : (int)gp.neighborCountSmoothing_;
  p.radius2_smoothing          = gp.radius2Smoothing_;
  p.radius2_boundary_detection = gp.radius2BoundaryDetection_;
  p.threshold_color_smoothing  = gp.thresholdColorSmoothing_;
  p.threshold_color_difference = gp.thresholdColorDifference_;
  p.threshold_color_variation  = gp.thresholdColorVariation_;

  uint8_t*  occ = g.inOcc.get<uint8_t>( F * oW * oH );
  uint16_t* geo = g.inGeo.get<uint16_t>( F * M * W * H );
  uint16_t* att = hasAttr ? g.inAtt.get<uint16_t>( F * M * 3 * W * H ) : nullptr;
  auto&     occVideo = context.getVideoOccupancyMap();
  for ( size_t f = 0; f < F; f++ ) {
    auto it = g.occOriginal.find( f );  // a frame generateOccupancyMap already binarised in place: use the saved copy
    if ( it != g.occOriginal.end() ) {
      std::memcpy( &occ[f * oW * oH], it->second.data(), oW * oH );
    } else {
      std::memcpy( &occ[f * oW * oH], occVideo.getFrame( f ).getChannel( 0 ).data(), oW * oH );
    }
    for ( size_t m = 0; m < M; m++ ) {
      auto& gf = streams ? geoVideos[m].getFrame( f ) : geoVideos[0].getFrame( f * M + m );
      std::memcpy( &geo[( f * M + m ) * W * H], gf.getChannel( 0 ).data(), W * H * 2 );
      if ( hasAttr ) {
        auto& a = streams ? attVideos[m].getFrame( f ) : attVideos[0].getFrame( f * M + m );
        for ( int c = 0; c < 3; c++ ) { std::memcpy( &att[( ( f * M + m ) * 3 + c ) * W * H], a.getChannel( c ).data(), W * H * 2 ); }
      }
    }
  }
  rb200::AtlasTables tables;  // the atlas layer's patches as flat rows (rb200_atlas_export.h)
  rb200::exportAtlas( context, tables );
  rb200_frames fr{occ, geo, att, nullptr, nullptr};
  if ( gp.useAuxSeperateVideo_ && ( gp.useAdditionalPointsPatch_ || gp.enhancedOccupancyMapCode_ ) ) {
    // raw / EOM points in the auxiliary video: channel 0 of context.getVideoRawPointsGeometry() (raw coordinates,
    // PCCCodec.cpp:895-897), the three channels of getVideoRawPointsAttribute() (raw and EOM colours, :1524-1580)
    auto&        rawGeo = context.getVideoRawPointsGeometry();
    auto&        rawAtt = context.getVideoRawPointsAttribute();
    const bool   haveGeo = gp.useAdditionalPointsPatch_ && rawGeo.getFrameCount() >= F;
    const size_t Wa = haveGeo ? rawGeo.getFrame( 0 ).getWidth() : rawAtt.getFrame( 0 ).getWidth(),
                 Ha = haveGeo ? rawGeo.getFrame( 0 ).getHeight() : rawAtt.getFrame( 0 ).getHeight();
    p.use_aux_separate_video = 1;
    p.aux_width = (int)Wa, p.aux_height = (int)Ha;
    if ( haveGeo ) {
      uint16_t* ag = g.inAuxGeo.get<uint16_t>( F * Wa * Ha );
      for ( size_t f = 0; f < F; f++ ) { std::memcpy( &ag[f * Wa * Ha], rawGeo.getFrame( f ).getChannel( 0 ).data(), Wa * Ha * 2 ); }
      fr.aux_geometry = ag;
    }
    if ( hasAttr ) {
      uint16_t* aa     = g.inAuxAtt.get<uint16_t>( F * 3 * Wa * Ha );
      for ( size_t f = 0; f < F; f++ ) {
        for ( int c = 0; c < 3; c++ ) { std::memcpy( &aa[( f * 3 + c ) * Wa * Ha], rawAtt.getFrame( f ).getChannel( c ).data(), Wa * Ha * 2 ); }
      }
      fr.aux_attribute = aa;
    }
  }
  rb200_atlas  at = tables.view();
  ensureContext();
  RB( rb200_gof_begin( g.ctx, &p, (int)F ) );
  RB( rb200_gof_upload( g.ctx, &fr, &at ) );
  if ( gp.pointLocalReconstruction_ && !gp.singleMapPixelInterleaving_ ) {
    // what PCCDecoder::setPointLocalReconstruction / setPLRData (PCCDecoder.cpp:528-591) left in the context and patches
    std::vector<rb200_plr_mode> modes;
    for ( size_t i = 0; i < context.getPointLocalReconstructionModeNumber(); i++ ) {
      const auto& m = context.getPointLocalReconstructionMode( i );
      modes.push_back( rb200_plr_mode{(uint8_t)m.interpolate_, (uint8_t)m.filling_, m.minD1_, m.neighbor_} );
    }
    std::vector<uint8_t> blockMode;
    std::vector<int64_t> blockOff{0};
    for ( size_t f = 0; f < F; f++ ) {
      for ( auto& s : context[f].getTile( 0 ).getPatches() ) {
        for ( size_t v0 = 0; v0 < s.getSizeV0(); v0++ ) {
          for ( size_t u0 = 0; u0 < s.getSizeU0(); u0++ ) { blockMode.push_back( s.getPointLocalReconstructionMode( u0, v0 ) ); }
        }
        blockOff.push_back( (int64_t)blockMode.size() );
      }
    }
    if ( blockMode.empty() ) { blockMode.push_back( 0 ); }
    rb200_plr plr{(int32_t)modes.size(), modes.data(), blockMode.data(), blockOff.data()};
    RB( rb200_gof_set_plr( g.ctx, &plr ) );
  }
  RB( rb200_reconstruct( g.ctx ) );
  g.counts.resize( F );
  RB( rb200_frame_counts_get( g.ctx, g.counts.data() ) );
  g.off.assign( F + 1, 0 );
  for ( size_t f = 0; f < F; f++ ) { g.off[f + 1] = g.off[f] + (size_t)g.counts[f].total; }
  const size_t N = g.off[F];
  if ( N ) {
    rb200_cloud_host h{};
    h.positions      = g.pos[0].get<int16_t>( N * 3 );
    h.boundary_types = g.typ[0].get<uint16_t>( N );
    h.colors16       = hasAttr ? g.col[0].get<uint16_t>( N * 3 ) : nullptr;
    h.partition      = g.part.get<uint32_t>( N );
    h.point_to_pixel = g.p2p.get<uint32_t>( N * 3 );
    RB( rb200_download_gof( g.ctx, &h ) );
  }
  g.context       = &context;
  g.frames        = F;
  g.gp            = gp;
  g.bDecoder      = bDecoder;
  g.reconstructed = true;
  g.geo = g.transfer = g.color = false;
}

void checkFrame( const char* who, size_t f, size_t points ) {
  if ( !g.reconstructed || f >= g.frames || points != (size_t)g.counts[f].total ) {
    std::fprintf( stderr, "rabbit_b200: %s called for a cloud that generatePointCloud did not produce (frame %zu, %zu points)\n", who,
                  f, points );
    std::exit( -1 );
  }
}

}  // namespace

// instrumentation for the tests: the launch / transfer counters of the shim's context (tests assert the GPU ran)
extern "C" int rb200_shim_stats( rb200_launch_stats* out, int reset ) {
  if ( !g.ctx ) {
    *out = rb200_launch_stats{};
    return RB200_OK;
  }
  return rb200_stats_get( g.ctx, out, reset );
}

namespace pcc {

void PCCCodec::generateOccupancyMap( PCCFrameContext& tile, PCCImageOccupancyMap& videoFrame, const size_t occupancyPrecision,
                                     const size_t thresholdLossyOM, const bool enhancedOccupancyMapForDepthFlag ) {
  const size_t f = tile.getFrameIndex();
  if ( f == 0 || g.occOriginal.count( f ) ) {  // a new GOF starts
    g.occOriginal.clear();
    g.rawSeen.clear();
    g.reconstructed = g.geo = g.transfer = g.color = false;
  }
  g.occFresh = true;
  if ( tile.getLeftTopXInFrame() != 0 || tile.getLeftTopYInFrame() != 0 ) { unsupported( "a tile that does not start at (0, 0)" ); }
  g.occOriginal[f]   = videoFrame.getChannel( 0 );  // the reconstruction of the GOF starts from the decoded samples
  g.thresholdLossyOM = thresholdLossyOM;
  // the caller-visible results of the original — tile.getOccupancyMap() and the video frame thresholded in place
  // (:1597-1600) — computed on the GPU for this frame
  ensureContext();
  const size_t width = tile.getWidth(), height = tile.getHeight();
  auto&        occupancyMap = tile.getOccupancyMap();
  occupancyMap.resize( width * height, 0 );
  if ( width * height == 0 ) { return; }
  if ( width % occupancyPrecision || height % occupancyPrecision || videoFrame.getWidth() != width / occupancyPrecision ||
       videoFrame.getHeight() != height / occupancyPrecision ) {
    unsupported( "an occupancy video that is not exactly tile size / occupancy precision" );
  }
  RB( rb200_occupancy_map( g.ctx, videoFrame.getChannel( 0 ).data(), (int)( width / occupancyPrecision ),
                           (int)( height / occupancyPrecision ), (int)occupancyPrecision, (int)thresholdLossyOM,
                           enhancedOccupancyMapForDepthFlag ? 1 : 0, occupancyMap.data() ) );
}

void PCCCodec::generateBlockToPatchFromOccupancyMapVideo( PCCContext& context, PCCFrameContext& tile, size_t frameIdx,
                                                          PCCImageOccupancyMap& occupancyMapImage, const size_t occupancyResolution,
                                                          const size_t occupancyPrecision ) {
  // tile.getBlockToPatch() is filled by generatePointCloud below (the GPU builds it together with the reconstruction);
  // size it now, as the original does, so that callers that only look at its size keep working
  const size_t bw = context[frameIdx].getAtlasFrameWidth() / occupancyResolution;
  const size_t bh = context[frameIdx].getAtlasFrameHeight() / occupancyResolution;
  tile.getBlockToPatch().assign( bw * bh, 0 );
  if ( !g.occFresh ) {  // generateOccupancyMap was skipped for this frame (occupancy synthesis, PCCDecoder.cpp:362-366)
    if ( frameIdx == 0 || g.rawSeen.count( frameIdx ) ) {  // a new GOF starts
      g.occOriginal.clear();
      g.rawSeen.clear();
      g.reconstructed = g.geo = g.transfer = g.color = false;
    }
    g.rawSeen.insert( frameIdx );
  }
  g.occFresh = false;
  (void)occupancyMapImage;
  (void)occupancyPrecision;
}

void PCCCodec::generatePointCloud( PCCPointSet3& reconstruct, PCCContext& context, size_t frameIndex, size_t tileIndex,
                                   const GeneratePointCloudParameters& params, std::vector<uint32_t>& partition, bool bDecoder ) {
  auto& tile = context[frameIndex].getTile( tileIndex );
  if ( !g.reconstructed ) {
    checkSupported( context, params, tileIndex );
    // the second attribute map is a delta on the first when the stream says so (PCCDecoder.cpp:309-323); colorPointCloud
    // receives the list itself and repeats the reconstruction should it disagree
    bool relT1 = false;
    if ( params.multipleStreams_ && params.mapCountMinus1_ >= 1 ) {
      auto& sps = context.getVps();
      auto& ai  = sps.getAttributeInformation( 0 );
      if ( ai.getAttributeCount() > 0 && ai.getAttributeMapAbsoluteCodingPersistenceFlag( 0 ) == 0u ) {
        relT1 = !sps.getMapAbsoluteCodingEnableFlag( context.getAtlasIndex(), 1 );
      }
    }
    reconstructGof( context, params, bDecoder, relT1 );
  }
  if ( tileIndex != 0 || frameIndex >= g.frames ) { unsupported( "an atlas frame with several tiles (PCCDecoder.cpp:356-381)" ); }
  g.currentFrame                = frameIndex;
  const rb200_frame_counts& cnt = g.counts[frameIndex];
  const size_t              n = (size_t)cnt.total, o = g.off[frameIndex];
  reconstruct.resize( n );
  partition.resize( n );
  auto& pointToPixel = tile.getPointToPixel();
  pointToPixel.resize( n );
  if ( n ) {
    std::memcpy( reconstruct.getPositions().data(), static_cast<int16_t*>( g.pos[0].p ) + 3 * o, n * 6 );
    std::memcpy( reconstruct.getBoundaryPointTypes().data(), static_cast<uint16_t*>( g.typ[0].p ) + o, n * 2 );
    std::memcpy( partition.data(), static_cast<uint32_t*>( g.part.p ) + o, n * 4 );
    const uint32_t* p2p = static_cast<uint32_t*>( g.p2p.p ) + 3 * o;
    for ( size_t i = 0; i < n; i++ ) {
      pointToPixel[i] = PCCVector3<size_t>( p2p[3 * i], p2p[3 * i + 1], p2p[3 * i + 2] );
      reconstruct.setPointPatchIndex( i, (uint32_t)tileIndex, partition[i] );
    }
  }
  tile.setTotalNumberOfRegularPoints( (size_t)cnt.regular );
  tile.setTotalNumberOfEOMPoints( (size_t)cnt.eom );
  tile.setTotalNumberOfRawPoints( (size_t)cnt.raw );
  const size_t          W = g.P.width, H = g.P.height, R = g.P.occupancy_resolution;
  std::vector<uint32_t> b2p( ( W / R ) * ( H / R ) );
  RB( rb200_download_block_to_patch( g.ctx, (int)frameIndex, b2p.data() ) );
  tile.getBlockToPatch().assign( b2p.begin(), b2p.end() );
  if ( !params.pbfEnableFlag_ ) {  // (occupancy synthesis keeps its maps inside the patches; the tile's map stays as it was)
    std::vector<uint8_t> om( W * H );
    RB( rb200_download_occupancy( g.ctx, (int)frameIndex, om.data() ) );
    tile.getOccupancyMap().assign( om.begin(), om.end() );
  }
}

size_t PCCCodec::colorPointCloud( PCCPointSet3& reconstruct, PCCContext& context, PCCFrameContext& tile,
                                  const std::vector<bool>& absoluteT1List, const size_t multipleStreams, const uint8_t attributeCount,
                                  size_t accTilePointCount, const GeneratePointCloudParameters& params ) {
  const size_t f = tile.getFrameIndex();
  if ( accTilePointCount != 0 ) { unsupported( "an atlas frame with several tiles (PCCDecoder.cpp:356-381)" ); }
  checkFrame( "colorPointCloud", f, reconstruct.getPointCount() );
  if ( ( multipleStreams != 0 ) != ( g.P.multiple_streams != 0 ) ) { unsupported( "colorPointCloud with a stream layout other than generatePointCloud's" ); }
  const size_t n = reconstruct.getPointCount();
  if ( n == 0 ) { return accTilePointCount; }
  reconstruct.fillColor();  // :1319
  if ( attributeCount == 0 ) {
    for ( auto& color : reconstruct.getColors() ) { color[0] = color[1] = color[2] = 127; }  // :1327-1330
  } else {
    // the colours were gathered by the reprojection kernel; a delta-coded second map (:1387-1416) is part of that kernel,
    // so the reconstruction is repeated when the caller's list differs from what the stream announced
    const bool relT1 = multipleStreams && g.P.map_count_minus1 >= 1 && absoluteT1List.size() > 1 && !absoluteT1List[1];
    if ( relT1 != ( g.P.relative_t1 != 0 ) ) {
      if ( g.geo || g.transfer || g.color ) { unsupported( "a change of absoluteT1List after post-processing has started" ); }
      reconstructGof( *g.context, g.gp, g.bDecoder, relT1 );
      checkFrame( "colorPointCloud", f, n );
    }
    std::memcpy( reconstruct.getColors16bit().data(), static_cast<uint16_t*>( g.col[0].p ) + 3 * g.off[f], n * 6 );
  }
  (void)context;
  (void)params;
  return accTilePointCount + tile.getTotalNumberOfRegularPoints() + tile.getTotalNumberOfEOMPoints() +
         tile.getTotalNumberOfRawPoints();
}

void PCCCodec::smoothPointCloudPostprocess( PCCPointSet3& reconstruct, const PCCColorTransform colorTransform,
                                            const GeneratePointCloudParameters& params, std::vector<uint32_t>& partition ) {
  const size_t f = g.currentFrame;
  if ( reconstruct.getPointCount() == 0 ) { return; }  // :64
  checkFrame( "smoothPointCloudPostprocess", f, reconstruct.getPointCount() );
  // (gridSmoothing_ == 0: the non-grid smoothPointCloud, :1106-1157 — its debug colouring of the moved points, :1143, is not
  // reproduced: every caller overwrites the 8-bit colours afterwards)
  if ( !g.geo ) {
    RB( rb200_smooth_geometry( g.ctx ) );
    const size_t N = g.off[g.frames];
    rb200_cloud_host h{};
    h.positions      = g.pos[1].get<int16_t>( N * 3 );
    h.boundary_types = g.typ[1].get<uint16_t>( N );
    RB( rb200_download_gof( g.ctx, &h ) );
    RB( rb200_frame_counts_get( g.ctx, g.counts.data() ) );
    g.geo = true;
  }
  const size_t n = reconstruct.getPointCount(), o = g.off[f];
  std::memcpy( reconstruct.getPositions().data(), static_cast<int16_t*>( g.pos[1].p ) + 3 * o, n * 6 );
  std::memcpy( reconstruct.getBoundaryPointTypes().data(), static_cast<uint16_t*>( g.typ[1].p ) + o, n * 2 );
  (void)colorTransform;
  (void)partition;
}

bool PCCPointSet3::transferColors16bitBP( PCCPointSet3& target, const int filterType, const int32_t searchRange,
                                          const bool losslessAttribute, const int numNeighborsColorTransferFwd,
                                          const int numNeighborsColorTransferBwd, const bool useDistWeightedAverageFwd,
                                          const bool useDistWeightedAverageBwd, const bool skipAvgIfIdenticalSourcePointPresentFwd,
                                          const bool skipAvgIfIdenticalSourcePointPresentBwd, const double distOffsetFwd,
                                          const double distOffsetBwd, double maxGeometryDist2Fwd, double maxGeometryDist2Bwd,
                                          double maxColorDist2Fwd, double maxColorDist2Bwd, const bool excludeColorOutlier,
                                          const double thresholdColorOutlierDist ) const {
  const size_t f = g.currentFrame;
  // exactly the decoder's call (PCCDecoder.cpp:447-465); the encoder's other argument sets are not on this path
  const bool decoderCall = filterType == 1 && searchRange == 0 && numNeighborsColorTransferFwd == 8 &&
                           numNeighborsColorTransferBwd == 1 && useDistWeightedAverageFwd && useDistWeightedAverageBwd &&
                           skipAvgIfIdenticalSourcePointPresentFwd && !skipAvgIfIdenticalSourcePointPresentBwd &&
                           distOffsetFwd == 4 && distOffsetBwd == 4 && maxGeometryDist2Fwd >= 512 && maxGeometryDist2Bwd >= 512 &&
                           maxColorDist2Fwd >= 131072 && maxColorDist2Bwd >= 131072 && !excludeColorOutlier;
  if ( !decoderCall ) { unsupported( "transferColors16bitBP with arguments other than the decoder's (PCCDecoder.cpp:447-465)" ); }
  if ( getPointCount() == 0 || !hasColors() ) { return false; }  // :1147
  checkFrame( "transferColors16bitBP", f, target.getPointCount() );
  if ( getPointCount() != target.getPointCount() || losslessAttribute != ( g.P.attribute_rgb444 != 0 ) ) {
    unsupported( "transferColors16bitBP between clouds other than a decoded frame and its smoothed copy" );
  }
  (void)thresholdColorOutlierDist;
  target.addColors16bit();
  const size_t n = target.getPointCount(), o = g.off[f];
  if ( !g.geo ) {  // gridSmoothing_ == 0: both clouds are the reconstruction, no point is of type 3, nothing changes (:1163-1164)
    std::memcpy( target.getColors16bit().data(), static_cast<uint16_t*>( g.col[0].p ) + 3 * o, n * 6 );
    return true;
  }
  if ( !g.transfer ) {
    RB( rb200_transfer_colors( g.ctx ) );
    rb200_cloud_host h{};
    h.colors16 = g.col[1].get<uint16_t>( g.off[g.frames] * 3 );
    RB( rb200_download_gof( g.ctx, &h ) );
    g.transfer = true;
  }
  std::memcpy( target.getColors16bit().data(), static_cast<uint16_t*>( g.col[1].p ) + 3 * o, n * 6 );
  return true;
}

void PCCCodec::colorSmoothing( PCCPointSet3& reconstruct, const PCCColorTransform colorTransform,
                               const GeneratePointCloudParameters& params ) {
  const size_t f = g.currentFrame;
  if ( reconstruct.getPointCount() == 0 ) { return; }
  checkFrame( "colorSmoothing", f, reconstruct.getPointCount() );
  if ( !g.color ) {
    RB( rb200_smooth_color( g.ctx ) );
    rb200_cloud_host h{};
    h.colors16 = g.col[2].get<uint16_t>( g.off[g.frames] * 3 );
    RB( rb200_download_gof( g.ctx, &h ) );
    RB( rb200_frame_counts_get( g.ctx, g.counts.data() ) );
    g.color = true;
  }
  std::memcpy( reconstruct.getColors16bit().data(), static_cast<uint16_t*>( g.col[2].p ) + 3 * g.off[f], reconstruct.getPointCount() * 6 );
  (void)colorTransform;
  (void)params;
}

namespace {
rb200_metrics_params metricsParams( const PCCMetricsParameters& p ) {
  if ( p.computeLidar_ || p.computeReflectance_ ) { unsupported( "lidar / reflectance metrics" ); }
  rb200_metrics_params mp{};
  mp.compute_c2c = p.computeC2c_, mp.compute_c2p = p.computeC2p_, mp.compute_color = p.computeColor_;
  mp.compute_hausdorff = p.computeHausdorff_, mp.drop_duplicates = (int)p.dropDuplicates_;
  mp.neighbors_proc = (int)p.neighborsProc_, mp.resolution = (float)p.resolution_;
  return mp;
}
QualityMetrics fillQuality( const PCCMetricsParameters& params, const rb200_quality& s ) {
  QualityMetrics q;
  q.setParameters( params );
  q.psnr_    = params.resolution_;
  q.c2cMse_  = s.c2c_mse, q.c2cPsnr_ = s.c2c_psnr, q.c2cHausdorff_ = s.c2c_hausdorff, q.c2cHausdorffPsnr_ = s.c2c_hausdorff_psnr;
  q.c2pMse_  = s.c2p_mse, q.c2pPsnr_ = s.c2p_psnr, q.c2pHausdorff_ = s.c2p_hausdorff, q.c2pHausdorffPsnr_ = s.c2p_hausdorff_psnr;
  for ( int k = 0; k < 3; k++ ) { q.colorMse_[k] = s.color_mse[k], q.colorPsnr_[k] = s.color_psnr[k]; }
  return q;
}
// the C ABI takes the normals through the source view: the normal cloud must hold the source's points in the source's order
bool sameOrder( const PCCPointSet3& normals, const PCCPointSet3& source ) {
  return normals.getPointCount() == source.getPointCount() &&
         std::memcmp( normals.positions_.data(), source.positions_.data(), source.getPointCount() * 6 ) == 0;
}
std::vector<float> floatNormals( const PCCPointSet3& normals ) {
  std::vector<float> out( normals.getPointCount() * 3 );
  for ( size_t k = 0; k < normals.getPointCount(); k++ ) {
    for ( int c = 0; c < 3; c++ ) { out[3 * k + c] = (float)normals.normals_[k][c]; }
  }
  return out;
}
rb200_cloud_view viewOf( const PCCPointSet3& pc ) {
  return rb200_cloud_view{reinterpret_cast<const int16_t*>( pc.positions_.data() ),
                          pc.hasColors() ? reinterpret_cast<const uint8_t*>( pc.colors_.data() ) : nullptr, nullptr,
                          (int64_t)pc.getPointCount()};
}
}  // namespace

void PCCMetrics::compute( const PCCGroupOfFrames& sources, const PCCGroupOfFrames& reconstructs, const PCCGroupOfFrames& normals ) {
  if ( normals.getFrameCount() != 0 && sources.getFrameCount() != normals.getFrameCount() ) { params_.computeC2p_ = false; }  // :338-340
  if ( sources.getFrameCount() != reconstructs.getFrameCount() ) {  // :341-347
    printf( "Error: group of frames must have same numbers of frames. ( src = %zu rec = %zu norm = %zu ) \n", sources.getFrameCount(),
            reconstructs.getFrameCount(), normals.getFrameCount() );
    exit( -1 );
  }
  const size_t n = sources.getFrameCount();
  if ( n == 0 ) { return; }
  const bool                      useNormals = normals.getFrameCount() == n;
  rb200_metrics_params            mp         = metricsParams( params_ );
  std::vector<rb200_cloud_view>   vs( n ), vr( n );
  std::vector<std::vector<float>> nrm( n );
  for ( size_t i = 0; i < n; i++ ) {
    if ( sources[i].getPointCount() == 0 || reconstructs[i].getPointCount() == 0 ) { unsupported( "metrics of an empty cloud" ); }
    vs[i] = viewOf( sources[i] );
    vr[i] = viewOf( reconstructs[i] );
    if ( useNormals && normals[i].getPointCount() > 0 ) {
      if ( !sameOrder( normals[i], sources[i] ) ) { unsupported( "a normal cloud that is not the source cloud point for point" ); }
      nrm[i]        = floatNormals( normals[i] );
      vs[i].normals = nrm[i].data();
    }
  }
  ensureContext();
  std::vector<rb200_metrics_result> res( n );
  const int                         st = rb200_metrics( g.ctx, &mp, (int)n, vs.data(), vr.data(), res.data() );
  if ( st != RB200_OK && st != RB200_ERR_TIE_OVERFLOW ) { die( st ); }
  for ( size_t i = 0; i < n; i++ ) {
    sourcePoints_.push_back( (size_t)res[i].source_points );
    reconstructPoints_.push_back( (size_t)res[i].rec_points );
    sourceDuplicates_.push_back( (size_t)res[i].source_after_dedup );
    reconstructDuplicates_.push_back( (size_t)res[i].rec_after_dedup );
    quality1_.push_back( fillQuality( params_, res[i].q1 ) );
    quality2_.push_back( fillQuality( params_, res[i].q2 ) );
    qualityF_.push_back( fillQuality( params_, res[i].qf ) );
  }
}

// the per-pair overload (:371-385): the clouds are compared as they are (the caller removed the duplicates, :358-361).
// The reference also leaves the transferred normals in `source` / `reconstruct` (copyNormals / scaleNormals); no caller
// of the path reads them afterwards, and they are not produced here.
void PCCMetrics::compute( PCCPointSet3& source, PCCPointSet3& reconstruct, const PCCPointSet3& normalSource ) {
  rb200_metrics_params mp = metricsParams( params_ );
  mp.drop_duplicates      = 0;
  if ( source.getPointCount() == 0 || reconstruct.getPointCount() == 0 ) { unsupported( "metrics of an empty cloud" ); }
  rb200_cloud_view   vs = viewOf( source ), vr = viewOf( reconstruct );
  std::vector<float> nrm;
  if ( normalSource.getPointCount() > 0 ) {
    if ( !sameOrder( normalSource, source ) ) { unsupported( "a normal cloud that is not the source cloud point for point" ); }
    nrm        = floatNormals( normalSource );
    vs.normals = nrm.data();
  }
  ensureContext();
  rb200_metrics_result res{};
  const int            st = rb200_metrics( g.ctx, &mp, 1, &vs, &vr, &res );
  if ( st != RB200_OK && st != RB200_ERR_TIE_OVERFLOW ) { die( st ); }
  quality1_.push_back( fillQuality( params_, res.q1 ) );
  quality2_.push_back( fillQuality( params_, res.q2 ) );
  qualityF_.push_back( fillQuality( params_, res.qf ) );
}

}  // namespace pcc
