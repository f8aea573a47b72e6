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
//   PCCMetrics::compute( sources, reconstructs, normals )  PccLibMetrics/source/PCCMetrics.cpp:334-385
// The originals stay linked under the names rb200_orig_* (oracle/Makefile renames the symbols with objcopy), and every
// replaced body falls back to its original whenever the request is outside what the CUDA path implements (multiple
// tiles, auxiliary video, multiple streams, PBF, other transfer-filter arguments ...):
// an unsupported mode therefore gives the reference's result, never a different one.
//
// Per-frame calls, per-GOF execution: the reference calls these functions frame by frame; the CUDA path processes the
// whole GOF in one batched launch sequence.  The first call of a stage for a GOF runs that stage for every frame on
// the GPU, every call then copies the frame it was asked for into the caller's containers.
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

using namespace pcc;

// the reference's own bodies (renamed by objcopy; Itanium ABI: `this` is the first argument)
extern "C" {
void   rb200_orig_generateOccupancyMap( PCCCodec*, PCCFrameContext&, PCCImageOccupancyMap&, size_t, size_t, bool );
void   rb200_orig_generateBlockToPatch( PCCCodec*, PCCContext&, PCCFrameContext&, size_t, PCCImageOccupancyMap&, size_t, size_t );
void   rb200_orig_generatePointCloud( PCCCodec*, PCCPointSet3&, PCCContext&, size_t, size_t, const GeneratePointCloudParameters&,
                                      std::vector<uint32_t>&, bool );
size_t rb200_orig_colorPointCloud( PCCCodec*, PCCPointSet3&, PCCContext&, PCCFrameContext&, const std::vector<bool>&, size_t,
                                   uint8_t, size_t, const GeneratePointCloudParameters& );
void   rb200_orig_smoothPointCloudPostprocess( PCCCodec*, PCCPointSet3&, PCCColorTransform, const GeneratePointCloudParameters&,
                                               std::vector<uint32_t>& );
void   rb200_orig_colorSmoothing( PCCCodec*, PCCPointSet3&, PCCColorTransform, const GeneratePointCloudParameters& );
bool   rb200_orig_transferColors16bitBP( const PCCPointSet3*, PCCPointSet3&, int, int32_t, bool, int, int, bool, bool, bool, bool,
                                         double, double, double, double, double, double, bool, double );
void   rb200_orig_metricsCompute( PCCMetrics*, const PCCGroupOfFrames&, const PCCGroupOfFrames&, const PCCGroupOfFrames& );
}

namespace {

struct Gof {  // the GOF currently resident on the GPU
  rb200_ctx*                      ctx     = nullptr;
  const PCCContext*               context = nullptr;
  size_t                          frames  = 0;
  rb200_params                    P{};
  bool                            reconstructed = false, geo = false, transfer = false, color = false;
  std::vector<rb200_frame_counts> counts;
  std::map<size_t, std::vector<uint8_t>> occOriginal;  // occupancy video of the frames seen by generateOccupancyMap
  size_t                          thresholdLossyOM = 0;
  bool                            eom              = false;
  size_t                          currentFrame     = 0;
  bool                            active           = false;  // false: the CUDA path declined this GOF (fallback)
};
Gof g;  // the decoder's frame loop is single-threaded (PCCDecoder.cpp:330) and PCCCodec is not re-entrant per instance

[[noreturn]] void die( int status ) {
  // the reference prints and exits (PCCPatch.cpp:237-245 exits 180; everything else -1)
  std::fprintf( stderr, "rabbit_b200: %s\n", rb200_error_string( g.ctx ) );
  std::exit( status == RB200_ERR_PATCH_OUT_OF_CANVAS ? 180 : -1 );
}
#define RB( call )                             \
  do {                                         \
    int st__ = ( call );                       \
    if ( st__ != RB200_OK ) { die( st__ ); }   \
  } while ( 0 )

bool supported( PCCContext& context, const GeneratePointCloudParameters& p ) {
  if ( p.pbfEnableFlag_ || p.useAuxSeperateVideo_ || p.multipleStreams_ || p.mapCountMinus1_ > 1 || p.occupancyResolution_ != 16 ) {
    return false;
  }
  if ( ( p.pointLocalReconstruction_ || p.singleMapPixelInterleaving_ ) &&
       ( p.mapCountMinus1_ != 0 || p.enhancedOccupancyMapCode_ || p.surfaceThickness_ < 1 ) ) {
    return false;  // combinations the C ABI refuses (rb200_gof_begin)
  }
  for ( size_t f = 0; f < context.size(); f++ ) {
    if ( context[f].getNumTilesInAtlasFrame() != 1 ) { return false; }
    auto& tile = context[f].getTile( 0 );
    if ( tile.getLeftTopXInFrame() != 0 || tile.getLeftTopYInFrame() != 0 || tile.getUseRawPointsSeparateVideo() ) { return false; }
  }
  return true;
}

// PCCContext -> flat structs of the C ABI, upload, reconstruction of every frame
void reconstructGof( PCCContext& context, const GeneratePointCloudParameters& gp, bool bDecoder ) {
  const size_t F = context.size();
  const size_t W = context[0].getAtlasFrameWidth(), H = context[0].getAtlasFrameHeight();
  const size_t M = gp.mapCountMinus1_ + 1, P = gp.occupancyPrecision_;
  const size_t oW = W / P, oH = H / P;
  auto&        ai       = context.getVps().getAttributeInformation( 0 );
  const bool   hasAttr  = ai.getAttributeCount() > 0;
  auto&        geoVideo = context.getVideoGeometryMultiple()[0];
  rb200_params& p = g.P;
  std::memset( &p, 0, sizeof( p ) );
  p.width = (int)W, p.height = (int)H;
  p.occupancy_resolution        = (int)gp.occupancyResolution_;
  p.occupancy_precision         = (int)P;
  p.threshold_lossy_om          = (int)g.thresholdLossyOM;
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
  p.attribute_count             = hasAttr ? 1 : 0;
  p.attribute_rgb444 = hasAttr && context.getVideoAttributesMultiple()[0].getFrameCount() > 0 &&
                       context.getVideoAttributesMultiple()[0].getColorFormat() == PCCCOLORFORMAT::RGB444;
  p.geometry_bitdepth_3d       = (int)gp.geometryBitDepth3D_;
  p.flag_geometry_smoothing    = gp.flagGeometrySmoothing_;
  p.grid_smoothing             = gp.gridSmoothing_;
  p.grid_size                  = (int)gp.gridSize_;
  p.apply_geo_smoothing        = 1;  // the caller decides which stage functions it calls
  p.attr_transfer_filter_type  = 1;  // keep the pre-smoothing cloud: transferColors16bitBP may follow
  p.flag_color_smoothing       = gp.flagColorSmoothing_;
  p.apply_attr_smoothing       = 1;
  p.threshold_smoothing        = gp.thresholdSmoothing_;
  p.threshold_color_smoothing  = gp.thresholdColorSmoothing_;
  p.threshold_color_difference = gp.thresholdColorDifference_;
  p.threshold_color_variation  = gp.thresholdColorVariation_;

  std::vector<uint8_t>  occ( F * oW * oH );
  std::vector<uint16_t> geo( F * M * W * H ), att( hasAttr ? F * M * 3 * W * H : 0 );
  auto&                 occVideo = context.getVideoOccupancyMap();
  for ( size_t f = 0; f < F; f++ ) {
    auto it = g.occOriginal.find( f );  // a frame generateOccupancyMap already binarised in place: use the saved copy
    if ( it != g.occOriginal.end() ) {
      std::memcpy( &occ[f * oW * oH], it->second.data(), oW * oH );
    } else {
      std::memcpy( &occ[f * oW * oH], occVideo.getFrame( f ).getChannel( 0 ).data(), oW * oH );
    }
    for ( size_t m = 0; m < M; m++ ) {
      std::memcpy( &geo[( f * M + m ) * W * H], geoVideo.getFrame( f * M + m ).getChannel( 0 ).data(), W * H * 2 );
      if ( hasAttr ) {
        auto& a = context.getVideoAttributesMultiple()[0].getFrame( f * M + m );
        for ( int c = 0; c < 3; c++ ) { std::memcpy( &att[( ( f * M + m ) * 3 + c ) * W * H], a.getChannel( c ).data(), W * H * 2 ); }
      }
    }
  }
  std::vector<rb200_patch>     patches;
  std::vector<rb200_eom_patch> eoms;
  std::vector<rb200_raw_patch> raws;
  std::vector<int32_t>         pOff{0}, eOff{0}, rOff{0}, members;
  for ( size_t f = 0; f < F; f++ ) {
    auto& tile = context[f].getTile( 0 );
    for ( auto& s : tile.getPatches() ) {
      rb200_patch d{};
      d.u0 = (int)s.getU0(), d.v0 = (int)s.getV0(), d.size_u0 = (int)s.getSizeU0(), d.size_v0 = (int)s.getSizeV0();
      d.u1 = (int)s.getU1(), d.v1 = (int)s.getV1(), d.d1 = (int)s.getD1();
      d.normal_axis = (int)s.getNormalAxis(), d.tangent_axis = (int)s.getTangentAxis(), d.bitangent_axis = (int)s.getBitangentAxis();
      d.projection_mode = (int)s.getProjectionMode(), d.orientation = (int)s.getPatchOrientation();
      d.lod_x = (int)s.getLodScaleX(), d.lod_y = (int)s.getLodScaleY();
      d.axis_of_additional_plane = (int)s.getAxisOfAdditionalPlane();
      d.size2d_x_px = (int)s.getPatchSize2DXInPixel(), d.size2d_y_px = (int)s.getPatchSize2DYInPixel();
      patches.push_back( d );
    }
    pOff.push_back( (int32_t)patches.size() );
    for ( auto& s : tile.getEomPatches() ) {
      rb200_eom_patch e{};
      e.u0 = (int)s.u0_, e.v0 = (int)s.v0_, e.member_begin = (int)members.size(), e.member_count = (int)s.memberPatches_.size();
      e.eom_count = (int)s.eomCount_;
      for ( auto m : s.memberPatches_ ) { members.push_back( (int32_t)m ); }
      eoms.push_back( e );
    }
    eOff.push_back( (int32_t)eoms.size() );
    for ( auto& s : tile.getRawPointsPatches() ) {
      rb200_raw_patch r{};
      r.u0 = (int)s.u0_, r.v0 = (int)s.v0_, r.size_u0 = (int)s.sizeU0_, r.size_v0 = (int)s.sizeV0_;
      r.u1 = (int)s.u1_, r.v1 = (int)s.v1_, r.d1 = (int)s.d1_, r.num_points = (int)s.getNumberOfRawPoints();
      raws.push_back( r );
    }
    rOff.push_back( (int32_t)raws.size() );
  }
  if ( members.empty() ) { members.push_back( 0 ); }
  rb200_frames fr{occ.data(), geo.data(), hasAttr ? att.data() : nullptr};
  rb200_atlas  at{patches.data(), pOff.data(), eoms.empty() ? nullptr : eoms.data(), eOff.data(), members.data(),
                  raws.empty() ? nullptr : raws.data(), rOff.data()};
  if ( !g.ctx ) { RB( rb200_create( 0, &g.ctx ) ); }
  RB( rb200_enable_stage_snapshots( g.ctx, 1 ) );  // the caller walks the stages frame by frame
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
  g.context       = &context;
  g.frames        = F;
  g.reconstructed = true;
  g.geo = g.transfer = g.color = false;
}

}  // namespace

namespace pcc {

void PCCCodec::generateOccupancyMap( PCCFrameContext& tile, PCCImageOccupancyMap& videoFrame, const size_t occupancyPrecision,
                                     const size_t thresholdLossyOM, const bool enhancedOccupancyMapForDepthFlag ) {
  const size_t f = tile.getFrameIndex();
  if ( f == 0 || g.occOriginal.count( f ) ) {  // a new GOF starts
    g.occOriginal.clear();
    g.reconstructed = g.geo = g.transfer = g.color = false;
    g.active                                       = false;
  }
  g.occOriginal[f]   = videoFrame.getChannel( 0 );  // the original body thresholds the video sample in place (:1597-1600)
  g.thresholdLossyOM = thresholdLossyOM;
  g.eom              = enhancedOccupancyMapForDepthFlag;
  // keep the caller-visible side effects of the original (binarised video frame, tile.getOccupancyMap()) by running it;
  // the map used for the reconstruction is recomputed on the GPU from the saved samples
  rb200_orig_generateOccupancyMap( this, tile, videoFrame, occupancyPrecision, thresholdLossyOM, enhancedOccupancyMapForDepthFlag );
}

void PCCCodec::generateBlockToPatchFromOccupancyMapVideo( PCCContext& context, PCCFrameContext& tile, size_t frameIdx,
                                                          PCCImageOccupancyMap& occupancyMapImage, const size_t occupancyResolution,
                                                          const size_t occupancyPrecision ) {
  // tile.getBlockToPatch() is filled by generatePointCloud below (the GPU builds it together with the reconstruction);
  // size it now, as the original does, so that callers that only look at its size keep working
  const size_t bw = context[frameIdx].getAtlasFrameWidth() / occupancyResolution;
  const size_t bh = context[frameIdx].getAtlasFrameHeight() / occupancyResolution;
  tile.getBlockToPatch().assign( bw * bh, 0 );
  (void)occupancyMapImage;
  (void)occupancyPrecision;
}

void PCCCodec::generatePointCloud( PCCPointSet3& reconstruct, PCCContext& context, size_t frameIndex, size_t tileIndex,
                                   const GeneratePointCloudParameters& params, std::vector<uint32_t>& partition, bool bDecoder ) {
  if ( !g.reconstructed ) { g.active = supported( context, params ) && tileIndex == 0; }
  auto& tile = context[frameIndex].getTile( tileIndex );
  if ( !g.active ) {
    rb200_orig_generateBlockToPatch( this, context, tile, frameIndex, context.getVideoOccupancyMap().getFrame( frameIndex ),
                                     params.occupancyResolution_, params.occupancyPrecision_ );
    rb200_orig_generatePointCloud( this, reconstruct, context, frameIndex, tileIndex, params, partition, bDecoder );
    return;
  }
  if ( !g.reconstructed ) { reconstructGof( context, params, bDecoder ); }
  g.currentFrame                = frameIndex;
  const rb200_frame_counts& cnt = g.counts[frameIndex];
  const size_t              n   = (size_t)cnt.total;
  reconstruct.resize( n );
  partition.resize( n );
  std::vector<uint32_t> p2p( n * 3 );
  rb200_cloud_host      h{};
  h.positions      = n ? reinterpret_cast<int16_t*>( reconstruct.getPositions().data() ) : nullptr;
  h.boundary_types = n ? reconstruct.getBoundaryPointTypes().data() : nullptr;
  h.partition      = n ? partition.data() : nullptr;
  h.point_to_pixel = n ? p2p.data() : nullptr;
  RB( rb200_download_frame_stage( g.ctx, (int)frameIndex, 0, &h ) );
  auto& pointToPixel = tile.getPointToPixel();
  pointToPixel.resize( n );
  for ( size_t i = 0; i < n; i++ ) {
    pointToPixel[i] = PCCVector3<size_t>( p2p[3 * i], p2p[3 * i + 1], p2p[3 * i + 2] );
    reconstruct.setPointPatchIndex( i, (uint32_t)tileIndex, partition[i] );
  }
  tile.setTotalNumberOfRegularPoints( (size_t)cnt.regular );
  tile.setTotalNumberOfEOMPoints( (size_t)cnt.eom );
  tile.setTotalNumberOfRawPoints( (size_t)cnt.raw );
  const size_t          W = g.P.width, H = g.P.height, R = g.P.occupancy_resolution;
  std::vector<uint32_t> b2p( ( W / R ) * ( H / R ) );
  RB( rb200_download_block_to_patch( g.ctx, (int)frameIndex, b2p.data() ) );
  tile.getBlockToPatch().assign( b2p.begin(), b2p.end() );
  std::vector<uint8_t> om( W * H );
  RB( rb200_download_occupancy( g.ctx, (int)frameIndex, om.data() ) );
  tile.getOccupancyMap().assign( om.begin(), om.end() );
}

size_t PCCCodec::colorPointCloud( PCCPointSet3& reconstruct, PCCContext& context, PCCFrameContext& tile,
                                  const std::vector<bool>& absoluteT1List, const size_t multipleStreams, const uint8_t attributeCount,
                                  size_t accTilePointCount, const GeneratePointCloudParameters& params ) {
  const size_t f = tile.getFrameIndex();
  if ( !g.active || !g.reconstructed || multipleStreams || accTilePointCount != 0 ||
       reconstruct.getPointCount() != (size_t)g.counts[f].total ) {
    return rb200_orig_colorPointCloud( this, reconstruct, context, tile, absoluteT1List, multipleStreams, attributeCount,
                                       accTilePointCount, params );
  }
  const size_t n = reconstruct.getPointCount();
  if ( n == 0 ) { return accTilePointCount; }
  reconstruct.fillColor();  // :1319
  if ( attributeCount == 0 ) {
    for ( auto& color : reconstruct.getColors() ) { color[0] = color[1] = color[2] = 127; }  // :1327-1330
  } else {
    rb200_cloud_host h{};
    h.colors16 = reinterpret_cast<uint16_t*>( reconstruct.getColors16bit().data() );  // gathered by the reprojection kernel
    RB( rb200_download_frame_stage( g.ctx, (int)f, 0, &h ) );
  }
  return accTilePointCount + tile.getTotalNumberOfRegularPoints() + tile.getTotalNumberOfEOMPoints() +
         tile.getTotalNumberOfRawPoints();
}

void PCCCodec::smoothPointCloudPostprocess( PCCPointSet3& reconstruct, const PCCColorTransform colorTransform,
                                            const GeneratePointCloudParameters& params, std::vector<uint32_t>& partition ) {
  const size_t f = g.currentFrame;
  if ( !g.active || !g.reconstructed || reconstruct.getPointCount() != (size_t)g.counts[f].total ) {
    rb200_orig_smoothPointCloudPostprocess( this, reconstruct, colorTransform, params, partition );
    return;
  }
  if ( !g.geo ) {
    RB( rb200_smooth_geometry( g.ctx ) );
    g.geo = true;
  }
  if ( reconstruct.getPointCount() == 0 ) { return; }
  rb200_cloud_host h{};
  h.positions      = reinterpret_cast<int16_t*>( reconstruct.getPositions().data() );
  h.boundary_types = reconstruct.getBoundaryPointTypes().data();
  RB( rb200_download_frame_stage( g.ctx, (int)f, 1, &h ) );
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
  // exactly the decoder's call (PCCDecoder.cpp:447-465); anything else runs the original body
  const bool decoderCall = filterType == 1 && searchRange == 0 && numNeighborsColorTransferFwd == 8 &&
                           numNeighborsColorTransferBwd == 1 && useDistWeightedAverageFwd && useDistWeightedAverageBwd &&
                           skipAvgIfIdenticalSourcePointPresentFwd && !skipAvgIfIdenticalSourcePointPresentBwd &&
                           distOffsetFwd == 4 && distOffsetBwd == 4 && maxGeometryDist2Fwd >= 512 && maxGeometryDist2Bwd >= 512 &&
                           maxColorDist2Fwd >= 131072 && maxColorDist2Bwd >= 131072 && !excludeColorOutlier;
  if ( !decoderCall || !g.active || !g.geo || target.getPointCount() != (size_t)g.counts[f].total ||
       getPointCount() != target.getPointCount() || losslessAttribute != ( g.P.attribute_rgb444 != 0 ) ) {
    return rb200_orig_transferColors16bitBP( this, target, filterType, searchRange, losslessAttribute, numNeighborsColorTransferFwd,
                                             numNeighborsColorTransferBwd, useDistWeightedAverageFwd, useDistWeightedAverageBwd,
                                             skipAvgIfIdenticalSourcePointPresentFwd, skipAvgIfIdenticalSourcePointPresentBwd,
                                             distOffsetFwd, distOffsetBwd, maxGeometryDist2Fwd, maxGeometryDist2Bwd,
                                             maxColorDist2Fwd, maxColorDist2Bwd, excludeColorOutlier, thresholdColorOutlierDist );
  }
  if ( getPointCount() == 0 || !hasColors() ) { return false; }  // :1147
  if ( !g.transfer ) {
    RB( rb200_transfer_colors( g.ctx ) );
    g.transfer = true;
  }
  target.addColors16bit();
  rb200_cloud_host h{};
  h.colors16 = reinterpret_cast<uint16_t*>( target.getColors16bit().data() );
  RB( rb200_download_frame_stage( g.ctx, (int)f, 2, &h ) );
  return true;
}

void PCCCodec::colorSmoothing( PCCPointSet3& reconstruct, const PCCColorTransform colorTransform,
                               const GeneratePointCloudParameters& params ) {
  const size_t f = g.currentFrame;
  if ( !g.active || !g.reconstructed || reconstruct.getPointCount() != (size_t)g.counts[f].total ) {
    rb200_orig_colorSmoothing( this, reconstruct, colorTransform, params );
    return;
  }
  if ( !g.color ) {
    RB( rb200_smooth_color( g.ctx ) );
    g.color = true;
  }
  if ( reconstruct.getPointCount() == 0 ) { return; }
  rb200_cloud_host h{};
  h.colors16 = reinterpret_cast<uint16_t*>( reconstruct.getColors16bit().data() );
  RB( rb200_download_frame_stage( g.ctx, (int)f, 3, &h ) );
}

void PCCMetrics::compute( const PCCGroupOfFrames& sources, const PCCGroupOfFrames& reconstructs, const PCCGroupOfFrames& normals ) {
  const size_t n = sources.getFrameCount();
  bool         ok = n > 0 && n == reconstructs.getFrameCount() && params_.neighborsProc_ >= 1 && params_.neighborsProc_ <= 4 &&
            !params_.computeLidar_ && !params_.computeReflectance_ && ( normals.getFrameCount() == 0 || normals.getFrameCount() == n );
  for ( size_t i = 0; ok && i < n; i++ ) {
    ok = sources[i].getPointCount() > 0 && reconstructs[i].getPointCount() > 0 && sources[i].hasColors() == reconstructs[i].hasColors();
    if ( ok && normals.getFrameCount() ) {
      // the C ABI takes the normal cloud through the source view: same points in the same order
      ok = normals[i].getPointCount() == sources[i].getPointCount() &&
           std::memcmp( normals[i].positions_.data(), sources[i].positions_.data(), sources[i].getPointCount() * 6 ) == 0;
    }
  }
  if ( !ok ) {  // includes the error paths of the original (:337-347), which prints and exits itself
    rb200_orig_metricsCompute( this, sources, reconstructs, normals );
    return;
  }
  rb200_metrics_params mp{};
  mp.compute_c2c = params_.computeC2c_, mp.compute_c2p = params_.computeC2p_, mp.compute_color = params_.computeColor_;
  mp.compute_hausdorff = params_.computeHausdorff_, mp.drop_duplicates = (int)params_.dropDuplicates_;
  mp.neighbors_proc = (int)params_.neighborsProc_, mp.resolution = (float)params_.resolution_;
  std::vector<rb200_cloud_view>   vs( n ), vr( n );
  std::vector<std::vector<float>> nrm( n );
  for ( size_t i = 0; i < n; i++ ) {
    vs[i] = rb200_cloud_view{reinterpret_cast<const int16_t*>( sources[i].positions_.data() ),
                             sources[i].hasColors() ? reinterpret_cast<const uint8_t*>( sources[i].colors_.data() ) : nullptr, nullptr,
                             (int64_t)sources[i].getPointCount()};
    vr[i] = rb200_cloud_view{reinterpret_cast<const int16_t*>( reconstructs[i].positions_.data() ),
                             reconstructs[i].hasColors() ? reinterpret_cast<const uint8_t*>( reconstructs[i].colors_.data() ) : nullptr,
                             nullptr, (int64_t)reconstructs[i].getPointCount()};
    if ( normals.getFrameCount() ) {
      nrm[i].resize( normals[i].getPointCount() * 3 );
      for ( size_t k = 0; k < normals[i].getPointCount(); k++ ) {
        for ( int c = 0; c < 3; c++ ) { nrm[i][3 * k + c] = (float)normals[i].normals_[k][c]; }
      }
      vs[i].normals = nrm[i].data();
    }
  }
  if ( !g.ctx ) { RB( rb200_create( 0, &g.ctx ) ); }
  std::vector<rb200_metrics_result> res( n );
  const int                         st = rb200_metrics( g.ctx, &mp, (int)n, vs.data(), vr.data(), res.data() );
  if ( st != RB200_OK && st != RB200_ERR_TIE_OVERFLOW ) { die( st ); }
  auto fill = [&]( const rb200_quality& s ) {
    QualityMetrics q;
    q.setParameters( params_ );
    q.psnr_    = params_.resolution_;
    q.c2cMse_  = s.c2c_mse, q.c2cPsnr_ = s.c2c_psnr, q.c2cHausdorff_ = s.c2c_hausdorff, q.c2cHausdorffPsnr_ = s.c2c_hausdorff_psnr;
    q.c2pMse_  = s.c2p_mse, q.c2pPsnr_ = s.c2p_psnr, q.c2pHausdorff_ = s.c2p_hausdorff, q.c2pHausdorffPsnr_ = s.c2p_hausdorff_psnr;
    for ( int k = 0; k < 3; k++ ) { q.colorMse_[k] = s.color_mse[k], q.colorPsnr_[k] = s.color_psnr[k]; }
    return q;
  };
  for ( size_t i = 0; i < n; i++ ) {
    sourcePoints_.push_back( (size_t)res[i].source_points );
    reconstructPoints_.push_back( (size_t)res[i].rec_points );
    sourceDuplicates_.push_back( (size_t)res[i].source_after_dedup );
    reconstructDuplicates_.push_back( (size_t)res[i].rec_after_dedup );
    quality1_.push_back( fill( res[i].q1 ) );
    quality2_.push_back( fill( res[i].q2 ) );
    qualityF_.push_back( fill( res[i].qf ) );
  }
}

}  // namespace pcc
