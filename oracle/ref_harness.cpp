// ref_harness.cpp — drives the UNMODIFIED reference implementation (compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/librabbit_ref.so) on flat buffers.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (rabbit-transcoding_b200/) may load this; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// It executes exactly the decoder's per-frame sequence of PccLibDecoder/source/PCCDecoder.cpp:330-508:
//   generateOccupancyMap (:363) -> generateBlockToPatchFromOccupancyMapVideo (:373) ->
//   generatePointCloud (:380) -> appendPointSet (:381) -> colorPointCloud (:391) ->
//   [smoothPointCloudPostprocess (:437) -> transferColors16bitBP (:449)] -> [colorSmoothing (:498)] ->
//   convertYUV16ToRGB8 | copyRGB16ToRGB8 (:503/:506)
// with stage snapshots after each step, and PCCMetrics::compute for the metrics.  Inputs use the structs of
// include/rabbit_b200.h so the CUDA path and the reference see byte-identical data.

// Standard headers first: the access-specifier override below must only touch the reference's headers.
#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <queue>
#include <set>
#include <sstream>
#include <stack>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <condition_variable>
#include <sys/time.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include "nanoflann.hpp"
#include "tbb/tbb.h"

#define private public
#define protected public
#include "PCCCommon.h"
#include "PCCImage.h"
#include "PCCVideo.h"
#include "PCCPointSet.h"
#include "PCCPatch.h"
#include "PCCContext.h"
#include "PCCFrameContext.h"
#include "PCCInternalColorConverter.h"
#include "PCCGroupOfFrames.h"
#include "PCCCodec.h"
#include "PCCKdTree.h"
#include "KDTreeVectorOfVectorsAdaptor.h"
#include "PCCMetricsParameters.h"
#include "PCCMetrics.h"
#undef private
#undef protected

#include "rabbit_b200.h"
#include "../rabbit-transcoding_b200/host/rb200_atlas_export.h"

using namespace pcc;

namespace {

struct StageSnap {
  bool                     valid = false;
  std::vector<int16_t>     pos;     // [N][3]
  std::vector<uint16_t>    col16;   // [N][3]
  std::vector<uint8_t>     col8;    // [N][3]
  std::vector<uint16_t>    btype;   // [N]
};

struct FrameOut {
  StageSnap             stage[5];
  std::vector<uint32_t> partition;
  std::vector<uint32_t> pointToPixel;  // [N][3]
  std::vector<uint32_t> blockToPatch;
  std::vector<uint8_t>  occupancy;
  rb200_frame_counts    counts{};
  uint8_t               md5Ordered[16]{};
  uint8_t               md5Canonical[16]{};
  double                msReconstruct = 0, msPost = 0;
  double                msStage[5]    = {0, 0, 0, 0, 0};
};

struct Oracle : public PCCCodec {};  // reach the protected generateOccupancyMap / generateBlockToPatch...

void snap( StageSnap& s, PCCPointSet3& pc ) {
  const size_t n = pc.getPointCount();
  s.valid        = true;
  s.pos.resize( n * 3 );
  s.col16.resize( n * 3 );
  s.col8.resize( n * 3 );
  s.btype.resize( n );
  if ( n ) {
    std::memcpy( s.pos.data(), pc.positions_.data(), n * 6 );
    if ( pc.colors16bit_.size() == n ) { std::memcpy( s.col16.data(), pc.colors16bit_.data(), n * 6 ); }
    if ( pc.colors_.size() == n ) { std::memcpy( s.col8.data(), pc.colors_.data(), n * 3 ); }
    for ( size_t i = 0; i < n; i++ ) { s.btype[i] = pc.boundaryPointTypes_[i]; }
  }
}

struct QuietStdout {
  int saved = -1;
  QuietStdout() {
    fflush( stdout );
    saved  = dup( 1 );
    int dn = open( "/dev/null", O_WRONLY );
    dup2( dn, 1 );
    close( dn );
  }
  ~QuietStdout() {
    fflush( stdout );
    dup2( saved, 1 );
    close( saved );
  }
};

void fillGpc( GeneratePointCloudParameters& g, const rb200_params& p ) {
  g.occupancyResolution_           = p.occupancy_resolution;
  g.occupancyPrecision_            = p.occupancy_precision;
  g.enableSizeQuantization_        = p.enable_size_quantization != 0;
  g.gridSmoothing_                 = p.grid_smoothing != 0;
  g.gridSize_                      = p.grid_size;
  g.neighborCountSmoothing_        = p.neighbor_count_smoothing;
  g.radius2Smoothing_              = p.radius2_smoothing;
  g.radius2BoundaryDetection_      = p.radius2_boundary_detection;
  g.thresholdSmoothing_            = p.threshold_smoothing;
  g.rawPointColorFormat_           = 0;
  g.nbThread_                      = 1;
  g.multipleStreams_               = p.multiple_streams != 0;  // streams: video m holds map m of every frame
  g.absoluteD1_                    = p.absolute_d1 != 0;
  g.surfaceThickness_              = p.surface_thickness > 0 ? p.surface_thickness : 4;
  g.thresholdColorSmoothing_       = p.threshold_color_smoothing;
  g.cgridSize_                     = 0;
  g.thresholdColorDifference_      = p.threshold_color_difference;
  g.thresholdColorVariation_       = p.threshold_color_variation;
  g.flagGeometrySmoothing_         = p.flag_geometry_smoothing != 0;
  g.flagColorSmoothing_            = p.flag_color_smoothing != 0;
  g.enhancedOccupancyMapCode_      = p.enhanced_occupancy_map_code != 0;
  g.EOMFixBitCount_                = p.eom_fix_bit_count;
  g.thresholdLossyOM_              = p.threshold_lossy_om;
  g.removeDuplicatePoints_         = p.remove_duplicate_points != 0;
  g.mapCountMinus1_                = p.map_count_minus1;
  g.pointLocalReconstruction_      = p.point_local_reconstruction != 0;
  g.singleMapPixelInterleaving_    = p.single_map_pixel_interleaving != 0;
  g.useAdditionalPointsPatch_      = p.use_additional_points_patch != 0;
  g.useAuxSeperateVideo_           = p.use_aux_separate_video != 0;
  g.plrlNumberOfModes_             = 0;
  g.geometryBitDepth3D_            = p.geometry_bitdepth_3d;
  g.geometry3dCoordinatesBitdepth_ = p.geometry_bitdepth_3d;
  g.pbfEnableFlag_                 = p.pbf_enable != 0;  // setPostProcessingSeiParameters, PCCDecoder.cpp:627-640
  g.pbfPassesCount_                = p.pbf_passes_count;
  g.pbfFilterSize_                 = p.pbf_filter_size;
  g.pbfLog2Threshold_              = p.pbf_log2_threshold;
}

}  // namespace

// point local reconstruction tables for the next ref_gof_run (same layout as rb200_gof_set_plr); consumed by that run
static const rb200_plr* g_pending_plr = nullptr;

struct ref_gof {
  int                   nFrames = 0;
  std::vector<FrameOut> frames;
};

extern "C" {

void ref_gof_set_plr( const rb200_plr* plr ) { g_pending_plr = plr; }

// the patch tables of every frame as PCCDecoder::createPatchFrameDataStructure leaves them in the tiles
// (PCCDecoder.cpp:869-1238), built from the flat rows
static void fillPatchTables( PCCContext& context, const rb200_params& p, int nFrames, const rb200_atlas* at ) {
  const size_t W = p.width, H = p.height;
  for ( int f = 0; f < nFrames; f++ ) {
    auto& afc = context[f];
    afc.setAtlasFrameWidth( W );
    afc.setAtlasFrameHeight( H );
    afc.setNumTilesInAtlasFrame( 1 );
    auto& tile = afc.getTile( 0 );
    tile.setFrameIndex( f );
    tile.setTileIndex( 0 );
    tile.setLeftTopXInFrame( 0 );
    tile.setLeftTopYInFrame( 0 );
    tile.setUseRawPointsSeparateVideo( p.use_aux_separate_video != 0 );  // PCCDecoder.cpp: asps.getAuxiliaryVideoEnabledFlag
    if ( p.use_aux_separate_video ) {  // setTilePartitionSizeAfti, PCCDecoder.cpp:1820-1824 (one tile)
      context[f].setAuxVideoWidth( p.aux_width );
      context[f].resizeAuxTileLeftTopY( 1, 0 );
      context[f].resizeAuxTileHeight( 1, p.aux_height );
    }
    tile.setLog2PatchQuantizerSizeX( p.log2_quantizer_x );
    tile.setLog2PatchQuantizerSizeY( p.log2_quantizer_y );
    auto& patches = tile.getPatches();
    for ( int i = at->patch_offset[f]; i < at->patch_offset[f + 1]; i++ ) {
      const rb200_patch& s = at->patches[i];
      PCCPatch           d;
      d.setIndex( i - at->patch_offset[f] );
      d.setFrameIndex( f );
      d.setTileIndex( 0 );
      d.setU0( s.u0 );
      d.setV0( s.v0 );
      d.setSizeU0( s.size_u0 );
      d.setSizeV0( s.size_v0 );
      d.setU1( s.u1 );
      d.setV1( s.v1 );
      d.setD1( s.d1 );
      d.setOccupancyResolution( p.occupancy_resolution );
      d.setAxis( s.axis_of_additional_plane, s.normal_axis, s.tangent_axis, s.bitangent_axis, s.projection_mode );
      d.setPatchOrientation( s.orientation );
      d.setLodScaleX( s.lod_x );
      d.setLodScaleYIdc( s.lod_y );
      d.setPatchSize2DXInPixel( s.size2d_x_px );
      d.setPatchSize2DYInPixel( s.size2d_y_px );
      if ( p.point_local_reconstruction && g_pending_plr ) {  // PCCDecoder::setPLRData (PCCDecoder.cpp:552-578), block level
        d.allocOneLayerData();
        d.getPointLocalReconstructionLevel() = 0;
        for ( int v0 = 0; v0 < s.size_v0; v0++ ) {
          for ( int u0 = 0; u0 < s.size_u0; u0++ ) {
            d.setPointLocalReconstructionMode( u0, v0, g_pending_plr->block_mode[g_pending_plr->block_offset[i] + v0 * s.size_u0 + u0] );
          }
        }
      }
      patches.push_back( d );
    }
    if ( at->eom_patches && at->eom_offset ) {
      for ( int j = at->eom_offset[f]; j < at->eom_offset[f + 1]; j++ ) {
        const rb200_eom_patch& s = at->eom_patches[j];
        PCCEomPatch            e;
        e.u0_                  = s.u0;
        e.v0_                  = s.v0;
        e.sizeU_               = 0;
        e.sizeV_               = 0;
        e.isPatchInAuxVideo_   = p.use_aux_separate_video != 0;
        e.tileIndex_           = 0;
        e.frameIndex_          = f;
        e.eomCount_            = s.eom_count;
        e.occupancyResolution_ = p.occupancy_resolution;
        for ( int k = 0; k < s.member_count; k++ ) { e.memberPatches_.push_back( at->eom_members[s.member_begin + k] ); }
        tile.getEomPatches().push_back( e );
      }
    }
    size_t totalRaw = 0;
    if ( at->raw_patches && at->raw_offset ) {
      for ( int j = at->raw_offset[f]; j < at->raw_offset[f + 1]; j++ ) {
        const rb200_raw_patch& s = at->raw_patches[j];
        PCCRawPointsPatch      r;
        r.u0_                  = s.u0;
        r.v0_                  = s.v0;
        r.sizeU0_              = s.size_u0;
        r.sizeV0_              = s.size_v0;
        r.u1_                  = s.u1;
        r.v1_                  = s.v1;
        r.d1_                  = s.d1;
        r.occupancyResolution_ = p.occupancy_resolution;
        r.isPatchInAuxVideo_   = false;
        r.tileIndex_           = 0;
        r.frameIndex_          = f;
        r.setNumberOfRawPoints( s.num_points );
        tile.getRawPointsPatches().push_back( r );
        totalRaw += s.num_points;
      }
    }
    // PCCDecoder::createPatchFrameDataStructure sets this from the syntax (PCCDecoder.cpp:1150-1238)
    tile.setTotalNumberOfRawPoints( p.use_additional_points_patch ? totalRaw : 0 );
    {  // ... and the EOM total (:1237): generateRawPointsAttributefromVideo sizes the EOM colours with it before the frame is built
      size_t totalEom = 0;
      for ( auto& e : tile.getEomPatches() ) { totalEom += e.eomCount_; }
      tile.setTotalNumberOfEOMPoints( p.enhanced_occupancy_map_code ? totalEom : 0 );
    }
  }

}

// Runs the reference on a whole GOF.  keep_mask bit s keeps the snapshot after stage s
// (0 reconstruct+colour, 1 geometry smoothing, 2 colour transfer, 3 colour smoothing, 4 RGB8).
// n_threads > 1 = frame-parallel (one PCCCodec instance per worker; frames are independent).
// want_canonical_md5: also compute computeChecksum(true) (slow: nested std::map).
ref_gof* ref_gof_run( const rb200_params* pp,
                      int                 nFrames,
                      const rb200_frames* fr,
                      const rb200_atlas*  at,
                      uint32_t            keep_mask,
                      int                 n_threads,
                      int                 want_canonical_md5,
                      int                 quiet ) {
  const rb200_params& p = *pp;
  const size_t        W = p.width, H = p.height, P = p.occupancy_precision, M = p.map_count_minus1 + 1;
  const size_t        oW = W / P, oH = H / P;
  std::unique_ptr<QuietStdout> q;
  if ( quiet ) { q.reset( new QuietStdout ); }

  PCCContext context;
  context.resizeAtlas( 1 );
  context.setAtlasIndex( 0 );
  context.allocateAtlasHLS( 1 );
  auto& asps = context.addAtlasSequenceParameterSet();
  asps.setPatchPrecedenceOrderFlag( p.patch_precedence_reverse != 0 );
  auto& vps = context.addV3CParameterSet();
  vps.init( 0, 0, (uint16_t)W, (uint16_t)H, 30, (uint32_t)( M - 1 ), false, false, true, true, p.attribute_count > 0 );
  vps.allocateMap( 0 );
  auto& ai = vps.getAttributeInformation( 0 );
  ai.setAttributeCount( p.attribute_count > 0 ? 1 : 0 );
  ai.allocate();
  if ( p.attribute_count > 0 ) {
    ai.setAttributeDimensionMinus1( 0, 2 );
    ai.setAttributeDimensionPartitionsMinus1( 0, 0 );
    ai.setAttribute2dBitdepthMinus1( 0, 7 );
  }
  context.setActiveVpsId( 0 );
  context.getAtlas( 0 ).allocateVideoFrames( context, 0 );
  context.resize( nFrames );

  // ---- frame ingest (PCCImage layout: planar channels_, PCCImage.h:205-212) ----
  auto& occVideo = context.getVideoOccupancyMap();
  occVideo.resize( nFrames );
  auto& geoVideos = context.getVideoGeometryMultiple();
  const bool streams = p.multiple_streams != 0;  // PCCCodec.cpp:609-618: video m, frame f instead of video 0, frame f*M+m
  geoVideos.resize( streams ? M : 1 );
  for ( size_t m = 0; m < ( streams ? M : 1 ); m++ ) { geoVideos[m].resize( streams ? nFrames : nFrames * M ); }
  for ( int f = 0; f < nFrames; f++ ) {
    auto& o = occVideo.getFrame( f );
    o.resize( oW, oH, PCCCOLORFORMAT::YUV444 );
    std::memcpy( o.getChannel( 0 ).data(), fr->occupancy + (size_t)f * oW * oH, oW * oH );
    for ( size_t m = 0; m < M; m++ ) {
      auto& g = streams ? geoVideos[m].getFrame( f ) : geoVideos[0].getFrame( f * M + m );
      g.resize( W, H, PCCCOLORFORMAT::YUV444 );
      std::memcpy( g.getChannel( 0 ).data(), fr->geometry + ( (size_t)f * M + m ) * W * H, W * H * 2 );
    }
  }
  if ( p.attribute_count > 0 ) {
    auto& attrVideos = context.getVideoAttributesMultiple();
    if ( streams && attrVideos.size() < M ) { attrVideos.resize( M ); }
    for ( size_t m = 0; m < ( streams ? M : 1 ); m++ ) { attrVideos[m].resize( streams ? nFrames : nFrames * M ); }
    for ( int f = 0; f < nFrames; f++ ) {
      for ( size_t m = 0; m < M; m++ ) {
        auto& a = streams ? attrVideos[m].getFrame( f ) : attrVideos[0].getFrame( f * M + m );
        a.resize( W, H, p.attribute_rgb444 ? PCCCOLORFORMAT::RGB444 : PCCCOLORFORMAT::YUV444 );
        a.setDeprecatedColorFormat( p.attribute_rgb444 ? 0 : 1 );  // PCCVideoDecoder.cpp:130-136
        for ( int c = 0; c < 3; c++ ) {
          std::memcpy( a.getChannel( c ).data(), fr->attribute + ( ( (size_t)f * M + m ) * 3 + c ) * W * H, W * H * 2 );
        }
      }
    }
  }

  if ( p.use_aux_separate_video ) {  // the auxiliary video's frames (PCCDecoder.cpp:187-230 decodes them into these)
    const size_t Wa = p.aux_width, Ha = p.aux_height;
    auto&        rawGeo = context.getVideoRawPointsGeometry();
    rawGeo.resize( nFrames );
    for ( int f = 0; f < nFrames; f++ ) {
      auto& g = rawGeo.getFrame( f );
      g.resize( Wa, Ha, PCCCOLORFORMAT::YUV444 );
      if ( fr->aux_geometry ) { std::memcpy( g.getChannel( 0 ).data(), fr->aux_geometry + (size_t)f * Wa * Ha, Wa * Ha * 2 ); }
    }
    if ( p.attribute_count > 0 ) {
      auto& rawAtt = context.getVideoRawPointsAttribute();
      rawAtt.resize( nFrames );
      for ( int f = 0; f < nFrames; f++ ) {
        auto& a = rawAtt.getFrame( f );
        a.resize( Wa, Ha, PCCCOLORFORMAT::YUV444 );
        for ( int c = 0; c < 3; c++ ) {
          std::memcpy( a.getChannel( c ).data(), fr->aux_attribute + ( (size_t)f * 3 + c ) * Wa * Ha, Wa * Ha * 2 );
        }
      }
    }
  }

  if ( p.point_local_reconstruction && g_pending_plr ) {  // PCCDecoder::setPointLocalReconstruction (PCCDecoder.cpp:528-541)
    for ( int i = 0; i < g_pending_plr->n_modes; i++ ) {
      const rb200_plr_mode&        m = g_pending_plr->modes[i];
      PointLocalReconstructionMode mode = {m.interpolate != 0, m.filling != 0, m.min_d1, m.neighbor};
      context.addPointLocalReconstructionMode( mode );
    }
  }
  fillPatchTables( context, p, nFrames, at );

  GeneratePointCloudParameters gpc;
  fillGpc( gpc, p );
  context.setOccupancyPrecision( P );

  ref_gof* out  = new ref_gof;
  out->nFrames  = nFrames;
  out->frames.resize( nFrames );
  std::atomic<int> next( 0 );

  auto worker = [&]() {
    Oracle codec;
    for ( ;; ) {
      const int f = next.fetch_add( 1 );
      if ( f >= nFrames ) { break; }
      FrameOut&             fo   = out->frames[f];
      auto&                 tile = context[f].getTile( 0 );
      PCCPointSet3          reconstruct;
      std::vector<uint32_t> partition;
      auto                  t0 = std::chrono::steady_clock::now();
      if ( !p.pbf_enable ) {  // PCCDecoder.cpp:362-366
        codec.generateOccupancyMap( tile, occVideo.getFrame( f ), P, p.threshold_lossy_om,
                                    p.enhanced_occupancy_map_code != 0 );
      }
      codec.generateBlockToPatchFromOccupancyMapVideo( context, tile, f, occVideo.getFrame( f ),
                                                       p.occupancy_resolution, P );
      if ( p.use_aux_separate_video ) {  // PCCDecoder.cpp:330-343 (the per-tile forms: the frame loop here runs frame-parallel)
        if ( p.use_additional_points_patch ) { codec.generateRawPointsGeometryfromVideo( context, tile, f ); }
        if ( p.attribute_count > 0 ) { codec.generateRawPointsAttributefromVideo( context, tile, f ); }
      }
      PCCPointSet3 tileRec;
      codec.generatePointCloud( tileRec, context, f, 0, gpc, partition, true );
      reconstruct.appendPointSet( tileRec );
      if ( p.attribute_count > 0 ) {
        reconstruct.addColors();
        reconstruct.addColors16bit();
        std::vector<bool> absoluteT1List( M, true );
        if ( M > 1 && p.relative_t1 ) { absoluteT1List[1] = false; }  // sps.getMapAbsoluteCodingEnableFlag, PCCDecoder.cpp:321
        codec.colorPointCloud( reconstruct, context, tile, absoluteT1List, p.multiple_streams != 0 ? 1 : 0, 1, 0, gpc );
      }
      auto t1          = std::chrono::steady_clock::now();
      fo.msReconstruct = std::chrono::duration<double, std::milli>( t1 - t0 ).count();
      fo.msStage[0]    = fo.msReconstruct;
      fo.counts.total   = (int64_t)reconstruct.getPointCount();
      fo.counts.regular = (int64_t)tile.getTotalNumberOfRegularPoints();
      fo.counts.eom     = (int64_t)tile.getTotalNumberOfEOMPoints();
      fo.counts.raw     = (int64_t)tile.getTotalNumberOfRawPoints();
      fo.partition      = partition;
      {
        auto& p2p = tile.getPointToPixel();
        fo.pointToPixel.resize( p2p.size() * 3 );
        for ( size_t i = 0; i < p2p.size(); i++ ) {
          for ( int k = 0; k < 3; k++ ) { fo.pointToPixel[i * 3 + k] = (uint32_t)p2p[i][k]; }
        }
        auto& b2p = tile.getBlockToPatch();
        fo.blockToPatch.resize( b2p.size() );
        for ( size_t i = 0; i < b2p.size(); i++ ) { fo.blockToPatch[i] = (uint32_t)b2p[i]; }
        auto& om = tile.getOccupancyMap();
        fo.occupancy.resize( om.size() );
        for ( size_t i = 0; i < om.size(); i++ ) { fo.occupancy[i] = (uint8_t)om[i]; }
      }
      if ( keep_mask & 1u ) { snap( fo.stage[0], reconstruct ); }

      // ---- post-processing, PCCDecoder.cpp:431-508 ----
      auto t2 = std::chrono::steady_clock::now();
      if ( p.apply_geo_smoothing && p.flag_geometry_smoothing ) {
        PCCPointSet3 tempFrameBuffer = reconstruct;
        // (neighbor_count_smoothing > 0 without grid smoothing: the encoder-side call with the non-grid filter, PCCCodec.cpp:141)
        if ( p.grid_smoothing || p.neighbor_count_smoothing > 0 ) {
          codec.smoothPointCloudPostprocess( reconstruct, COLOR_TRANSFORM_NONE, gpc, partition );
        }
        auto t3       = std::chrono::steady_clock::now();
        fo.msStage[1] = std::chrono::duration<double, std::milli>( t3 - t2 ).count();
        if ( keep_mask & 2u ) { snap( fo.stage[1], reconstruct ); }
        if ( p.attribute_count > 0 && p.attr_transfer_filter_type == 1 && !p.pbf_enable ) {  // :445
          tempFrameBuffer.transferColors16bitBP( reconstruct, 1, int32_t( 0 ), p.attribute_rgb444 != 0, 8, 1, true, true,
                                                 true, false, 4, 4, 1000, 1000, 1000 * 256, 1000 * 256 );
        }
        auto t4       = std::chrono::steady_clock::now();
        fo.msStage[2] = std::chrono::duration<double, std::milli>( t4 - t3 ).count();
        if ( keep_mask & 4u ) { snap( fo.stage[2], reconstruct ); }
      }
      for ( size_t i = 0; i < reconstruct.getPointCount(); i++ ) {
        if ( reconstruct.getBoundaryPointType( i ) == 3 ) { fo.counts.smoothed++; }
      }
      if ( p.attribute_count > 0 ) {
        auto t5 = std::chrono::steady_clock::now();
        if ( p.apply_attr_smoothing && p.flag_color_smoothing ) {
          std::vector<PCCColor16bit> before;
          before = reconstruct.colors16bit_;
          codec.colorSmoothing( reconstruct, COLOR_TRANSFORM_NONE, gpc );
          for ( size_t i = 0; i < before.size(); i++ ) {
            if ( !( before[i] == reconstruct.colors16bit_[i] ) ) { fo.counts.recolored++; }
          }
        }
        auto t6       = std::chrono::steady_clock::now();
        fo.msStage[3] = std::chrono::duration<double, std::milli>( t6 - t5 ).count();
        if ( keep_mask & 8u ) { snap( fo.stage[3], reconstruct ); }
        if ( !p.attribute_rgb444 ) {
          reconstruct.convertYUV16ToRGB8();
        } else {
          reconstruct.copyRGB16ToRGB8();
        }
        auto t7       = std::chrono::steady_clock::now();
        fo.msStage[4] = std::chrono::duration<double, std::milli>( t7 - t6 ).count();
      }
      auto t8   = std::chrono::steady_clock::now();
      fo.msPost = std::chrono::duration<double, std::milli>( t8 - t2 ).count();
      if ( keep_mask & 16u ) { snap( fo.stage[4], reconstruct ); }
      {
        auto d = reconstruct.computeChecksum( false );
        std::memcpy( fo.md5Ordered, d.data(), 16 );
        if ( want_canonical_md5 ) {
          auto c = reconstruct.computeChecksum( true );
          std::memcpy( fo.md5Canonical, c.data(), 16 );
        }
      }
    }
  };
  if ( n_threads <= 1 ) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for ( int t = 0; t < n_threads; t++ ) { pool.emplace_back( worker ); }
    for ( auto& t : pool ) { t.join(); }
  }
  return out;
}

void ref_gof_free( ref_gof* g ) { delete g; }

int64_t ref_gof_count( ref_gof* g, int frame, int stage ) {
  auto& s = g->frames[frame].stage[stage];
  return s.valid ? (int64_t)s.btype.size() : -1;
}

int ref_gof_fetch( ref_gof* g, int frame, int stage, const rb200_cloud_host* dst ) {
  auto& fo = g->frames[frame];
  auto& s  = fo.stage[stage];
  if ( !s.valid ) { return 1; }
  const size_t n = s.btype.size();
  if ( dst->positions ) { std::memcpy( dst->positions, s.pos.data(), n * 6 ); }
  if ( dst->colors16 ) { std::memcpy( dst->colors16, s.col16.data(), n * 6 ); }
  if ( dst->colors ) { std::memcpy( dst->colors, s.col8.data(), n * 3 ); }
  if ( dst->boundary_types ) { std::memcpy( dst->boundary_types, s.btype.data(), n * 2 ); }
  if ( dst->partition ) { std::memcpy( dst->partition, fo.partition.data(), fo.partition.size() * 4 ); }
  if ( dst->point_to_pixel ) { std::memcpy( dst->point_to_pixel, fo.pointToPixel.data(), fo.pointToPixel.size() * 4 ); }
  return 0;
}

void ref_gof_counts( ref_gof* g, int frame, rb200_frame_counts* out ) { *out = g->frames[frame].counts; }
int64_t ref_gof_partition_size( ref_gof* g, int frame ) { return (int64_t)g->frames[frame].partition.size(); }
int64_t ref_gof_point_to_pixel_size( ref_gof* g, int frame ) { return (int64_t)g->frames[frame].pointToPixel.size() / 3; }

void ref_gof_md5( ref_gof* g, int frame, int canonical, uint8_t* out16 ) {
  std::memcpy( out16, canonical ? g->frames[frame].md5Canonical : g->frames[frame].md5Ordered, 16 );
}
void ref_gof_block_to_patch( ref_gof* g, int frame, uint32_t* dst ) {
  auto& v = g->frames[frame].blockToPatch;
  std::memcpy( dst, v.data(), v.size() * 4 );
}
void ref_gof_occupancy( ref_gof* g, int frame, uint8_t* dst ) {
  auto& v = g->frames[frame].occupancy;
  std::memcpy( dst, v.data(), v.size() );
}
// which: 0 reconstruction (PCCDecoder.cpp:331-428), 1 postProcessing (:429-510), 10+s = stage s
double ref_gof_time_ms( ref_gof* g, int frame, int which ) {
  auto& fo = g->frames[frame];
  if ( which == 0 ) { return fo.msReconstruct; }
  if ( which == 1 ) { return fo.msPost; }
  if ( which >= 10 && which < 15 ) { return fo.msStage[which - 10]; }
  return -1;
}

static void viewToCloud( const rb200_cloud_view& v, PCCPointSet3& pc, bool withNormals ) {
  pc.resize( v.count );
  if ( v.count ) { std::memcpy( pc.positions_.data(), v.positions, v.count * 6 ); }
  if ( v.colors ) {
    pc.addColors();
    if ( v.count ) { std::memcpy( pc.colors_.data(), v.colors, v.count * 3 ); }
  }
  if ( withNormals && v.normals ) {
    pc.addNormals();
    for ( int64_t i = 0; i < v.count; i++ ) {
      for ( int k = 0; k < 3; k++ ) { pc.normals_[i][k] = v.normals[i * 3 + k]; }
    }
  }
}

static void fillQuality( rb200_quality& d, const QualityMetrics& s ) {
  std::memset( &d, 0, sizeof( d ) );
  d.c2c_mse            = s.c2cMse_;
  d.c2c_psnr           = s.c2cPsnr_;
  d.c2p_mse            = s.c2pMse_;
  d.c2p_psnr           = s.c2pPsnr_;
  d.c2c_hausdorff      = s.c2cHausdorff_;
  d.c2c_hausdorff_psnr = s.c2cHausdorffPsnr_;
  d.c2p_hausdorff      = s.c2pHausdorff_;
  d.c2p_hausdorff_psnr = s.c2pHausdorffPsnr_;
  for ( int i = 0; i < 3; i++ ) {
    d.color_mse[i]  = s.colorMse_[i];
    d.color_psnr[i] = s.colorPsnr_[i];
  }
}

// PCCMetrics::compute( sources, reconstructs, normals ) for one frame (PCCMetrics.cpp:334-385).
// `normals` (optional) is the source-with-normals cloud the reference loads from --normalDataPath.
int ref_metrics( const rb200_metrics_params* mp,
                 const rb200_cloud_view*     src,
                 const rb200_cloud_view*     rec,
                 const rb200_cloud_view*     normals,
                 rb200_metrics_result*       out,
                 double*                     ms ) {
  PCCMetricsParameters params;
  params.computeC2c_       = mp->compute_c2c != 0;
  params.computeC2p_       = mp->compute_c2p != 0;
  params.computeColor_     = mp->compute_color != 0;
  params.computeHausdorff_ = mp->compute_hausdorff != 0;
  params.dropDuplicates_   = mp->drop_duplicates;
  params.neighborsProc_    = mp->neighbors_proc;
  params.resolution_       = mp->resolution;
  params.computeLidar_ = params.computeReflectance_ = false;
  PCCGroupOfFrames sources, recs, norms;
  sources.setFrameCount( 1 );
  recs.setFrameCount( 1 );
  viewToCloud( *src, sources[0], false );
  viewToCloud( *rec, recs[0], false );
  if ( normals && normals->count > 0 && normals->normals ) {
    norms.setFrameCount( 1 );
    viewToCloud( *normals, norms[0], true );
  }
  PCCMetrics metrics;
  metrics.setParameters( params );
  auto t0 = std::chrono::steady_clock::now();
  metrics.compute( sources, recs, norms );
  auto t1 = std::chrono::steady_clock::now();
  if ( ms ) { *ms = std::chrono::duration<double, std::milli>( t1 - t0 ).count(); }
  std::memset( out, 0, sizeof( *out ) );
  fillQuality( out->q1, metrics.quality1_[0] );
  fillQuality( out->q2, metrics.quality2_[0] );
  fillQuality( out->qf, metrics.qualityF_[0] );
  out->source_points      = metrics.sourcePoints_[0];
  out->source_after_dedup = metrics.sourceDuplicates_[0];
  out->rec_points         = metrics.reconstructPoints_[0];
  out->rec_after_dedup    = metrics.reconstructDuplicates_[0];
  return 0;
}

// PCCPointSet3::removeDuplicate (PCCPointSet.cpp:169-218)
int64_t ref_remove_duplicates( const rb200_cloud_view* in, int drop, int16_t* outPos, uint8_t* outCol ) {
  PCCPointSet3 pc, res;
  viewToCloud( *in, pc, false );
  pc.removeDuplicate( res, drop );
  const size_t n = res.getPointCount();
  if ( outPos && n ) { std::memcpy( outPos, res.positions_.data(), n * 6 ); }
  if ( outCol && n && res.colors_.size() == n ) { std::memcpy( outCol, res.colors_.data(), n * 3 ); }
  return (int64_t)n;
}

// exact kNN through the reference's nanoflann wrapper (PCCKdTree.cpp:42-79): used to pin the
// nanoflann-order restatement of oracle/ (transferColors16bitBP tie order).
int ref_knn( const int16_t* cloud, int64_t n, const int16_t* queries, int64_t nq, int k, int64_t* outIdx, double* outDist ) {
  PCCPointSet3 pc;
  pc.resize( n );
  std::memcpy( pc.positions_.data(), cloud, n * 6 );
  PCCKdTree   tree( pc );
  PCCNNResult res;
  for ( int64_t i = 0; i < nq; i++ ) {
    PCCPoint3D qp( queries[i * 3], queries[i * 3 + 1], queries[i * 3 + 2] );
    tree.search( qp, k, res );
    for ( int j = 0; j < k; j++ ) {
      outIdx[i * k + j]  = j < (int)res.size() ? (int64_t)res.indices( j ) : -1;
      outDist[i * k + j] = j < (int)res.size() ? res.dist( j ) : -1.0;
    }
  }
  return 0;
}

// PCCKdTree::searchRadius (PCCKdTree.cpp:69-79); sorted == 0: the same index queried with SearchParams::sorted = false
// (nanoflann's traversal order, before std::sort)
int ref_knn_radius( const int16_t* cloud, int64_t n, const int16_t* queries, int64_t nq, double radius2, int maxResults, int sorted,
                    int64_t* outIdx, double* outDist, int32_t* outCount ) {
  PCCPointSet3 pc;
  pc.resize( n );
  std::memcpy( pc.positions_.data(), cloud, n * 6 );
  PCCKdTree tree( pc );
  typedef KDTreeVectorOfVectorsAdaptor<PCCPointSet3, PCCType, float, 3, metric_L2_Simple_2, size_t> Adaptor;  // PCCKdTree.cpp:42
  Adaptor raw( 3, pc, 10 );
  for ( int64_t i = 0; i < nq; i++ ) {
    PCCPoint3D qp( queries[i * 3], queries[i * 3 + 1], queries[i * 3 + 2] );
    std::vector<std::pair<size_t, double> > ret;
    {
      nanoflann::SearchParams sp;
      sp.sorted = false;
      raw.index->radiusSearch( &qp[0], radius2, ret, sp );
    }
    outCount[i] = (int32_t)ret.size();
    if ( sorted ) {
      PCCNNResult res;
      tree.searchRadius( qp, (size_t)maxResults, radius2, res );
      for ( int j = 0; j < maxResults; j++ ) {
        outIdx[i * maxResults + j]  = j < (int)res.count() ? (int64_t)res.indices( j ) : -1;
        outDist[i * maxResults + j] = j < (int)res.count() ? res.dist( j ) : -1.0;
      }
    } else {
      for ( int j = 0; j < maxResults; j++ ) {
        outIdx[i * maxResults + j]  = j < (int)ret.size() ? (int64_t)ret[j].first : -1;
        outDist[i * maxResults + j] = j < (int)ret.size() ? ret[j].second : -1.0;
      }
    }
  }
  return 0;
}

// PCCPointSet3::write / read (PCCPointSet.cpp:359-757) on a flat cloud: pins the PLY wire format
int ref_write_ply( const rb200_cloud_view* in, const char* path ) {
  PCCPointSet3 pc;
  viewToCloud( *in, pc, false );
  return pc.write( path, false ) ? 0 : 1;
}
int64_t ref_read_ply( const char* path, int16_t* outPos, uint8_t* outCol, int64_t cap ) {
  PCCPointSet3 pc;
  if ( !pc.read( path ) ) { return -1; }
  const int64_t n = (int64_t)pc.getPointCount();
  if ( n > cap ) { return n; }
  if ( outPos && n ) { std::memcpy( outPos, pc.positions_.data(), n * 6 ); }
  if ( outCol && n && pc.colors_.size() == (size_t)n ) { std::memcpy( outCol, pc.colors_.data(), n * 3 ); }
  return n;
}

// PCCVideoDecoder's inverse colour conversion (PccLibDecoder/source/PCCVideoDecoder.cpp:125-146, :365): the decoded
// 4:2:0 frame through the reference's own PCCInternalColorConverter<uint16_t>::convert( "YUV420ToYUV444_<bits>_<filter>" ).
// y [H][W], u / v [H/2][W/2] samples as uint16; out [3][H][W] uint16 (16-bit 4:4:4)
int ref_yuv420_to_yuv444( const uint16_t* y, const uint16_t* u, const uint16_t* v, int W, int H, int bitdepth, int filter,
                          uint16_t* out ) {
  PCCVideo<uint16_t, 3> src, dst;
  src.resize( 1 );
  src[0].resize( W, H, PCCCOLORFORMAT::YUV420 );
  std::copy( y, y + (size_t)W * H, src[0].getChannel( 0 ).begin() );
  std::copy( u, u + (size_t)( W / 2 ) * ( H / 2 ), src[0].getChannel( 1 ).begin() );
  std::copy( v, v + (size_t)( W / 2 ) * ( H / 2 ), src[0].getChannel( 2 ).begin() );
  PCCInternalColorConverter<uint16_t> conv;
  conv.convert( "YUV420ToYUV444_" + std::to_string( bitdepth ) + "_" + std::to_string( filter ), src, dst, "", "" );
  if ( dst.getFrameCount() != 1 || (int)dst[0].getWidth() != W || (int)dst[0].getHeight() != H ) { return -1; }
  for ( int c = 0; c < 3; c++ ) { std::copy( dst[0].getChannel( c ).begin(), dst[0].getChannel( c ).end(), out + (size_t)c * W * H ); }
  return 0;
}

// PCCImage<uint16_t,3>::set (PCCImage.h:97-138) as the HM wrapper calls it (PCCHMLibVideoDecoderImpl.cpp:360-363): decoder
// picture planes (Pel = int16, dense strides here) -> stored samples with the rounding shift and clamp.
// out: Y [H][W] then U, V [H/2][W/2] uint16
int ref_image_set_yuv420( const int16_t* y, const int16_t* u, const int16_t* v, int W, int H, int shift, uint16_t* out ) {
  PCCImage<uint16_t, 3> img;
  img.set( y, u, v, (size_t)W, (size_t)H, (size_t)W, (size_t)( W / 2 ), (size_t)( H / 2 ), (size_t)( W / 2 ), (int16_t)shift,
           PCCCOLORFORMAT::YUV420, false );
  uint16_t* o = out;
  for ( int c = 0; c < 3; c++ ) { o = std::copy( img.getChannel( c ).begin(), img.getChannel( c ).end(), o ); }
  return 0;
}

// PCCImage<T, 3>::convertBitdepth (PCCImage.cpp:258-299) on one plane, in place: T = uint16_t (geometry video,
// PCCDecoder.cpp:148-149) or uint8_t (occupancy video, :119)
int ref_convert_bitdepth_u16( uint16_t* plane, int W, int H, int in, int out, int msb ) {
  PCCImage<uint16_t, 3> img;
  img.resize( (size_t)W, (size_t)H, PCCCOLORFORMAT::YUV444 );
  std::copy( plane, plane + (size_t)W * H, img.getChannel( 0 ).begin() );
  img.convertBitdepth( (uint8_t)in, (uint8_t)out, msb != 0 );
  std::copy( img.getChannel( 0 ).begin(), img.getChannel( 0 ).end(), plane );
  return 0;
}
int ref_convert_bitdepth_u8( uint8_t* plane, int W, int H, int in, int out, int msb ) {
  PCCImage<uint8_t, 3> img;
  img.resize( (size_t)W, (size_t)H, PCCCOLORFORMAT::YUV444 );
  std::copy( plane, plane + (size_t)W * H, img.getChannel( 0 ).begin() );
  img.convertBitdepth( (uint8_t)in, (uint8_t)out, msb != 0 );
  std::copy( img.getChannel( 0 ).begin(), img.getChannel( 0 ).end(), plane );
  return 0;
}

// patch-table export (rabbit-transcoding_b200/host/rb200_atlas_export.h): the rows are turned into the reference's tile
// containers exactly as for a GOF run, exported again, and handed back — they must come back unchanged.
// out_* arrays are sized like the inputs; returns the number of patches, or -1 when a table does not fit.
int ref_atlas_export_roundtrip( const rb200_params* pp, int nFrames, const rb200_atlas* at, rb200_patch* outPatches,
                                int32_t* outPatchOffset, rb200_eom_patch* outEoms, int32_t* outEomOffset, int32_t* outMembers,
                                rb200_raw_patch* outRaws, int32_t* outRawOffset ) {
  PCCContext context;
  context.resizeAtlas( 1 );
  context.setAtlasIndex( 0 );
  context.allocateAtlasHLS( 1 );
  context.addAtlasSequenceParameterSet();
  auto& vps = context.addV3CParameterSet();
  vps.init( 0, 0, (uint16_t)pp->width, (uint16_t)pp->height, 30, (uint32_t)pp->map_count_minus1, false, false, true, true, true );
  vps.allocateMap( 0 );
  context.setActiveVpsId( 0 );
  context.resize( nFrames );
  fillPatchTables( context, *pp, nFrames, at );
  rb200::AtlasTables t;
  rb200::exportAtlas( context, t );
  std::copy( t.patches.begin(), t.patches.end(), outPatches );
  std::copy( t.patchOffset.begin(), t.patchOffset.end(), outPatchOffset );
  if ( outEoms ) { std::copy( t.eoms.begin(), t.eoms.end(), outEoms ); }
  if ( outEomOffset ) { std::copy( t.eomOffset.begin(), t.eomOffset.end(), outEomOffset ); }
  if ( outMembers ) { std::copy( t.members.begin(), t.members.end(), outMembers ); }
  if ( outRaws ) { std::copy( t.raws.begin(), t.raws.end(), outRaws ); }
  if ( outRawOffset ) { std::copy( t.rawOffset.begin(), t.rawOffset.end(), outRawOffset ); }
  return (int)t.patches.size();
}

int ref_abi_version( void ) { return RB200_ABI_VERSION; }

}  // extern "C"
