// rb200_atlas_export.h — patch-table export: the atlas layer of the reference (what PCCDecoder::
// createPatchFrameDataStructure, PccLibDecoder/source/PCCDecoder.cpp:869-1238, leaves in every tile: PCCPatch,
// PCCEomPatch and PCCRawPointsPatch objects) as the flat rows of include/rabbit_b200.h, once per GOF.
// Host glue that compiles against the unmodified reference headers; used by the shim (PCCCodecB200.cpp) and usable on
// its own by a host that keeps the reference's containers but drives the C ABI itself.
#pragma once
#include <cstdint>
#include <vector>

#include "PCCContext.h"
#include "PCCFrameContext.h"
#include "PCCPatch.h"
#include "rabbit_b200.h"

namespace rb200 {

struct AtlasTables {
  std::vector<rb200_patch>     patches;
  std::vector<rb200_eom_patch> eoms;
  std::vector<rb200_raw_patch> raws;
  std::vector<int32_t>         patchOffset{0}, eomOffset{0}, rawOffset{0}, members;
  // the view the C ABI takes; valid while this object lives
  rb200_atlas view() {
    if ( members.empty() ) { members.push_back( 0 ); }
    return rb200_atlas{patches.data(),
                       patchOffset.data(),
                       eoms.empty() ? nullptr : eoms.data(),
                       eomOffset.data(),
                       members.data(),
                       raws.empty() ? nullptr : raws.data(),
                       rawOffset.data()};
  }
};

// one PCCPatch (PCCPatch.h:353-408) -> one row; setViewId (PCCPatch.cpp:111-137) has been applied by the atlas layer, so
// the axes and the projection mode are read back from the patch
inline rb200_patch exportPatch( const pcc::PCCPatch& s ) {
  rb200_patch d{};
  d.u0 = (int32_t)s.getU0(), d.v0 = (int32_t)s.getV0(), d.size_u0 = (int32_t)s.getSizeU0(), d.size_v0 = (int32_t)s.getSizeV0();
  d.u1 = (int32_t)s.getU1(), d.v1 = (int32_t)s.getV1(), d.d1 = (int32_t)s.getD1();
  d.normal_axis = (int32_t)s.getNormalAxis(), d.tangent_axis = (int32_t)s.getTangentAxis(), d.bitangent_axis = (int32_t)s.getBitangentAxis();
  d.projection_mode = (int32_t)s.getProjectionMode(), d.orientation = (int32_t)s.getPatchOrientation();
  d.lod_x = (int32_t)s.getLodScaleX(), d.lod_y = (int32_t)s.getLodScaleY();
  d.axis_of_additional_plane = (int32_t)s.getAxisOfAdditionalPlane();
  d.size2d_x_px = (int32_t)s.getPatchSize2DXInPixel(), d.size2d_y_px = (int32_t)s.getPatchSize2DYInPixel();
  return d;
}

// every frame of the context (single tile per atlas frame: tile 0), frames [0, context.size())
inline void exportAtlas( pcc::PCCContext& context, AtlasTables& out ) {
  out = AtlasTables{};
  for ( size_t f = 0; f < context.size(); f++ ) {
    auto& tile = context[f].getTile( 0 );
    for ( auto& s : tile.getPatches() ) { out.patches.push_back( exportPatch( s ) ); }
    out.patchOffset.push_back( (int32_t)out.patches.size() );
    for ( auto& s : tile.getEomPatches() ) {  // PCCEomPatch, PCCPatch.h:439-451
      rb200_eom_patch e{};
      e.u0 = (int32_t)s.u0_, e.v0 = (int32_t)s.v0_, e.member_begin = (int32_t)out.members.size();
      e.member_count = (int32_t)s.memberPatches_.size(), e.eom_count = (int32_t)s.eomCount_;
      for ( auto m : s.memberPatches_ ) { out.members.push_back( (int32_t)m ); }
      out.eoms.push_back( e );
    }
    out.eomOffset.push_back( (int32_t)out.eoms.size() );
    for ( auto& s : tile.getRawPointsPatches() ) {  // PCCRawPointsPatch, PCCPatch.h:453-...
      rb200_raw_patch r{};
      r.u0 = (int32_t)s.u0_, r.v0 = (int32_t)s.v0_, r.size_u0 = (int32_t)s.sizeU0_, r.size_v0 = (int32_t)s.sizeV0_;
      r.u1 = (int32_t)s.u1_, r.v1 = (int32_t)s.v1_, r.d1 = (int32_t)s.d1_, r.num_points = (int32_t)s.getNumberOfRawPoints();
      out.raws.push_back( r );
    }
    out.rawOffset.push_back( (int32_t)out.raws.size() );
  }
}

}  // namespace rb200
