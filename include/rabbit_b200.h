/* rabbit_b200.h — C ABI of the B200-native V-PCC reconstruction / smoothing / metrics hot path.
 *
 * This is the drop-in boundary beneath the reference's C++ entry points.  The
 * reference (mic-rud/RABBIT-Transcoding, a fork of MPEG TMC2 v15) has no FFI for this path: its
 * boundary is a set of C++ member functions of pcc::PCCCodec / pcc::PCCMetrics.  Every entry point
 * below names the reference function(s) whose *body* it replaces (paths relative to
 * /root/reference/source/lib).  INTEGRATION.md shows the reference-side shim that marshals the
 * reference's containers (PCCContext / PCCFrameContext / PCCPatch / PCCVideo / PCCPointSet3) into
 * these flat calls.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types.  Pointers named "host or device" are
 *    resolved with cudaMemcpyDefault (UVA), so pinned host, pageable host and device memory all work.
 *  - every call returns an rb200_status; nothing here calls exit().  The reference's behaviour
 *    (printf + exit(code), PCCPatch.cpp:237-245, PCCMetrics.cpp:342-346) is restored by the C++ shim.
 *  - all work is enqueued on the context's CUDA stream (rb200_set_stream); calls that return data to
 *    the host synchronise that stream themselves.
 *  - one context per GPU, not re-entrant (the reference's PCCCodec keeps per-instance scratch too,
 *    PccLibCommon/include/PCCCodec.h:415-423).
 */
#ifndef RABBIT_B200_H
#define RABBIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB200_ABI_VERSION 3

typedef enum rb200_status {
  RB200_OK                   = 0,
  RB200_ERR_INVALID          = 1, /* bad argument / inconsistent sizes                                  */
  RB200_ERR_UNSUPPORTED      = 2, /* a reference mode this build does not implement (fails loudly)      */
  RB200_ERR_CUDA             = 3, /* CUDA runtime error; see rb200_error_string                         */
  RB200_ERR_NOMEM            = 4,
  RB200_ERR_STATE            = 5, /* call order violated (e.g. smoothing before reconstruction)         */
  RB200_ERR_TIE_OVERFLOW     = 6, /* metrics: a nearest-distance tie set exceeded 30 (PCCMetrics.cpp:88) */
  RB200_ERR_PATCH_OUT_OF_CANVAS = 180 /* PccLibCommon/source/PCCPatch.cpp:237-245 exits with this code */
} rb200_status;

/* ------------------------------------------------------------------------------------------------
 * Patch table rows.  One row per patch per frame; fields are the PCCPatch members the path reads
 * (PccLibCommon/include/PCCPatch.h:353-408), already resolved by PCCDecoder::createPatchFrameDataStructure
 * (PccLibDecoder/source/PCCDecoder.cpp:869-1238).  setViewId (PCCPatch.cpp:111-137) is applied by the
 * caller: axes and projection mode are stored explicitly.
 * ---------------------------------------------------------------------------------------------- */
typedef struct rb200_patch {
  int32_t u0, v0;           /* canvas position, in occupancyResolution blocks (u0_, v0_)               */
  int32_t size_u0, size_v0; /* patch size in blocks (sizeU0_, sizeV0_)                                  */
  int32_t u1, v1, d1;       /* tangent / bitangent / normal shift (u1_, v1_, d1_)                       */
  int32_t normal_axis, tangent_axis, bitangent_axis; /* 0..2                                           */
  int32_t projection_mode;  /* 0: d + d1, 1: max(d1 - d, 0)   (PCCPatch.h:177-186)                      */
  int32_t orientation;      /* enum PCCPatchOrientation 0..8  (PCCBitstreamCommon.h:120-130)            */
  int32_t lod_x, lod_y;     /* levelOfDetailX_/Y_                                                       */
  int32_t axis_of_additional_plane; /* 0, or 1..3 for 45-degree planes (PCCCodec.cpp:2503-2524)         */
  int32_t size2d_x_px, size2d_y_px; /* getPatchSize2DX/YInPixel, used by size quantisation :571-597     */
} rb200_patch;

/* PCCEomPatch (PCCPatch.h:439-451): members are indices into eom_members[] */
typedef struct rb200_eom_patch {
  int32_t u0, v0;               /* in blocks                                                            */
  int32_t member_begin, member_count; /* memberPatches_ = eom_members[member_begin .. +member_count)    */
  int32_t eom_count;            /* eomCount_ (only summed into TotalNumberOfEOMPoints, :854)            */
} rb200_eom_patch;

/* PCCRawPointsPatch (PCCPatch.h:453-…), non-auxiliary-video case of PCCCodec.cpp:894-949 */
typedef struct rb200_raw_patch {
  int32_t u0, v0, size_u0, size_v0; /* in blocks of occupancy_resolution                                */
  int32_t u1, v1, d1;               /* offsets added to the X / Y / Z planes                            */
  int32_t num_points;               /* numberOfRawPoints_                                               */
} rb200_raw_patch;

/* ------------------------------------------------------------------------------------------------
 * Per-GOF parameters: the fields of GeneratePointCloudParameters (PCCCodec.h:62-100) that the path
 * reads, the decoder switches of PCCDecoderParameters.cpp:114-157, and the atlas geometry.
 * ---------------------------------------------------------------------------------------------- */
typedef struct rb200_params {
  int32_t width, height;            /* atlas frame size (single tile: tile == atlas)                    */
  int32_t occupancy_resolution;     /* block size R (16)                                                */
  int32_t occupancy_precision;      /* p: occupancy video is (W/p)x(H/p)                                */
  int32_t threshold_lossy_om;       /* oi.getLossyOccupancyCompressionThreshold(), PCCDecoder.cpp:364   */
  int32_t map_count_minus1;         /* 0 or 1                                                           */
  int32_t absolute_d1;
  int32_t remove_duplicate_points;
  int32_t enhanced_occupancy_map_code;
  int32_t eom_fix_bit_count;
  int32_t enable_size_quantization;
  int32_t log2_quantizer_x, log2_quantizer_y;
  int32_t patch_precedence_reverse; /* bDecoder && asps.patchPrecedenceOrderFlag, PCCCodec.cpp:627-629  */
  int32_t use_additional_points_patch; /* raw patches in the geometry video                             */
  int32_t total_raw_points_known;   /* tile.getTotalNumberOfRawPoints() is set by the caller's syntax layer */
  int32_t single_map_pixel_interleaving; /* generatePoints :350-471: one map, layers on a checkerboard; needs
                                          * map_count_minus1 == 0, surface_thickness >= 1, no EOM / multiple streams */
  int32_t point_local_reconstruction;    /* generatePoints :472-496; needs one map and rb200_gof_set_plr */
  int32_t pbf_enable;                    /* Rec-2 occupancy synthesis (patch border filtering, PCCCodec.cpp:541-554):
                                          * one or two maps, no EOM / pixel interleaving / PLR / size quantisation,
                                          * every patch at level of detail 1 without an additional plane; the
                                          * pbf_* fields at the end of this struct hold the SEI's parameters   */
  int32_t multiple_streams;         /* sps.getMultipleMapStreamsPresentFlag: the caller still hands planes
                                     * as [F][M][..][H][W] (map m of frame f = frame f of stream m)       */
  int32_t attribute_count;          /* 0: colours become 127 (PCCCodec.cpp:1327-1330)                   */
  int32_t attribute_rgb444;         /* 1: copyRGB16ToRGB8 instead of convertYUV16ToRGB8                 */
  int32_t geometry_bitdepth_3d;     /* geometryBitDepth3D_ (10 / 11)                                    */
  /* geometry smoothing SEI (PCCDecoder.cpp:707-722) + decoder switch applyGeoSmoothingType             */
  int32_t flag_geometry_smoothing;
  int32_t grid_smoothing;
  int32_t grid_size;
  int32_t apply_geo_smoothing;      /* params_.applyGeoSmoothingType_ != 0                              */
  int32_t attr_transfer_filter_type;/* params_.attrTransferFilterType_: 0 none, 1 transferColors16bitBP */
  /* attribute smoothing SEI (PCCDecoder.cpp:754-775) + decoder switch applyAttrSmoothingType           */
  int32_t flag_color_smoothing;
  int32_t apply_attr_smoothing;
  int32_t relative_t1;              /* multiple_streams only: map 1 of the attribute is a delta on map 0
                                     * (!absoluteT1List[1], PCCCodec.cpp:1387-1416; CTC condition T1-from-rec-T0) */
  int32_t surface_thickness;        /* surfaceThickness_ (ctc-common.cfg:25: 4); read by pixel interleaving only */
  double  threshold_smoothing;
  double  threshold_color_smoothing;
  double  threshold_color_difference;
  double  threshold_color_variation;
  /* occupancy synthesis SEI (PCCDecoder.cpp:627-640): pbfPassesCount_, pbfFilterSize_, pbfLog2Threshold_ */
  int32_t pbf_passes_count;
  int32_t pbf_filter_size;
  int32_t pbf_log2_threshold;
  /* the non-grid smoothPointCloud (PCCCodec.cpp:1106-1157), reached through smoothPointCloudPostprocess (:141) when
   * grid_smoothing == 0 — the encoder's reconstruction (PCCEncoder.cpp:7905-7907); the decoder skips geometry smoothing
   * in that case (PCCDecoder.cpp:436) and leaves neighbor_count_smoothing at 0, which rb200_decode_gof takes as "skip" */
  int32_t neighbor_count_smoothing;   /* neighborCountSmoothing_ (4 * 16)                                 */
  double  radius2_smoothing;          /* radius2Smoothing_ (64)                                           */
  double  radius2_boundary_detection; /* radius2BoundaryDetection_ (64)                                   */
  /* raw (missed) points carried in the auxiliary video (asps.getAuxiliaryVideoEnabledFlag, PCCDecoder.cpp:783;
   * tile.getUseRawPointsSeparateVideo, PCCCodec.cpp:606, :895-897, :1334, :1436-1439): the raw patches address the
   * auxiliary frames (rb200_frames.aux_geometry / aux_attribute) instead of the atlas; their colours are the low 8
   * bits of the auxiliary attribute samples (they pass through PCCColor3B, PCCCodec.cpp:1541-1543).  With EOM the
   * colours of the EOM points come from the auxiliary attribute video as well (:1551-1580), their synthetic pixel
   * addresses start at (0, 0) and the occupancy map is not marked (:852-853, :880). */
  int32_t use_aux_separate_video;
  int32_t aux_width, aux_height;      /* size of the auxiliary video frames                               */
  int32_t reserved0;
} rb200_params;

/* Decoded video planes of one GOF (what PCCVideoDecoder leaves in PCCContext, PCCContext.h:48-50).
 * host or device pointers. */
typedef struct rb200_frames {
  const uint8_t*  occupancy; /* [F][H/p][W/p]      channel 0 of PCCVideoOccupancyMap frame f             */
  const uint16_t* geometry;  /* [F][M][H][W]       channel 0 of geometry frame f*M+m (PCCCodec.cpp:613)   */
  const uint16_t* attribute; /* [F][M][3][H][W]    4:4:4 16-bit attribute frame f*M+m; NULL if none       */
  /* params.use_aux_separate_video: context.getVideoRawPointsGeometry() / getVideoRawPointsAttribute()          */
  const uint16_t* aux_geometry;  /* [F][aux_height][aux_width]     channel 0 of auxiliary geometry frame f       */
  const uint16_t* aux_attribute; /* [F][3][aux_height][aux_width]  auxiliary attribute frame f; NULL if none     */
} rb200_frames;

/* Decoder-native planes of one GOF: the planar 4:2:0 frames a video decoder (HM, libav, NVDEC) leaves behind, BEFORE
 * PCCVideoDecoder's inverse colour conversion (PccLibDecoder/source/PCCVideoDecoder.cpp:125-146, :365 ->
 * PCCInternalColorConverter<T>::convertYUV420ToYUV444, PccLibColorConverter/source/PCCInternalColorConverter.cpp:
 * 456-486), which then runs on the device.  host or device pointers. */
typedef struct rb200_frames_yuv420 {
  const uint8_t* occupancy;       /* [F][H/p][W/p]  as rb200_frames                                                    */
  const void*    geometry;        /* [F][M][H][W]   luma samples of geometry_sample_bytes each                        */
  const void*    attribute;       /* [F][M] frames { Y [H][W], U [H/2][W/2], V [H/2][W/2] } of attribute_sample_bytes */
  int32_t geometry_sample_bytes;  /* 1 or 2                                                                            */
  int32_t attribute_sample_bytes; /* 1 or 2                                                                            */
  int32_t attribute_bitdepth;     /* 8 or 10: the "<bits>" of "YUV420ToYUV444_<bits>_<filter>" (nbyte = 1 for 8)       */
  int32_t upsampling_filter;      /* "<filter>": index into g_filter420to444 (0..7), decoder parameter upsamplingFilter */
  /* PCCImage::set (PCCImage.h:97-138) as the decoder wrappers call it (PCCHMLibVideoDecoderImpl.cpp:360-363) with
   * shiftbits = internal bit depth - output bit depth: every sample becomes
   * clamp( (sample + (1 << (shift-1))) >> shift, 0, (1 << (10 - shift)) - 1 ); 0 = samples are copied as they are */
  int32_t geometry_shift;
  int32_t attribute_shift;
  /* PCCImage::convertBitdepth (PccLibCommon/source/PCCImage.cpp:258-299) as the decoder runs it on every decoded geometry
   * video (PCCDecoder.cpp:148-149, :173: decoder output bit depth -> gi.getGeometry2dBitdepthMinus1() + 1) and on the
   * occupancy video (PCCDecoder.cpp:119: 8 -> oi.getOccupancy2DBitdepthMinus1() + 1), fused into the ingest kernels:
   * diff = in - out >= 0: msb_align ? sample >> diff : min( sample, 2^out - 1 );  diff < 0: msb_align ? sample << -diff : sample.
   * bitdepth_out == 0: no conversion. */
  int32_t geometry_bitdepth_in, geometry_bitdepth_out, geometry_msb_align;
  int32_t occupancy_bitdepth_out, occupancy_msb_align; /* occupancy input depth is 8 */
} rb200_frames_yuv420;

/* Decoder surfaces of one GOF as a hardware decoder leaves them (libav AV_PIX_FMT_CUDA frames of NVDEC in RABBIT's
 * --useCuda path, PccLibTranscoder/source/PCCTranscoder.cpp:693-704, :791-817: data[0] = luma, data[1] = interleaved
 * chroma, linesize[] = pitch): one surface per video frame, pitched, NV12 (1-byte samples) or P010 / P016 (2-byte samples,
 * value in the high bits).  Pointers must be device memory or device-accessible (pinned) host memory: the planes are
 * gathered by a kernel, so a device-resident decode has NO host-to-device traffic for the planes. */
typedef struct rb200_surface {
  const void* luma;     /* rows of `pitch_luma` bytes                                                              */
  const void* chroma;   /* interleaved U, V rows (H/2 rows of W samples) of `pitch_chroma` bytes; NULL: luma only   */
  int32_t     pitch_luma, pitch_chroma;
} rb200_surface;
typedef struct rb200_frames_nv12 {
  const rb200_surface* occupancy; /* [F]     luma = occupancy video frame f, (W/p) x (H/p), always 1-byte samples    */
  const rb200_surface* geometry;  /* [F][M]  luma = geometry frame                                                   */
  const rb200_surface* attribute; /* [F][M]  luma + chroma of the 4:2:0 attribute frame; NULL if none                */
  int32_t sample_bytes;           /* geometry and attribute samples: 1 (NV12) or 2 (P010 / P016)                     */
  int32_t sample_lsb_shift;       /* 2-byte samples are shifted right by this on read (P010: 6)                      */
  /* the conversion parameters of rb200_frames_yuv420; its pointers and sample-byte fields are ignored */
  rb200_frames_yuv420 conversion;
} rb200_frames_nv12;

/* Patch tables of one GOF.  host pointers.  *_offset arrays have F+1 entries. */
typedef struct rb200_atlas {
  const rb200_patch*     patches;
  const int32_t*         patch_offset;
  const rb200_eom_patch* eom_patches;   /* may be NULL */
  const int32_t*         eom_offset;    /* may be NULL */
  const int32_t*         eom_members;   /* may be NULL */
  const rb200_raw_patch* raw_patches;   /* may be NULL */
  const int32_t*         raw_offset;    /* may be NULL */
} rb200_atlas;

/* Host-side destination of rb200_download_frame.  Any pointer may be NULL (skipped).  Byte layouts are
 * the reference's own std::vector element layouts so the shim can memcpy straight into PCCPointSet3
 * (PCCPointSet.h:520-531, PCCMath.h:449-455). */
typedef struct rb200_cloud_host {
  int16_t*  positions;      /* [N][3]   positions_                                                       */
  uint16_t* colors16;       /* [N][3]   colors16bit_                                                     */
  uint8_t*  colors;         /* [N][3]   colors_   (valid after rb200_convert_rgb8)                       */
  uint16_t* boundary_types; /* [N]      boundaryPointTypes_                                              */
  uint32_t* partition;      /* [N]      partition[] (patch index; EOM/raw points: patch count)           */
  uint32_t* point_to_pixel; /* [N][3]   tile.getPointToPixel(): x, y, layer                              */
} rb200_cloud_host;

typedef struct rb200_frame_counts {
  int64_t total, regular, eom, raw; /* setTotalNumberOf{Regular,EOM,Raw}Points, PCCCodec.cpp:841,886,948 */
  int64_t smoothed;                 /* points moved by geometry smoothing (type 3)                       */
  int64_t recolored;                /* points changed by colour smoothing                                */
} rb200_frame_counts;

typedef struct rb200_ctx rb200_ctx;

/* ---- context ---------------------------------------------------------------------------------- */
int         rb200_abi_version(void);
int         rb200_create(int cuda_device, rb200_ctx** out);
void        rb200_destroy(rb200_ctx* ctx);
const char* rb200_error_string(const rb200_ctx* ctx); /* last error text of this context            */
int         rb200_set_stream(rb200_ctx* ctx, void* cuda_stream /* cudaStream_t, NULL = own stream */);
int         rb200_synchronize(rb200_ctx* ctx);

/* pinned (page-locked) host memory for staging planes and results: copies to / from it are true DMA transfers */
void*       rb200_host_alloc(size_t bytes);
void        rb200_host_free(void* p);

/* ---- PCCCodec::generateOccupancyMap (PccLibCommon/source/PCCCodec.cpp:1584-1606) for ONE frame, as the decoder calls it
 *      before it knows the GOF (PCCDecoder.cpp:361-365): `video` [oH][oW] (host) is thresholded IN PLACE exactly as the
 *      reference does it (once per full-resolution pixel, i.e. precision^2 times per sample, :1597-1600), `map_out`
 *      [oH * precision][oW * precision] uint32 (host) is tile.getOccupancyMap(). ----------------------------------------- */
int rb200_occupancy_map(rb200_ctx* ctx, uint8_t* video, int o_width, int o_height, int precision, int threshold_lossy_om,
                        int enhanced_occupancy_map, uint32_t* map_out);

/* ---- frame ingest: replaces PCCImage::set / PCCVideo containers on the path (PCCImage.h:97-138) --- */
int rb200_gof_begin(rb200_ctx* ctx, const rb200_params* params, int n_frames);
int rb200_gof_upload(rb200_ctx* ctx, const rb200_frames* frames, const rb200_atlas* atlas);

/* Point local reconstruction (params.point_local_reconstruction): the mode table the decoder builds in
 * PCCDecoder::setPointLocalReconstruction (PccLibDecoder/source/PCCDecoder.cpp:528-550; entry 0 is {0,0,0,1}) and, for
 * every block of every patch, the mode PCCPatch::getPointLocalReconstructionMode(u0, v0) resolves to
 * (PCCPatch.h:283-289, filled by PCCDecoder::setPLRData :552-591).  Call after rb200_gof_upload, before
 * rb200_reconstruct.  Host pointers; copied. */
typedef struct rb200_plr_mode {
  uint8_t interpolate, filling, min_d1, neighbor; /* PointLocalReconstructionMode, PCCPLRInformation.h:40-45 */
} rb200_plr_mode;
typedef struct rb200_plr {
  int32_t               n_modes;
  const rb200_plr_mode* modes;        /* [n_modes]                                                              */
  const uint8_t*        block_mode;   /* patch i (atlas order) owns block_mode[block_offset[i] + v0 * size_u0 + u0] */
  const int64_t*        block_offset; /* [total patches + 1]                                                    */
} rb200_plr;
int rb200_gof_set_plr(rb200_ctx* ctx, const rb200_plr* plr);
/* the same with decoder-native planes: replaces PCCImage::set (PCCImage.h:97-138) + the inverse colour conversion of
 * PCCVideoDecoder (PCCVideoDecoder.cpp:125-146, :365; PCCInternalColorConverter.cpp:456-486, :596-611, :669-695, :582-594) */
int rb200_gof_upload_yuv420(rb200_ctx* ctx, const rb200_frames_yuv420* frames, const rb200_atlas* atlas);
/* the same from pitched NV12 / P010 decoder surfaces (host arrays of rb200_surface; the surfaces themselves in device or
 * pinned memory) */
int rb200_gof_upload_nv12(rb200_ctx* ctx, const rb200_frames_nv12* frames, const rb200_atlas* atlas);
/* the planes the reconstruction reads as they sit in HBM after an upload: geometry [H][W], attribute [3][H][W] uint16 of
 * frame `frame`, map `map` (either pointer may be NULL); used to check the ingest conversion */
int rb200_download_planes(rb200_ctx* ctx, int frame, int map, uint16_t* geometry, uint16_t* attribute);

/* ---- reconstruction: PCCCodec::generateOccupancyMap (PCCCodec.cpp:1584-1606) +
 *      generateBlockToPatchFromOccupancyMapVideo (:1725-1763) + generatePointCloud (:517-978, incl.
 *      generatePoints :327-515, EOM :669-779/:846-891, raw :894-949, identifyBoundaryPoints :266-325) +
 *      colorPointCloud (:1308-1449), for every frame of the GOF in one batched launch sequence. ------ */
int rb200_reconstruct(rb200_ctx* ctx);

/* ---- PCCCodec::smoothPointCloudPostprocess (:52-147) + smoothPointCloudGrid/gridFiltering (:1000-1104) */
int rb200_smooth_geometry(rb200_ctx* ctx);

/* ---- PCCPointSet3::transferColors16bitBP as called at PCCDecoder.cpp:447-465 (PCCPointSet.cpp:1126-1485) */
int rb200_transfer_colors(rb200_ctx* ctx);

/* ---- PCCCodec::colorSmoothing (:149-236) + gridFilteringColor / smoothPointCloudColorLC (:1182-1306) */
int rb200_smooth_color(rb200_ctx* ctx);

/* ---- PCCPointSet3::convertYUV16ToRGB8 / copyRGB16ToRGB8 (PCCPointSet.h:121-166) ------------------ */
int rb200_convert_rgb8(rb200_ctx* ctx);
/* test hook: n host colour triples [n][3] through the conversion kernel of rb200_convert_rgb8 (a short
 * evaluation of convertYUV16ToRGB8 that falls back to the reference's sequence of double operations near rounding ties), or,
 * force_f64 != 0, through the double arithmetic alone; rgb [n][3] on the host */
int rb200_debug_yuv16_to_rgb8(rb200_ctx* ctx, const uint16_t* yuv, int64_t n, uint8_t* rgb, int force_f64);

/* test hook: the sparse smoothing tables of this context start 2^shrink times smaller (and the luma lists at 4 entries),
 * so that the overflow -> regrow -> repeat path runs on small inputs; 0 restores the production sizes */
int rb200_debug_set_grid_shrink(rb200_ctx* ctx, int shrink);

/* ---- the decoder's whole per-frame sequence (PCCDecoder.cpp:330-508) governed by params ---------- */
int rb200_decode_gof(rb200_ctx* ctx);

/* ---- results ---------------------------------------------------------------------------------- */
int rb200_frame_counts_get(rb200_ctx* ctx, rb200_frame_counts* out /* [n_frames] */);
int rb200_download_frame(rb200_ctx* ctx, int frame, const rb200_cloud_host* dst);
/* A frame-by-frame caller (the reference's decoder loop) sees the state every stage left although each stage runs for
 * the whole GOF at once: enable before rb200_reconstruct, then ask for the state after stage 0 reconstruction + colour
 * fetch, 1 geometry smoothing, 2 attribute re-transfer, 3 colour smoothing, >= 4 current. */
int rb200_enable_stage_snapshots(rb200_ctx* ctx, int enable);
int rb200_download_frame_stage(rb200_ctx* ctx, int frame, int stage, const rb200_cloud_host* dst);
/* all frames of the GOF back to back in frame order (frame f = counts[f].total points): one packed copy per field */
int rb200_download_gof(rb200_ctx* ctx, const rb200_cloud_host* dst);
/* block-to-patch map (tile.getBlockToPatch(), value = patch index + 1, 0 = none), [H/R][W/R] uint32 */
int rb200_download_block_to_patch(rb200_ctx* ctx, int frame, uint32_t* dst);
/* full-resolution occupancy map (tile.getOccupancyMap()), [H][W] uint8, after EOM marks */
int rb200_download_occupancy(rb200_ctx* ctx, int frame, uint8_t* dst);

/* ---- metrics: PCCMetrics::compute (PccLibMetrics/source/PCCMetrics.cpp:334-385) ------------------ */
typedef struct rb200_metrics_params { /* PCCMetricsParameters fields the path reads */
  int32_t compute_c2c, compute_c2p, compute_color, compute_hausdorff;
  int32_t drop_duplicates;  /* 0 keep, 1 drop, 2 average colours (default 2)                            */
  int32_t neighbors_proc;   /* 0 first NN, 1/2 average of the tie set (default 1)                       */
  float   resolution;       /* PSNR peak (1023 vox10, 2047 vox11)                                       */
  int32_t reserved;
} rb200_metrics_params;

typedef struct rb200_cloud_view { /* host or device pointers */
  const int16_t* positions; /* [n][3]                                                                  */
  const uint8_t* colors;    /* [n][3] RGB8 or NULL                                                      */
  const float*   normals;   /* [n][3] or NULL (D2 needs them on the source, PCCMetrics.cpp:371-375)      */
  int64_t        count;
} rb200_cloud_view;

/* One direction A->B of QualityMetrics::compute (PCCMetrics.cpp:75-231): the double accumulators and
 * the float results derived from them exactly as :204-226 does. */
typedef struct rb200_quality {
  double  sse_c2c, sse_c2p, sse_color[3], max_c2c, max_c2p;
  int64_t num;
  float   c2c_mse, c2c_psnr, c2p_mse, c2p_psnr, c2c_hausdorff, c2c_hausdorff_psnr, c2p_hausdorff,
      c2p_hausdorff_psnr, color_mse[3], color_psnr[3];
} rb200_quality;

typedef struct rb200_metrics_result {
  rb200_quality q1, q2, qf;   /* quality1_ (src->rec), quality2_ (rec->src), qualityF_ (:299-332)        */
  int64_t source_points, source_after_dedup, rec_points, rec_after_dedup;
  int32_t tie_overflow;       /* >0: some tie set was larger than 30 (reference result order-dependent)  */
  int32_t reserved;
} rb200_metrics_result;

/* metrics for n_pairs (source, reconstruction[, source-with-normals]) pairs in one batched launch
 * sequence.  reconstruct[i].positions == NULL means "frame i of the GOF resident in this context". */
int rb200_metrics(rb200_ctx* ctx, const rb200_metrics_params* params, int n_pairs,
                  const rb200_cloud_view* sources, const rb200_cloud_view* reconstructs,
                  rb200_metrics_result* results /* [n_pairs] */);

/* Source clouds that stay in DEVICE memory across calls (the transcode loop compares the same source frames with the
 * reconstruction of every rate point, /root/reference transcode.sh:25-37 → PccAppMetrics per rate): with on != 0,
 * a call whose sources are the same device pointers / counts, with the same drop_duplicates and normals, as the previous
 * call keeps their de-duplicated points, column tables and gathered normals (removeDuplicate + copyNormals of the
 * source, PCCMetrics.cpp:353-375) and only indexes the reconstructions.  The caller promises the buffers' contents did
 * not change in between; calling this again (on or off) drops what is kept.  The results are the same either way:
 * integer quantities exactly, the double sums to the last bits that the order of the normals' atomic additions leaves
 * open from run to run in any case (PSNR far inside 1e-6 dB). */
int rb200_metrics_cache_sources(rb200_ctx* ctx, int on);

/* ---- multi-GPU: the one exchange step of the path.  Frames / GOFs / streams are sharded over one context
 *      per GPU with no data-path collective; what ranks exchange is one fixed-size record of accumulators per frame.
 *      A C / C++ host packs its results, all-gathers the records with its own communicator (ncclAllGather / MPI_Allgather
 *      on RB200_METRICS_RECORD doubles per frame) and unpacks: the float results are derived again from the sums exactly as
 *      QualityMetrics::compute (PCCMetrics.cpp:204-226) and operator+ (:299-332) do, so they equal the owning rank's.
 *      (rabbit_transcoding_b200.dist.gather_metrics is the torch.distributed form of the same.) -------------------------- */
#define RB200_METRICS_RECORD 24
int rb200_metrics_pack(int frame, const rb200_metrics_result* result, double* record /* [RB200_METRICS_RECORD] */);
int rb200_metrics_unpack(const double* record, const rb200_metrics_params* params, int* frame, rb200_metrics_result* out);

/* PCCPointSet3::removeDuplicate(out, dropDuplicates) (PCCPointSet.cpp:169-218): lexicographic sort +
 * merge.  Returns the number of output points; out_* may be NULL to only count. */
int rb200_remove_duplicates(rb200_ctx* ctx, const rb200_cloud_view* in, int drop_duplicates,
                            int16_t* out_positions, uint8_t* out_colors, int64_t* out_count);

/* ---- PCCKdTree::search (PccLibCommon/source/PCCKdTree.cpp:61-66): k nearest neighbours of every query in `cloud`
 *      in nanoflann's result order (ties in tree-traversal order), k = 1..8.  host or device pointers.
 *      out_idx / out_dist are [nq][k]; missing entries (cloud smaller than k) are -1. ------------------------------- */
int rb200_kdtree_search(rb200_ctx* ctx, const int16_t* cloud, int64_t n, const int16_t* queries, int64_t nq, int k,
                        int64_t* out_idx, double* out_dist);
/* PCCKdTree::searchRadius (PCCKdTree.cpp:69-79): nanoflann radiusSearch (dist < radius2), std::sort by distance and, for
 * equal distances, by index (the vendored IndexDist_Sorter, nanoflann.hpp:193-200), cut to max_results — what the non-grid smoothPointCloud asks per point (PCCCodec.cpp:1120).  sorted == 0 returns the unsorted
 * traversal order instead (nanoflann SearchParams::sorted = false).  out_idx / out_dist are [nq][max_results] (missing
 * entries -1), out_count [nq] is the number of points inside the radius before the cut. */
int rb200_kdtree_search_radius(rb200_ctx* ctx, const int16_t* cloud, int64_t n, const int16_t* queries, int64_t nq,
                               double radius2, int max_results, int sorted, int64_t* out_idx, double* out_dist,
                               int32_t* out_count);

/* ---- PCCPointSet3::computeChecksum( false ) (PCCPointSet.cpp:222-245): MD5 of positions || RGB8 of a decoded frame ---- */
int rb200_frame_md5(rb200_ctx* ctx, int frame, uint8_t* out16);
/* PCCPointSet3::computeChecksum( true ) (:221-229, reorder :258-296): MD5 after the canonical reordering — points in
 * lexicographic (x, y, z) order, one per position, colour = integer mean of its duplicates (the conformance checksum
 * of PCCDecoder.cpp:420; north_star's "bit-exact after canonical sorting") */
int rb200_frame_md5_canonical(rb200_ctx* ctx, int frame, uint8_t* out16);
/* ---- PCCPointSet3::write( file, asAscii = false ) (:359-457): binary little-endian PLY, float xyz + uchar rgb,
 *      byte-identical to the reference's file for a decoded frame.  Records are packed on the device. ------------------- */
int rb200_write_ply(rb200_ctx* ctx, int frame, const char* path);
/* ---- PCCPointSet3::read (:459-757) for x, y, z (+ red, green, blue): ascii or binary little-endian.  out_pos NULL or
 *      capacity too small: only *out_count (and *has_colors) are filled. ------------------------------------------------ */
int rb200_read_ply(const char* path, int16_t* out_pos, uint8_t* out_col, int64_t capacity, int64_t* out_count, int* has_colors);

/* ---- instrumentation --------------------------------------------------------------------------- */
typedef struct rb200_launch_stats {
  int64_t kernel_launches;    /* number of this library's kernels launched since the last reset        */
  int64_t h2d_bytes, d2h_bytes;
} rb200_launch_stats;
int rb200_stats_get(rb200_ctx* ctx, rb200_launch_stats* out, int reset);
/* per-kernel CUDA-event timing: enable, run, then read name/ms pairs (for bench.py's roofline) */
int rb200_timing_enable(rb200_ctx* ctx, int enable);
int rb200_timing_get(rb200_ctx* ctx, int index, char* name, int name_cap, double* total_ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* RABBIT_B200_H */
