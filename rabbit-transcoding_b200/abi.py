"""ctypes mirror of include/rabbit_b200.h and loader of the CUDA library.

The structures here are field-for-field the C structs of include/rabbit_b200.h; they are also what the
test-only CPU checkers consume, so the CUDA path, the CPU
restatement and the unmodified reference all see byte-identical inputs.

There is NO CPU fallback: `load_library()` raises if librabbit_b200.so is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "librabbit_b200.so")

RB200_OK = 0
RB200_ERR_INVALID = 1
RB200_ERR_UNSUPPORTED = 2
RB200_ERR_CUDA = 3
RB200_ERR_NOMEM = 4
RB200_ERR_STATE = 5
RB200_ERR_TIE_OVERFLOW = 6
RB200_ERR_PATCH_OUT_OF_CANVAS = 180

i32 = C.c_int32
i64 = C.c_int64


class Patch(C.Structure):
    """rb200_patch (PCCPatch.h:353-408)"""
    _fields_ = [(n, i32) for n in (
        "u0", "v0", "size_u0", "size_v0", "u1", "v1", "d1",
        "normal_axis", "tangent_axis", "bitangent_axis", "projection_mode", "orientation",
        "lod_x", "lod_y", "axis_of_additional_plane", "size2d_x_px", "size2d_y_px")]


PATCH_DTYPE = np.dtype([(n, np.int32) for n, _ in Patch._fields_])
assert PATCH_DTYPE.itemsize == C.sizeof(Patch)


class EomPatch(C.Structure):
    _fields_ = [(n, i32) for n in ("u0", "v0", "member_begin", "member_count", "eom_count")]


EOM_DTYPE = np.dtype([(n, np.int32) for n, _ in EomPatch._fields_])


class RawPatch(C.Structure):
    _fields_ = [(n, i32) for n in ("u0", "v0", "size_u0", "size_v0", "u1", "v1", "d1", "num_points")]


RAW_DTYPE = np.dtype([(n, np.int32) for n, _ in RawPatch._fields_])


class Params(C.Structure):
    """rb200_params (GeneratePointCloudParameters, PCCCodec.h:62-100)"""
    _fields_ = [(n, i32) for n in (
        "width", "height", "occupancy_resolution", "occupancy_precision", "threshold_lossy_om",
        "map_count_minus1", "absolute_d1", "remove_duplicate_points", "enhanced_occupancy_map_code",
        "eom_fix_bit_count", "enable_size_quantization", "log2_quantizer_x", "log2_quantizer_y",
        "patch_precedence_reverse", "use_additional_points_patch", "total_raw_points_known",
        "single_map_pixel_interleaving", "point_local_reconstruction", "pbf_enable", "multiple_streams",
        "attribute_count", "attribute_rgb444", "geometry_bitdepth_3d",
        "flag_geometry_smoothing", "grid_smoothing", "grid_size", "apply_geo_smoothing",
        "attr_transfer_filter_type", "flag_color_smoothing", "apply_attr_smoothing", "relative_t1",
        "surface_thickness")] + [
        (n, C.c_double) for n in (
            "threshold_smoothing", "threshold_color_smoothing", "threshold_color_difference",
            "threshold_color_variation")] + [
        (n, i32) for n in ("pbf_passes_count", "pbf_filter_size", "pbf_log2_threshold", "neighbor_count_smoothing")] + [
        (n, C.c_double) for n in ("radius2_smoothing", "radius2_boundary_detection")] + [
        (n, i32) for n in ("use_aux_separate_video", "aux_width", "aux_height", "reserved0")]


class PlrMode(C.Structure):
    """rb200_plr_mode (PointLocalReconstructionMode, PCCPLRInformation.h:40-45)"""
    _fields_ = [("interpolate", C.c_uint8), ("filling", C.c_uint8), ("min_d1", C.c_uint8), ("neighbor", C.c_uint8)]


class Plr(C.Structure):
    _fields_ = [("n_modes", i32), ("modes", C.c_void_p), ("block_mode", C.c_void_p), ("block_offset", C.c_void_p)]


class Frames(C.Structure):
    _fields_ = [("occupancy", C.c_void_p), ("geometry", C.c_void_p), ("attribute", C.c_void_p),
                ("aux_geometry", C.c_void_p), ("aux_attribute", C.c_void_p)]


class FramesYuv420(C.Structure):
    _fields_ = [("occupancy", C.c_void_p), ("geometry", C.c_void_p), ("attribute", C.c_void_p),
                ("geometry_sample_bytes", i32), ("attribute_sample_bytes", i32), ("attribute_bitdepth", i32),
                ("upsampling_filter", i32), ("geometry_shift", i32), ("attribute_shift", i32),
                ("geometry_bitdepth_in", i32), ("geometry_bitdepth_out", i32), ("geometry_msb_align", i32),
                ("occupancy_bitdepth_out", i32), ("occupancy_msb_align", i32)]


class Surface(C.Structure):
    """rb200_surface: one pitched decoder surface (luma + interleaved chroma)"""
    _fields_ = [("luma", C.c_void_p), ("chroma", C.c_void_p), ("pitch_luma", i32), ("pitch_chroma", i32)]


class FramesNv12(C.Structure):
    _fields_ = [("occupancy", C.c_void_p), ("geometry", C.c_void_p), ("attribute", C.c_void_p),
                ("sample_bytes", i32), ("sample_lsb_shift", i32), ("conversion", FramesYuv420)]


class Atlas(C.Structure):
    _fields_ = [("patches", C.c_void_p), ("patch_offset", C.c_void_p),
                ("eom_patches", C.c_void_p), ("eom_offset", C.c_void_p), ("eom_members", C.c_void_p),
                ("raw_patches", C.c_void_p), ("raw_offset", C.c_void_p)]


class CloudHost(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("colors16", C.c_void_p), ("colors", C.c_void_p),
                ("boundary_types", C.c_void_p), ("partition", C.c_void_p), ("point_to_pixel", C.c_void_p)]


class FrameCounts(C.Structure):
    _fields_ = [(n, i64) for n in ("total", "regular", "eom", "raw", "smoothed", "recolored")]


class MetricsParams(C.Structure):
    _fields_ = [("compute_c2c", i32), ("compute_c2p", i32), ("compute_color", i32), ("compute_hausdorff", i32),
                ("drop_duplicates", i32), ("neighbors_proc", i32), ("resolution", C.c_float), ("reserved", i32)]


class CloudView(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("colors", C.c_void_p), ("normals", C.c_void_p), ("count", i64)]


class Quality(C.Structure):
    _fields_ = [("sse_c2c", C.c_double), ("sse_c2p", C.c_double), ("sse_color", C.c_double * 3),
                ("max_c2c", C.c_double), ("max_c2p", C.c_double), ("num", i64),
                ("c2c_mse", C.c_float), ("c2c_psnr", C.c_float), ("c2p_mse", C.c_float), ("c2p_psnr", C.c_float),
                ("c2c_hausdorff", C.c_float), ("c2c_hausdorff_psnr", C.c_float),
                ("c2p_hausdorff", C.c_float), ("c2p_hausdorff_psnr", C.c_float),
                ("color_mse", C.c_float * 3), ("color_psnr", C.c_float * 3)]


class MetricsResult(C.Structure):
    _fields_ = [("q1", Quality), ("q2", Quality), ("qf", Quality),
                ("source_points", i64), ("source_after_dedup", i64), ("rec_points", i64), ("rec_after_dedup", i64),
                ("tie_overflow", i32), ("reserved", i32)]


class LaunchStats(C.Structure):
    _fields_ = [("kernel_launches", i64), ("h2d_bytes", i64), ("d2h_bytes", i64)]


# every symbol include/rabbit_b200.h declares; tests check the .so exports all of them
EXPORTED_SYMBOLS = [
    "rb200_abi_version", "rb200_create", "rb200_destroy", "rb200_error_string", "rb200_set_stream",
    "rb200_synchronize", "rb200_host_alloc", "rb200_host_free", "rb200_occupancy_map", "rb200_gof_begin", "rb200_gof_upload", "rb200_gof_set_plr", "rb200_gof_upload_yuv420", "rb200_gof_upload_nv12", "rb200_download_planes", "rb200_reconstruct", "rb200_smooth_geometry",
    "rb200_transfer_colors", "rb200_smooth_color", "rb200_convert_rgb8", "rb200_debug_yuv16_to_rgb8", "rb200_debug_set_grid_shrink", "rb200_decode_gof",
    "rb200_frame_counts_get", "rb200_download_frame", "rb200_download_gof",
    "rb200_enable_stage_snapshots", "rb200_download_frame_stage", "rb200_download_block_to_patch",
    "rb200_download_occupancy", "rb200_metrics", "rb200_metrics_cache_sources", "rb200_metrics_pack", "rb200_metrics_unpack", "rb200_remove_duplicates", "rb200_kdtree_search", "rb200_kdtree_search_radius", "rb200_frame_md5", "rb200_frame_md5_canonical", "rb200_write_ply", "rb200_read_ply", "rb200_stats_get",
    "rb200_timing_enable", "rb200_timing_get",
]

# stages of the path that this build implements on the GPU (bench.py / tests pick their configs from these)
HAVE_TRANSFER = True   # rb200_transfer_colors (PCCPointSet3::transferColors16bitBP)
HAVE_METRICS = True    # rb200_metrics / rb200_remove_duplicates

_lib = None


def load_library(path=None):
    """dlopen librabbit_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"rabbit_b200: CUDA library {p} not built; run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback.")
    lib = C.CDLL(p)
    lib.rb200_error_string.restype = C.c_char_p
    lib.rb200_error_string.argtypes = [C.c_void_p]
    lib.rb200_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.rb200_destroy.argtypes = [C.c_void_p]
    lib.rb200_destroy.restype = None
    lib.rb200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.rb200_synchronize.argtypes = [C.c_void_p]
    lib.rb200_host_alloc.argtypes = [C.c_size_t]
    lib.rb200_host_alloc.restype = C.c_void_p
    lib.rb200_host_free.argtypes = [C.c_void_p]
    lib.rb200_host_free.restype = None
    lib.rb200_occupancy_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.rb200_gof_begin.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int]
    lib.rb200_gof_upload.argtypes = [C.c_void_p, C.POINTER(Frames), C.POINTER(Atlas)]
    lib.rb200_gof_set_plr.argtypes = [C.c_void_p, C.POINTER(Plr)]
    lib.rb200_gof_upload_yuv420.argtypes = [C.c_void_p, C.POINTER(FramesYuv420), C.POINTER(Atlas)]
    lib.rb200_gof_upload_nv12.argtypes = [C.c_void_p, C.POINTER(FramesNv12), C.POINTER(Atlas)]
    lib.rb200_download_planes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    for n in ("rb200_reconstruct", "rb200_smooth_geometry", "rb200_transfer_colors", "rb200_smooth_color",
              "rb200_convert_rgb8", "rb200_decode_gof"):
        getattr(lib, n).argtypes = [C.c_void_p]
    lib.rb200_debug_yuv16_to_rgb8.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, C.c_int]
    lib.rb200_debug_set_grid_shrink.argtypes = [C.c_void_p, C.c_int]
    lib.rb200_frame_counts_get.argtypes = [C.c_void_p, C.POINTER(FrameCounts)]
    lib.rb200_download_frame.argtypes = [C.c_void_p, C.c_int, C.POINTER(CloudHost)]
    lib.rb200_enable_stage_snapshots.argtypes = [C.c_void_p, C.c_int]
    lib.rb200_download_frame_stage.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(CloudHost)]
    lib.rb200_download_gof.argtypes = [C.c_void_p, C.POINTER(CloudHost)]
    lib.rb200_download_block_to_patch.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.rb200_download_occupancy.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.rb200_metrics.argtypes = [C.c_void_p, C.POINTER(MetricsParams), C.c_int, C.POINTER(CloudView),
                                  C.POINTER(CloudView), C.POINTER(MetricsResult)]
    lib.rb200_metrics_cache_sources.argtypes = [C.c_void_p, C.c_int]
    lib.rb200_metrics_pack.argtypes = [C.c_int, C.POINTER(MetricsResult), C.POINTER(C.c_double)]
    lib.rb200_metrics_unpack.argtypes = [C.POINTER(C.c_double), C.POINTER(MetricsParams), C.POINTER(C.c_int), C.POINTER(MetricsResult)]
    lib.rb200_remove_duplicates.argtypes = [C.c_void_p, C.POINTER(CloudView), C.c_int, C.c_void_p, C.c_void_p,
                                            C.POINTER(i64)]
    lib.rb200_kdtree_search.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, i64, C.c_int, C.c_void_p, C.c_void_p]
    lib.rb200_kdtree_search_radius.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, i64, C.c_double, C.c_int, C.c_int, C.c_void_p,
                                               C.c_void_p, C.c_void_p]
    lib.rb200_frame_md5.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.rb200_frame_md5_canonical.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.rb200_write_ply.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    lib.rb200_read_ply.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, i64, C.POINTER(i64), C.POINTER(C.c_int)]
    lib.rb200_stats_get.argtypes = [C.c_void_p, C.POINTER(LaunchStats), C.c_int]
    lib.rb200_timing_enable.argtypes = [C.c_void_p, C.c_int]
    lib.rb200_timing_get.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_double),
                                     C.POINTER(i64)]
    if path is None:
        _lib = lib
    return lib


def ptr(a):
    """raw address of a numpy array / torch tensor / None"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return int(a)


def plr_struct(plr):
    """rb200_plr over the arrays of a dict(modes=uint8[n, 4], block_mode=uint8[], block_offset=int64[])"""
    import numpy as np
    assert plr["modes"].dtype == np.uint8 and plr["modes"].ndim == 2 and plr["modes"].shape[1] == 4
    assert plr["block_mode"].dtype == np.uint8 and plr["block_offset"].dtype == np.int64
    s = Plr()
    s.n_modes = plr["modes"].shape[0]
    s.modes, s.block_mode, s.block_offset = ptr(plr["modes"]), ptr(plr["block_mode"]), ptr(plr["block_offset"])
    return s
