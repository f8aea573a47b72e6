"""GPU parity: the CUDA path through the C ABI vs the CPU checker (the unmodified reference when
oracle/_ref/librabbit_ref.so is present, else the C restatement), bit-exact at every stage."""
import numpy as np
import pytest

from util import run_stages

pytestmark = pytest.mark.gpu


def small(rb, **kw):
    args = dict(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=11, transfer_filter=0)
    args.update(kw)
    return rb.synthetic.generate_gof(**args)


def test_default_two_frames(rb, codec, checker_backend):
    run_stages(codec, small(rb), checker_backend, what="default")


def test_all_orientations_p2(rb, codec, checker_backend):
    run_stages(codec, small(rb, orientations=tuple(range(9)), occupancy_precision=2, seed=12), checker_backend,
               what="orient")


def test_precision1_single_map(rb, codec, checker_backend):
    run_stages(codec, small(rb, occupancy_precision=1, map_count=1, seed=13), checker_backend, what="p1m1")


def test_patch_precedence_reverse(rb, codec, checker_backend):
    run_stages(codec, small(rb, precedence_reverse=True, seed=14), checker_backend, what="reverse")


def test_relative_d1_keep_duplicates(rb, codec, checker_backend):
    g = small(rb, absolute_d1=False, seed=15)
    g.params.remove_duplicate_points = 0
    run_stages(codec, g, checker_backend, what="reld1")


def test_level_of_detail_and_45_degree_planes(rb, codec, checker_backend):
    """lod 2 patches (PCCPatch.h:201-207) and all three 45-degree projection planes (PCCCodec.cpp:2503-2524), including
    the points whose intermediate sums are negative and wrap to 0 in the reference"""
    for seed, kw in ((61, dict(lod=True, planes=False)), (62, dict(lod=False, planes=True)), (63, dict(lod=True, planes=True))):
        g = rb.synthetic.make_lod_and_oblique(small(rb, seed=seed, orientations=tuple(range(9))), seed=seed, **kw)
        ref = run_stages(codec, g, checker_backend, what=f"lod/oblique {kw}")
        if kw["planes"]:
            axes = set(int(a) for a in g.patches["axis_of_additional_plane"])
            assert {1, 2, 3} <= axes
            pos = ref.cloud(0, "reconstruct")["positions"]
            assert (pos == 0).any(axis=1).sum() > 0  # some intermediate went negative: the wrap-to-0 case is exercised
    # the same through the whole decoder sequence with the attribute re-transfer
    g = rb.synthetic.make_lod_and_oblique(small(rb, seed=64, transfer_filter=1), seed=64)
    run_stages(codec, g, checker_backend, stages=ALL_STAGES, what="lod/oblique rec-1")


def test_non_grid_smoothing(rb, codec, checker_backend):
    """smoothPointCloudPostprocess with gridSmoothing_ == 0 (PCCCodec.cpp:141 -> smoothPointCloud, :1106-1157; the encoder's
    reconstruction): radius search, sorted by distance and index, the first 64; most neighbourhoods hold more than 64 points,
    so the cut falls inside a shell of equal distances"""
    for seed, kw in ((71, dict()), (72, dict(occupancy_precision=2, orientations=tuple(range(9)))), (73, dict(map_count=1))):
        g = small(rb, seed=seed, **kw)
        g.params.grid_smoothing = 0
        g.params.neighbor_count_smoothing = 64  # PCCEncoderParameters.cpp:92-94
        g.params.radius2_smoothing = 64.0
        g.params.radius2_boundary_detection = 64.0
        ref = run_stages(codec, g, checker_backend, what=f"non-grid {kw}")
        a, b = ref.cloud(0, "reconstruct"), ref.cloud(0, "smooth_geometry")
        assert (a["positions"] != b["positions"]).any(axis=1).sum() > 20     # points really move
        assert (b["boundary_types"] == 2).sum() > 20                        # and boundary points are re-typed (:1131-1133)
    # other radii / counts, and the whole decoder-style sequence behind it (no point is of type 3: the re-transfer is a no-op)
    g = small(rb, seed=74, transfer_filter=1)
    g.params.grid_smoothing = 0
    g.params.neighbor_count_smoothing = 20
    g.params.radius2_smoothing = 30.5
    g.params.radius2_boundary_detection = 9.0
    g.params.threshold_smoothing = 3.0
    run_stages(codec, g, checker_backend, stages=ALL_STAGES, what="non-grid r2 30.5")
    # decodeGof takes the same branch
    codec.uploadGof(g)
    codec.decodeGof()
    ref = checker_backend.run_gof(g, keep=("rgb8",))
    for f in range(g.n_frames):
        got = codec.getPointCloud(f, fields=("positions", "colors"))
        want = ref.cloud(f, "rgb8")
        assert np.array_equal(got["positions"], want["positions"]) and np.array_equal(got["colors"], want["colors"])


def test_cell_sums_beyond_the_exact_float_range(rb, codec, checker_backend):
    """twelve patches decode to the same slab with a saturated luma: the colour cells hold a few hundred points and their
    float sums pass 2^24, where the reference's result depends on the ORDER of its float additions.
    Those cells are re-accumulated in emission order on the device (k_ordered_cells) and must still match bit for bit."""
    g = rb.synthetic.generate_stacked_gof(n_patches=12, n_frames=2, seed=5)
    ref = run_stages(codec, g, checker_backend, what="stacked")
    assert ref.counts(0).recolored > 0 and ref.counts(0).total > 12 * 32 * 32
    # the construction really leaves the exact range: one 4^3 cell holds far more than 2^24 / 65535 = 256 points
    pos = ref.cloud(0, "smooth_geometry")["positions"].astype(np.int64) // 4
    key = (pos[:, 0] << 32) | (pos[:, 1] << 16) | pos[:, 2]
    assert np.unique(key, return_counts=True)[1].max() > 300
    g = rb.synthetic.generate_stacked_gof(n_patches=12, n_frames=1, seed=6, transfer_filter=1)
    run_stages(codec, g, checker_backend, stages=ALL_STAGES, what="stacked rec-1")


def pbf(g, passes=None, filter_size=None, log2_threshold=2, grid_smoothing=True):
    """switch a GOF to Rec-2 occupancy synthesis with the encoder's defaults (PCCEncoderParameters.cpp:1132-1133)"""
    p = g.params
    prec = p.occupancy_precision
    p.pbf_enable = 1
    p.pbf_passes_count = passes or (1 if prec <= 2 else 2 if prec == 4 else 4)
    p.pbf_filter_size = filter_size or prec
    p.pbf_log2_threshold = log2_threshold
    p.flag_geometry_smoothing = 1  # the occupancy synthesis SEI sets it (PCCDecoder.cpp:631)
    p.grid_smoothing = 1 if grid_smoothing else 0
    return g


def test_occupancy_synthesis_pbf(rb, codec, checker_backend):
    """Rec-2: PatchBlockFiltering::patchBorderFiltering (PCCPatch.cpp:797-977) trims the block-upsampled occupancy of
    every patch against the border points of its 3-D neighbours; the points' boundary type is PCCPatch::isBorder"""
    if not hasattr(checker_backend, "run_gof") or type(checker_backend).__name__ != "Reference":
        pytest.skip("occupancy synthesis is pinned against the compiled reference only")
    for seed, kw, pk in ((71, dict(), dict()),
                         (72, dict(orientations=tuple(range(9)), occupancy_precision=2), dict()),
                         (73, dict(orientations=tuple(range(9)), map_count=1), dict(passes=3, filter_size=3, log2_threshold=3)),
                         (74, dict(occupancy_precision=8, bitdepth=9, width=512), dict()),
                         (75, dict(orientations=tuple(range(9))), dict(grid_smoothing=False))):
        g = pbf(small(rb, seed=seed, **kw), **pk)
        ref = run_stages(codec, g, checker_backend, what=f"pbf {kw} {pk}")
        plain = checker_backend.run_gof(small(rb, seed=seed, **kw), keep=("reconstruct",))
        assert ref.counts(0).total < plain.counts(0).total  # the filter really removed border pixels
        assert (ref.cloud(0, "reconstruct")["boundary_types"] == 1).any()
    # the whole decoder sequence in one call: the attribute re-transfer is skipped (PCCDecoder.cpp:445)
    g = pbf(small(rb, seed=76, transfer_filter=1, orientations=tuple(range(9))))
    ref = checker_backend.run_gof(g, keep=("rgb8",))
    codec.uploadGof(g)
    codec.decodeGof()
    for f in range(g.n_frames):
        assert codec.computeChecksum(f) == ref.md5(f)


def test_multiple_streams_relative_t1(rb, codec, checker_backend):
    """CTC condition T1-from-rec-T0: the second attribute map is a delta on the first (PCCCodec.cpp:1387-1416)"""
    g = rb.synthetic.make_relative_t1(small(rb, seed=18), seed=2)
    run_stages(codec, g, checker_backend, what="relt1")
    g = rb.synthetic.make_relative_t1(small(rb, seed=19, orientations=tuple(range(9)), occupancy_precision=2), seed=4)
    g.params.relative_t1 = 0  # multiple streams with an absolute second map: the planes are used as they are
    run_stages(codec, g, checker_backend, what="streams_abs_t1")


def test_smoothing_tables_overflow_and_regrow(rb, checker_backend):
    """the sparse smoothing grids start far too small (test hook): block pool, block table and luma lists overflow, the
    filters do nothing, the stage is repeated with larger tables — and the result is still the reference's"""
    c = rb.codec.PCCCodecB200(device=0)  # a fresh context: the growth state lives in the context
    assert c._lib.rb200_debug_set_grid_shrink(c._h, 9) == 0
    try:
        run_stages(c, small(rb, seed=31), checker_backend, what="regrow")
        run_stages(c, small(rb, seed=32, occupancy_precision=2, orientations=tuple(range(9))), checker_backend, what="regrow2")
    finally:
        c.close()


def _lossy(rb, precision):
    g = small(rb, seed=16, occupancy_precision=precision)
    rng = np.random.default_rng(5)
    g.occupancy[...] = np.where(g.occupancy != 0, rng.integers(1, 9, g.occupancy.shape),
                                rng.integers(0, 3, g.occupancy.shape)).astype(np.uint8)
    g.params.threshold_lossy_om = 2
    return g


def test_lossy_threshold_and_size_quantization(rb, codec, checker_backend):
    g = _lossy(rb, 1)
    g.params.enable_size_quantization = 1
    g.params.log2_quantizer_x = 2
    g.params.log2_quantizer_y = 3
    g.patches["size2d_x_px"] -= 5
    g.patches["size2d_y_px"] -= 3
    ref = run_stages(codec, g, checker_backend, what="lossy")
    assert ref.counts(0).total > 1000


def test_lossy_threshold_quirk_precision4(rb, codec, checker_backend):
    # PCCCodec.cpp:1597-1600 binarises the sample in place p*p times: threshold >= 1 with p > 1 empties the frame
    ref = run_stages(codec, _lossy(rb, 4), checker_backend, what="lossyquirk")
    assert ref.counts(0).total == 0


def test_eom(rb, codec, checker_backend):
    g = small(rb, eom=True, seed=17, geometry_smoothing=False, color_smoothing=False)  # lossless-style cfg
    run_stages(codec, g, checker_backend, stages=("reconstruct", "rgb8"), what="eom")


def test_eom_with_smoothing(rb, codec, checker_backend):
    g = small(rb, eom=True, seed=18, precedence_reverse=True)
    g.params.flag_geometry_smoothing = g.params.apply_geo_smoothing = 1
    g.params.flag_color_smoothing = g.params.apply_attr_smoothing = 1
    run_stages(codec, g, checker_backend, what="eomsmooth")


def test_raw_points(rb, codec, checker_backend):
    run_stages(codec, small(rb, raw_points=1000, seed=19), checker_backend, what="raw")


def test_raw_points_in_the_auxiliary_video(rb, codec, checker_backend):
    """asps.getAuxiliaryVideoEnabledFlag: the raw patches address context.getVideoRawPointsGeometry() (PCCCodec.cpp:895-897)
    and their colours come from the auxiliary attribute video through 8-bit PCCColor3B values (:1524-1549, :1436-1439)"""
    g = rb.synthetic.make_aux_video(small(rb, raw_points=1000, seed=21), seed=21)
    ref = run_stages(codec, g, checker_backend, what="aux raw")
    c = ref.cloud(0, "reconstruct")
    n = ref.counts(0).raw
    assert n == 1000 and (c["colors16"][-n:] < 256).all() and (c["colors16"][-n:] > 0).any()
    # the whole decoder sequence (Rec-1) and without attributes
    g = rb.synthetic.make_aux_video(small(rb, raw_points=333, seed=22, transfer_filter=1), seed=22)
    run_stages(codec, g, checker_backend, stages=ALL_STAGES, what="aux raw rec-1")
    g = rb.synthetic.make_aux_video(small(rb, raw_points=50, seed=23, color_smoothing=False), seed=23)
    g.params.attribute_count = 0
    ref = checker_backend.run_gof(g, keep=("reconstruct",))
    codec.uploadGof(g)
    codec.generatePointCloud()
    for f in range(g.n_frames):
        assert np.array_equal(codec.getPointCloud(f, fields=("positions",))["positions"], ref.cloud(f, "reconstruct")["positions"])
    # EOM with the auxiliary video: synthetic pixel addresses from (0, 0), no occupancy marks (:852-853, :880), colours of the
    # EOM points from the auxiliary attribute video (:1551-1580); with geometry smoothing on the boundary pass reads the
    # occupancy map at those addresses (:964-971)
    for seed, smooth in ((24, False), (25, True)):
        g = rb.synthetic.make_aux_video(small(rb, eom=True, seed=seed, geometry_smoothing=smooth, color_smoothing=smooth), seed=seed)
        if smooth:
            g.params.flag_geometry_smoothing = g.params.apply_geo_smoothing = 1
            g.params.flag_color_smoothing = g.params.apply_attr_smoothing = 1
        ref = run_stages(codec, g, checker_backend, stages=None if smooth else ("reconstruct", "rgb8"), what=f"aux eom smoothing {smooth}")
        c, n0, n = ref.cloud(0, "reconstruct"), ref.counts(0).regular, ref.counts(0).eom
        assert n > 1000 and (c["colors16"][n0:n0 + n] < 256).all() and (c["point_to_pixel"][n0] == 0).all()
    # EOM and raw points together
    g = rb.synthetic.make_aux_video(small(rb, eom=True, raw_points=400, seed=26, geometry_smoothing=False, color_smoothing=False), seed=26)
    ref = run_stages(codec, g, checker_backend, stages=("reconstruct", "rgb8"), what="aux eom + raw")
    assert ref.counts(0).raw == 400 and ref.counts(0).eom > 0
    # an eomCount_ that is not what the member patches produce cannot address the colours: refused
    g = rb.synthetic.make_aux_video(small(rb, eom=True, seed=27, geometry_smoothing=False, color_smoothing=False), seed=27)
    g.eom_patches[0]["eom_count"] += 1
    codec.uploadGof(g)
    with pytest.raises(rb.codec.RabbitError):
        codec.generatePointCloud()


def test_no_attributes(rb, codec, checker_backend):
    g = small(rb, seed=20, color_smoothing=False)
    g.params.attribute_count = 0
    ref = checker_backend.run_gof(g, keep=("reconstruct", "smooth_geometry"))
    codec.uploadGof(g)
    codec.generatePointCloud()
    codec.smoothPointCloudPostprocess()
    counts = codec.frameCounts()
    for f in range(g.n_frames):
        got = codec.getPointCloud(f, counts, fields=("positions", "boundary_types"))
        want = ref.cloud(f, "smooth_geometry")
        assert np.array_equal(got["positions"], want["positions"])
        assert np.array_equal(got["boundary_types"], want["boundary_types"])


def test_empty_frame_and_empty_gof(rb, codec, checker_backend):
    g = small(rb, seed=21)
    g.occupancy[1] = 0  # frame 1 decodes to zero points
    run_stages(codec, g, checker_backend, what="emptyframe")
    g.occupancy[...] = 0
    run_stages(codec, g, checker_backend, what="emptygof")


def test_decode_gof_matches_reference_md5(rb, codec, checker_backend):
    import hashlib
    g = small(rb, n_frames=3, seed=22)
    ref = checker_backend.run_gof(g, keep=("rgb8",))
    codec.uploadGof(g)
    codec.decodeGof()
    counts = codec.frameCounts()
    for f in range(g.n_frames):
        c = codec.getPointCloud(f, counts, fields=("positions", "colors"))
        md5 = hashlib.md5(c["positions"].tobytes() + c["colors"].tobytes()).hexdigest()  # PCCPointSet.cpp:232-245
        assert md5 == ref.md5(f), f"frame {f}: ordered MD5 differs"


def test_vox10_full_size_frame(rb, codec, checker_backend):
    g = rb.synthetic.generate_gof(n_frames=1, bitdepth=10, width=1280, scale=0.68, seed=23, transfer_filter=0,
                                  min_height_blocks=80)
    assert g.params.height >= 1280
    run_stages(codec, g, checker_backend, what="vox10")


def test_error_paths(rb, codec):
    g = small(rb, seed=24)
    g.patches["u0"][0] = 1000  # outside the canvas -> the reference exits 180 (PCCPatch.cpp:237-245)
    with pytest.raises(rb.codec.RabbitError) as e:
        codec.uploadGof(g)
    assert e.value.status == rb.abi.RB200_ERR_PATCH_OUT_OF_CANVAS
    g = pbf(small(rb, seed=24, map_count=1))
    g.params.single_map_pixel_interleaving = 1  # occupancy synthesis with pixel interleaving: refused, never approximated
    with pytest.raises(rb.codec.RabbitError) as e:
        codec.uploadGof(g)
    assert e.value.status == rb.abi.RB200_ERR_UNSUPPORTED
    g = small(rb, seed=24)
    g.params.point_local_reconstruction = 1  # needs a single map (and its mode tables)
    with pytest.raises(rb.codec.RabbitError) as e:
        codec.uploadGof(g)
    assert e.value.status == rb.abi.RB200_ERR_INVALID


# ---- attribute re-transfer (PCCPointSet3::transferColors16bitBP): the decoder's default Rec-1 flow ----
ALL_STAGES = ("reconstruct", "smooth_geometry", "transfer_colors", "smooth_color", "rgb8")


def test_transfer_colors_default(rb, codec, checker_backend):
    g = small(rb, transfer_filter=1, seed=51)
    ref = run_stages(codec, g, checker_backend, stages=ALL_STAGES, what="transfer")
    assert ref.counts(0).smoothed > 50


def test_transfer_colors_orientations_reverse(rb, codec, checker_backend):
    g = small(rb, transfer_filter=1, seed=52, orientations=tuple(range(9)), occupancy_precision=2, precedence_reverse=True,
              n_frames=3)
    run_stages(codec, g, checker_backend, stages=ALL_STAGES, what="transfer-orient")


def test_transfer_colors_heavy_noise(rb, codec, checker_backend):
    """more coding noise -> more moved points, longer candidate lists (std::sort beyond 16 elements)"""
    g = small(rb, transfer_filter=1, seed=53, noise_fraction=0.45, bitdepth=9, width=512, scale=0.8, n_frames=1)
    g.params.threshold_smoothing = 8.0
    ref = run_stages(codec, g, checker_backend, stages=ALL_STAGES, what="transfer-noise")
    assert ref.counts(0).smoothed > 2000


def test_transfer_colors_rgb444_lossless_attribute(rb, codec, checker_backend):
    g = small(rb, transfer_filter=1, seed=54)
    g.params.attribute_rgb444 = 1
    run_stages(codec, g, checker_backend, stages=ALL_STAGES, what="transfer-rgb444")


def test_vox10_full_size_frame_with_transfer(rb, codec, checker_backend):
    g = rb.synthetic.generate_gof(n_frames=1, bitdepth=10, width=1280, scale=0.68, seed=55, transfer_filter=1,
                                  height_blocks=80)
    ref = run_stages(codec, g, checker_backend, stages=ALL_STAGES, what="vox10-transfer")
    assert ref.counts(0).smoothed > 5000


# ---- BASELINE.json config 4 shapes: vox11 (11-bit, 2560-wide atlas), lossy two-layer and the lossless-style EOM variant ----
def test_vox10_r5_full_size_frame_precision2(rb, codec, checker_backend):
    """the r5 rate point (cfg/rate/ctc-r5.cfg: occupancyPrecision 2, colour cells of 2^3 voxels) at full vox10 size through
    the whole Rec-1 sequence"""
    g = rb.synthetic.generate_gof(n_frames=1, bitdepth=10, width=1280, scale=0.68, seed=0x0AB817 + 5, transfer_filter=1,
                                  occupancy_precision=2, max_depth=249)
    ref = run_stages(codec, g, checker_backend, stages=ALL_STAGES, what="vox10 r5")
    assert ref.counts(0).total > 600000 and ref.counts(0).recolored > 0


def test_vox11_frame_lossy_full_decoder(rb, codec, checker_backend):
    g = rb.synthetic.generate_gof(n_frames=1, bitdepth=11, width=2560, scale=0.55, seed=61, transfer_filter=1)
    assert g.params.geometry_bitdepth_3d == 11
    ref = run_stages(codec, g, checker_backend, stages=ALL_STAGES, what="vox11")
    assert ref.counts(0).total > 1500000


def test_vox11_frame_eom_lossless_style(rb, codec, checker_backend):
    g = rb.synthetic.generate_gof(n_frames=1, bitdepth=11, width=2560, scale=0.45, seed=62, transfer_filter=0, eom=True,
                                  geometry_smoothing=False, color_smoothing=False)
    ref = run_stages(codec, g, checker_backend, stages=("reconstruct", "rgb8"), what="vox11-eom")
    assert ref.counts(0).eom > 0


# ---- wire format and checksum (PCCPointSet3::write / read / computeChecksum) ----
def test_ply_and_md5_match_reference(rb, codec, checker_backend, tmp_path):
    g = small(rb, seed=65, transfer_filter=1)
    ref = checker_backend.run_gof(small(rb, seed=65, transfer_filter=1), keep=("rgb8",), canonical_md5=True)
    codec.uploadGof(g)
    codec.decodeGof()
    for f in range(g.n_frames):
        want = ref.cloud(f, "rgb8")
        assert codec.computeChecksum(f) == ref.md5(f)
        # computeChecksum( true ): the canonical order (PCCPointSet.cpp:258-296) — differs from the ordered one
        assert codec.computeChecksum(f, True) == ref.md5(f, canonical=True) != ref.md5(f)
        mine, theirs = str(tmp_path / f"b200_{f}.ply"), str(tmp_path / f"ref_{f}.ply")
        codec.write(f, mine)
        assert checker_backend.write_ply(want["positions"], want["colors"], theirs) == 0
        assert open(mine, "rb").read() == open(theirs, "rb").read(), "PLY file differs from PCCPointSet3::write"
        back = codec.read(theirs)
        assert np.array_equal(back["positions"], want["positions"]) and np.array_equal(back["colors"], want["colors"])
        rp, rc = checker_backend.read_ply(mine)
        assert np.array_equal(rp, want["positions"]) and np.array_equal(rc, want["colors"])


def _need_reference(chk):
    from oracle import checker
    if not isinstance(chk, checker.Reference):
        pytest.skip("pixel interleaving is checked against the unmodified reference (oracle/_ref) only")


def test_single_map_pixel_interleaving(rb, codec, checker_backend):
    """singleMapPixelInterleaving (generatePoints, PCCCodec.cpp:350-471): checkerboard layers, the other layer
    interpolated from the 4-neighbours, fill points; their colours by transferColorWeight (5-NN in nanoflann order)"""
    _need_reference(checker_backend)
    g = rb.synthetic.make_pixel_interleaved(small(rb, seed=21))
    ref = run_stages(codec, g, checker_backend, what="ilv")
    layers = ref.cloud(0, "reconstruct")["point_to_pixel"][:, 2]
    assert (layers == 100).sum() > 1000 and (layers == 0).sum() > 1000 and (layers == 1).sum() > 1000
    # every orientation, p = 2, thinner surface, duplicates kept
    g = rb.synthetic.make_pixel_interleaved(small(rb, seed=22, orientations=tuple(range(9)), occupancy_precision=2),
                                            surface_thickness=2)
    g.params.remove_duplicate_points = 0
    run_stages(codec, g, checker_backend, what="ilv_orient")


def test_pixel_interleaving_noisy_depth_and_full_decoder(rb, codec, checker_backend):
    """random depth codes drive both clamps of the interpolation and the size_t wrap of d1 - depth (:385-389); then the
    whole Rec-1 sequence including the attribute re-transfer runs on the interleaved cloud"""
    _need_reference(checker_backend)
    g = rb.synthetic.make_pixel_interleaved(small(rb, seed=23, transfer_filter=1), surface_thickness=6)
    rng = np.random.default_rng(77)
    noisy = rng.random(g.geometry.shape) < 0.02
    g.geometry[noisy] = rng.integers(0, 256, int(noisy.sum())).astype(np.uint16)
    g.params.geometry_bitdepth_3d = 9  # the noise pushes coordinates past 255; 2^bitdepth must still bound them
    run_stages(codec, g, checker_backend,
               stages=("reconstruct", "smooth_geometry", "transfer_colors", "smooth_color", "rgb8"), what="ilv_full")


def test_pixel_interleaving_argument_checks(rb, codec):
    g = rb.synthetic.make_pixel_interleaved(small(rb, seed=24))
    g.params.surface_thickness = 0
    with pytest.raises(Exception):
        codec.uploadGof(g)


def test_point_local_reconstruction(rb, codec, checker_backend):
    """pointLocalReconstruction (generatePoints :472-496, getDeltaNeighbors :238-264): per-block modes (interpolate,
    filling, minD1, neighbour window), second point and fills on layers 100 / 101, coloured by transferColorWeight"""
    _need_reference(checker_backend)
    g = rb.synthetic.make_plr(small(rb, seed=41, map_count=1), seed=1)
    ref = run_stages(codec, g, checker_backend, what="plr")
    layers = ref.cloud(0, "reconstruct")["point_to_pixel"][:, 2]
    assert (layers == 100).sum() > 1000 and (layers == 101).sum() > 1000
    g = rb.synthetic.make_plr(small(rb, seed=42, map_count=1, orientations=tuple(range(9)), occupancy_precision=2,
                                    transfer_filter=1), seed=2)
    run_stages(codec, g, checker_backend,
               stages=("reconstruct", "smooth_geometry", "transfer_colors", "smooth_color", "rgb8"), what="plr_full")


def test_point_local_reconstruction_needs_its_tables(rb, codec):
    g = rb.synthetic.make_plr(small(rb, seed=43, map_count=1), seed=3)
    g.plr = None
    codec.uploadGof(g)
    with pytest.raises(Exception):
        codec.generatePointCloud()


def test_interleaving_and_plr_with_lossy_occupancy_reverse_order(rb, codec, checker_backend):
    """both variable-count modes on top of the other switches of the reprojection: lossy occupancy threshold, patch
    size quantisation (the neighbour tests read the quantised map), reversed patch precedence, precision 1"""
    _need_reference(checker_backend)
    for tag in ("ilv", "plr"):
        g = _lossy(rb, 1)
        g.params.enable_size_quantization = 1
        g.params.log2_quantizer_x = 2
        g.params.log2_quantizer_y = 3
        g.patches["size2d_x_px"] -= 5
        g.patches["size2d_y_px"] -= 3
        g.params.patch_precedence_reverse = 1
        if tag == "ilv":
            rb.synthetic.make_pixel_interleaved(g, surface_thickness=3)
        else:
            h, w = g.params.height, g.params.width
            g.geometry = np.ascontiguousarray(g.geometry.reshape(g.n_frames, 2, h, w)[:, :1])
            g.attribute = np.ascontiguousarray(g.attribute.reshape(g.n_frames, 2, 3, h, w)[:, :1])
            g.params.map_count_minus1 = 0
            rb.synthetic.make_plr(g, seed=8)
        run_stages(codec, g, checker_backend, what="lossy_" + tag)


def test_interleaving_and_plr_with_raw_patches(rb, codec, checker_backend):
    """raw (missed-point) patches next to the variable-count modes: with pixel interleaving colorPointCloud applies the
    checkerboard rule to the raw points as well (:1367-1374: those on odd pixels are coloured by transferColorWeight),
    with point local reconstruction they sit on layer 0 and read the attribute frame"""
    _need_reference(checker_backend)
    g = rb.synthetic.make_pixel_interleaved(small(rb, seed=61, raw_points=700))
    ref = run_stages(codec, g, checker_backend, what="ilv_raw")
    assert ref.counts(0).raw == 700
    g = rb.synthetic.make_plr(small(rb, seed=62, raw_points=500, map_count=1), seed=9)
    ref = run_stages(codec, g, checker_backend, what="plr_raw")
    assert ref.counts(0).raw == 500
