import numpy as np

FIELDS = ("positions", "colors16", "colors", "boundary_types", "partition", "point_to_pixel")


def assert_cloud_equal(got, want, what, fields=FIELDS):
    """bit-exact comparison with a useful first-mismatch report"""
    for k in fields:
        g, w = got[k], want[k]
        assert g.shape == w.shape, f"{what}: {k} shape {g.shape} != {w.shape}"
        if not np.array_equal(g, w):
            bad = np.nonzero((g != w).reshape(len(g), -1).any(axis=1))[0]
            i = int(bad[0])
            ctx = {f: (got[f][i].tolist(), want[f][i].tolist()) for f in fields if len(got[f]) > i}
            raise AssertionError(f"{what}: {k} differs at {len(bad)}/{len(g)} points, first at {i}: (got, want) = {ctx}")


def run_stages(codec, gof, chk, stages=("reconstruct", "smooth_geometry", "smooth_color", "rgb8"), what=""):
    """run the CUDA path stage by stage and compare every stage with the checker's snapshot"""
    stages = stages or ("reconstruct", "smooth_geometry", "smooth_color", "rgb8")
    ref = chk.run_gof(gof, keep=stages)
    codec.uploadGof(gof)
    p = gof.params
    for st in stages:
        if st == "reconstruct":
            codec.generatePointCloud()
        elif st == "smooth_geometry":
            if p.apply_geo_smoothing and p.flag_geometry_smoothing and (p.grid_smoothing or p.neighbor_count_smoothing > 0):
                codec.smoothPointCloudPostprocess()
        elif st == "transfer_colors":
            if p.apply_geo_smoothing and p.flag_geometry_smoothing and p.attr_transfer_filter_type == 1:
                codec.transferColors16bitBP()
        elif st == "smooth_color":
            if p.apply_attr_smoothing and p.flag_color_smoothing:
                codec.colorSmoothing()
        elif st == "rgb8":
            codec.convertYUV16ToRGB8()
        counts = codec.frameCounts()
        for f in range(gof.n_frames):
            want = ref.cloud(f, st)
            got = codec.getPointCloud(f, counts)
            if st != "rgb8":
                got["colors"] = want["colors"]  # colours (u8) are only defined after the RGB conversion
            assert_cloud_equal(got, want, f"{what} frame {f} stage {st}")
            if st == "reconstruct":
                rc = ref.counts(f)
                assert (counts[f].total, counts[f].regular, counts[f].raw) == (rc.total, rc.regular, rc.raw)
                assert np.array_equal(codec.getBlockToPatch(f), ref.block_to_patch(f, p)), f"{what} blockToPatch"
                if not p.pbf_enable:  # (with occupancy synthesis the reference keeps no atlas-space map, PCCDecoder.cpp:362)
                    assert np.array_equal(codec.getOccupancyMap(f) != 0, ref.occupancy(f, p) != 0), f"{what} occupancy"
    counts = codec.frameCounts()
    for f in range(gof.n_frames):
        rc = ref.counts(f)
        assert counts[f].smoothed == rc.smoothed, f"{what} frame {f}: smoothed {counts[f].smoothed} != {rc.smoothed}"
        assert counts[f].recolored == rc.recolored, f"{what} frame {f}: recolored {counts[f].recolored} != {rc.recolored}"
    return ref
