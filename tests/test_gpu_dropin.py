"""The drop-in boundary: the reference's own classes and decoder sequence (oracle/ref_harness.cpp, unmodified reference
objects) with the hot member functions replaced by rabbit-transcoding_b200/host/PCCCodecB200.cpp, which calls the CUDA
library.  Everything the callers see — PCCPointSet3 contents after every stage, partition, pointToPixel, block-to-patch,
point counts, ordered MD5, PCCMetrics results — must equal the unmodified reference's."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import FIELDS, assert_cloud_equal

pytestmark = pytest.mark.gpu
STAGES = ("reconstruct", "smooth_geometry", "transfer_colors", "smooth_color", "rgb8")


@pytest.fixture(scope="module")
def backends():
    from oracle import checker
    if not (checker.have_reference() and checker.have_dropin()):
        pytest.skip("oracle/_ref/librabbit_ref.so / librabbit_dropin.so not built (make -C oracle ref dropin)")
    return checker.Reference(), checker.DropIn()


def compare(rb, backends, what, make=None, **kw):
    ref_b, drop_b = backends
    args = dict(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=71, transfer_filter=1)
    args.update(kw)
    g = make(rb.synthetic.generate_gof(**args)) if make else rb.synthetic.generate_gof(**args)
    want = ref_b.run_gof(g, keep=STAGES)
    drop_b.stats(reset=True)
    g2 = make(rb.synthetic.generate_gof(**args)) if make else rb.synthetic.generate_gof(**args)
    got = drop_b.run_gof(g2, keep=STAGES)  # fresh arrays: the original binarises in place
    st = drop_b.stats()
    # the results below were computed by the CUDA library: its kernels ran and the clouds came back over PCIe
    assert st.kernel_launches >= 5 * g.n_frames and st.d2h_bytes > 0 and st.h2d_bytes > 0, (st.kernel_launches, st.d2h_bytes)
    for f in range(g.n_frames):
        cw, cg = want.counts(f), got.counts(f)
        assert (cw.total, cw.regular, cw.eom, cw.raw, cw.smoothed, cw.recolored) == \
            (cg.total, cg.regular, cg.eom, cg.raw, cg.smoothed, cg.recolored), f"{what} frame {f} counts"
        for st in STAGES:
            try:
                w = want.cloud(f, st)
            except KeyError:
                continue
            assert_cloud_equal(got.cloud(f, st), w, f"{what} frame {f} stage {st}", FIELDS)
        assert got.md5(f) == want.md5(f)
        assert np.array_equal(got.block_to_patch(f, g.params), want.block_to_patch(f, g.params))
        assert np.array_equal(got.occupancy(f, g.params) != 0, want.occupancy(f, g.params) != 0)
    return g, want


def test_dropin_default_decoder_sequence(rb, backends):
    g, want = compare(rb, backends, "default")
    assert want.counts(0).smoothed > 0


def test_dropin_orientations_reverse_precision2(rb, backends):
    compare(rb, backends, "orient", orientations=tuple(range(9)), occupancy_precision=2, precedence_reverse=True, seed=72)


def test_dropin_eom_and_raw(rb, backends):
    compare(rb, backends, "eom", eom=True, geometry_smoothing=False, color_smoothing=False, transfer_filter=0, seed=73)
    compare(rb, backends, "raw", raw_points=700, transfer_filter=0, seed=74)


def test_dropin_raw_points_in_the_auxiliary_video(rb, backends):
    g, want = compare(rb, backends, "aux raw", make=lambda g: rb.synthetic.make_aux_video(g, seed=79), raw_points=600, seed=79)
    assert want.counts(0).raw == 600
    g, want = compare(rb, backends, "aux eom", make=lambda g: rb.synthetic.make_aux_video(g, seed=80), eom=True, geometry_smoothing=False,
                      color_smoothing=False, transfer_filter=0, seed=80)
    assert want.counts(0).eom > 0


def test_dropin_occupancy_synthesis(rb, backends):
    """Rec-2 through the shim: the decoder skips generateOccupancyMap and the re-transfer (PCCDecoder.cpp:362, :445), the
    patch border filter and the points' boundary types run on the GPU"""
    from test_gpu_parity import pbf
    g, want = compare(rb, backends, "pbf", make=pbf, orientations=tuple(range(9)), seed=77)
    assert (want.cloud(0, "reconstruct")["boundary_types"] == 1).any() and want.counts(0).smoothed > 0


def test_dropin_non_grid_smoothing(rb, backends):
    """smoothPointCloudPostprocess with gridSmoothing_ == 0 through the shim (the encoder's reconstruction, PCCCodec.cpp:141):
    the reference's 8-bit debug colouring of the moved points (:1143) is not reproduced — the callers overwrite the 8-bit
    colours (convertYUV16ToRGB8), as the rgb8 stage compared here shows"""
    def non_grid(g):
        g.params.grid_smoothing = 0
        g.params.neighbor_count_smoothing = 64
        g.params.radius2_smoothing = 64.0
        g.params.radius2_boundary_detection = 64.0
        return g
    ref_b, drop_b = backends
    args = dict(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=78, transfer_filter=1)
    want = ref_b.run_gof(non_grid(rb.synthetic.generate_gof(**args)), keep=STAGES)
    drop_b.stats(reset=True)
    got = drop_b.run_gof(non_grid(rb.synthetic.generate_gof(**args)), keep=STAGES)
    assert drop_b.stats().kernel_launches > 10
    moved = 0
    for f in range(2):
        for st in STAGES:
            fields = tuple(k for k in FIELDS if k != "colors" or st == "rgb8")
            assert_cloud_equal(got.cloud(f, st), want.cloud(f, st), f"non-grid frame {f} stage {st}", fields)
        assert got.md5(f) == want.md5(f)
        moved += int((want.cloud(f, "reconstruct")["positions"] != want.cloud(f, "smooth_geometry")["positions"]).any(axis=1).sum())
    assert moved > 100


def test_dropin_links_no_reference_body():
    """the shim has no way back into the reference's own bodies: no rb200_orig_* symbol, and every replaced member is
    defined exactly once (the shim's strong definition)"""
    from oracle import checker
    if not checker.have_dropin():
        pytest.skip("librabbit_dropin.so not built")
    out = subprocess.check_output(["nm", "-D", "--defined-only", checker.DROPIN_LIB], text=True)
    assert "rb200_orig" not in out
    for member in ("PCCCodec18generatePointCloud", "PCCCodec15colorPointCloud", "PCCCodec27smoothPointCloudPostprocess",
                   "PCCCodec14colorSmoothing", "PCCCodec20generateOccupancyMap", "PCCPointSet321transferColors16bitBP",
                   "PCCMetrics7computeERKNS_16PCCGroupOfFrames"):
        assert sum(1 for line in out.splitlines() if member in line and ".cold" not in line) == 1, member


def test_dropin_multiple_streams_relative_t1(rb, backends):
    """multiple map streams with a delta-coded second attribute map run on the GPU behind the member functions"""
    compare(rb, backends, "relt1", make=lambda g: rb.synthetic.make_relative_t1(g, seed=5), n_frames=2, seed=75)


def test_dropin_unsupported_mode_exits_like_the_reference(rb, backends):
    """a mode the CUDA path does not implement ends with a message and a non-zero exit status (the reference's error
    convention, PCCMetrics.cpp:342-346) — never with a result computed somewhere else"""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import rabbit_transcoding_b200 as rb\n"
        "from oracle import checker\n"
        "g = rb.synthetic.generate_gof(n_frames=1, bitdepth=8, width=256, scale=0.9, seed=79, transfer_filter=1)\n"
        "g.params.occupancy_resolution = 8\n"
        "checker.DropIn().run_gof(g, keep=('rgb8',), quiet=False)\n"
        "print('SURVIVED')\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "SURVIVED" not in r.stdout
    assert "not implemented on the CUDA path" in r.stderr, r.stderr[-2000:]


def test_dropin_pixel_interleaving_and_plr(rb, backends):
    """singleMapPixelInterleaving and pointLocalReconstruction run on the GPU behind the reference's member functions"""
    ref_b, drop_b = backends
    stages = ("reconstruct", "smooth_geometry", "transfer_colors", "smooth_color", "rgb8")
    for tag, make in (("ilv", lambda: rb.synthetic.make_pixel_interleaved(
                           rb.synthetic.generate_gof(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=77, transfer_filter=1))),
                      ("plr", lambda: rb.synthetic.make_plr(
                           rb.synthetic.generate_gof(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=78, transfer_filter=1,
                                                     map_count=1), seed=6))):
        want = ref_b.run_gof(make(), keep=stages)
        drop_b.stats(reset=True)
        got = drop_b.run_gof(make(), keep=stages)
        assert drop_b.stats().kernel_launches > 10
        for f in range(2):
            assert got.md5(f) == want.md5(f), tag
            for st in stages:
                assert_cloud_equal(got.cloud(f, st), want.cloud(f, st), f"dropin {tag} frame {f} {st}")


def test_dropin_metrics(rb, backends):
    from oracle import checker
    ref_b, drop_b = backends
    g = rb.synthetic.generate_gof(n_frames=1, bitdepth=8, width=256, scale=0.9, seed=76, transfer_filter=0)
    rec = ref_b.run_gof(g, keep=("rgb8",)).cloud(0, "rgb8")
    mp = checker.default_metrics_params(resolution=255.0)
    want, _ = ref_b.metrics(mp, g.sources[0], rec, g.sources[0])
    drop_b.stats(reset=True)
    got, _ = drop_b.metrics(mp, g.sources[0], rec, g.sources[0])
    assert drop_b.stats().kernel_launches > 5
    for tag in ("q1", "q2", "qf"):
        a, b = getattr(got, tag), getattr(want, tag)
        assert a.c2c_mse == b.c2c_mse
        assert abs(a.c2c_psnr - b.c2c_psnr) <= 1e-6 and abs(a.c2p_psnr - b.c2p_psnr) <= 1e-6
        for k in range(3):
            assert abs(a.color_psnr[k] - b.color_psnr[k]) <= 1e-6
    assert (got.source_after_dedup, got.rec_after_dedup) == (want.source_after_dedup, want.rec_after_dedup)
