"""GPU parity of the metrics path (duplicate removal, D1 / D2 / colour PSNR) through the C ABI against the
unmodified reference's PCCMetrics (oracle/_ref) on identical clouds.

Tolerances (BASELINE.json north_star): PSNR within 1e-6 dB; MSE floats within 1 ulp-ish relative 1e-6; the exact
integer quantities (duplicate-removed clouds, point counts, the d^2 sums behind D1) bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PSNR_TOL_DB = 1e-6


def _decoded_pair(rb, codec, seed=31, **kw):
    args = dict(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=seed, transfer_filter=0)
    args.update(kw)
    g = rb.synthetic.generate_gof(**args)
    codec.uploadGof(g)
    codec.decodeGof()
    counts = codec.frameCounts()
    recs = [codec.getPointCloud(f, counts, fields=("positions", "colors")) for f in range(g.n_frames)]
    return g, recs


def _close(a, b, what):
    if np.isinf(a) or np.isinf(b):
        assert a == b, f"{what}: {a} vs {b}"
    else:
        assert abs(a - b) <= PSNR_TOL_DB, f"{what}: {a} vs {b} (|d| = {abs(a - b)})"


def _compare(got, want, what, c2p=True):
    for tag in ("q1", "q2", "qf"):
        g, w = getattr(got, tag), getattr(want, tag)
        _close(g.c2c_psnr, w.c2c_psnr, f"{what} {tag} c2c_psnr")
        assert g.c2c_mse == w.c2c_mse, f"{what} {tag} c2c_mse {g.c2c_mse} vs {w.c2c_mse}"  # exact: integer sums
        if c2p:
            _close(g.c2p_psnr, w.c2p_psnr, f"{what} {tag} c2p_psnr")
            assert abs(g.c2p_mse - w.c2p_mse) <= 1e-6 * max(1e-30, abs(w.c2p_mse)), f"{what} {tag} c2p_mse"
        for k in range(3):
            _close(g.color_psnr[k], w.color_psnr[k], f"{what} {tag} color_psnr[{k}]")
            assert abs(g.color_mse[k] - w.color_mse[k]) <= 1e-6 * max(1e-30, abs(w.color_mse[k]))
    assert (got.source_points, got.source_after_dedup, got.rec_points, got.rec_after_dedup) == \
        (want.source_points, want.source_after_dedup, want.rec_points, want.rec_after_dedup), what


def test_remove_duplicates_bit_exact(rb, codec, checker_backend):
    g, recs = _decoded_pair(rb, codec)
    m = rb.metrics.PCCMetricsB200(codec)
    for drop in (1, 2):
        for f, rec in enumerate(recs):
            wp, wc = checker_backend.remove_duplicates(rec["positions"], rec["colors"], drop)
            got = m.removeDuplicate(rec, drop)
            assert got["positions"].shape == wp.shape, (drop, f, got["positions"].shape, wp.shape)
            assert len(wp) < len(rec["positions"]), "the decoded cloud should contain duplicates"
            assert np.array_equal(got["positions"], wp), f"drop {drop} frame {f}: positions / order differ"
            assert np.array_equal(got["colors"], wc), f"drop {drop} frame {f}: merged colours differ"


def test_remove_duplicates_edge_cases(rb, codec, checker_backend):
    m = rb.metrics.PCCMetricsB200(codec)
    rng = np.random.default_rng(3)
    # heavy duplication, negative coordinates, one long z column
    pos = rng.integers(-3, 4, size=(5000, 3)).astype(np.int16)
    pos[:1500, 0] = 2
    pos[:1500, 1] = -1
    pos[:1500, 2] = rng.integers(-700, 700, 1500)
    col = rng.integers(0, 256, size=(5000, 3)).astype(np.uint8)
    for drop in (1, 2):
        wp, wc = checker_backend.remove_duplicates(pos, col, drop)
        got = m.removeDuplicate(dict(positions=pos, colors=col), drop)
        assert np.array_equal(got["positions"], wp) and np.array_equal(got["colors"], wc)
    one = dict(positions=np.array([[5, 6, 7]], np.int16), colors=np.array([[1, 2, 3]], np.uint8))
    got = m.removeDuplicate(one)
    assert got["positions"].tolist() == [[5, 6, 7]] and got["colors"].tolist() == [[1, 2, 3]]


def test_d1_d2_colour_vs_reference(rb, codec, checker_backend):
    from oracle import checker
    g, recs = _decoded_pair(rb, codec, seed=32)
    mp = checker.default_metrics_params(resolution=255.0)
    m = rb.metrics.PCCMetricsB200(codec)
    m.setParameters(mp)
    res = m.compute(g.sources, recs, g.sources)
    for f in range(g.n_frames):
        want, _ = checker_backend.metrics(mp, g.sources[f], recs[f], g.sources[f])
        assert res[f].tie_overflow == 0
        _compare(res[f], want, f"frame {f}")
        assert np.isfinite(want.q1.c2p_psnr) and want.q1.c2c_mse > 0


def test_d1_colour_without_normals_and_resident_frames(rb, codec, checker_backend):
    """transcode.sh's configuration: no --normalDataPath, so D1 + colour only (PCCMetricsParameters.cpp:110);
    the reconstruction is taken from the GOF resident in the context (no host round trip)."""
    from oracle import checker
    g, recs = _decoded_pair(rb, codec, seed=33, n_frames=3)
    mp = checker.default_metrics_params(resolution=255.0, c2p=False)
    m = rb.metrics.PCCMetricsB200(codec)
    m.setParameters(mp)
    res = m.compute(g.sources, [None] * g.n_frames, None)
    for f in range(g.n_frames):
        want, _ = checker_backend.metrics(mp, g.sources[f], recs[f], None)
        _compare(res[f], want, f"resident frame {f}", c2p=False)


def test_hausdorff_dropdup_variants(rb, codec, checker_backend):
    from oracle import checker
    g, recs = _decoded_pair(rb, codec, seed=34, n_frames=1)
    for drop, nproc in ((1, 1), (2, 2), (2, 4), (2, 3)):
        mp = checker.default_metrics_params(resolution=255.0)
        mp.compute_hausdorff = 1
        mp.drop_duplicates = drop
        mp.neighbors_proc = nproc
        m = rb.metrics.PCCMetricsB200(codec)
        m.setParameters(mp)
        got = m.compute(g.sources[:1], recs[:1], g.sources[:1])[0]
        want, _ = checker_backend.metrics(mp, g.sources[0], recs[0], g.sources[0])
        _compare(got, want, f"drop {drop} nproc {nproc}")
        for tag in ("q1", "q2", "qf"):
            a, b = getattr(got, tag), getattr(want, tag)
            assert a.c2c_hausdorff == b.c2c_hausdorff
            _close(a.c2c_hausdorff_psnr, b.c2c_hausdorff_psnr, "c2c hausdorff psnr")
            _close(a.c2p_hausdorff_psnr, b.c2p_hausdorff_psnr, "c2p hausdorff psnr")


def test_neighbors_proc_0_first_nearest_in_traversal_order(rb, codec, checker_backend):
    """neighborsProc 0 (PCCMetrics.cpp:126, :177): the colour of result.indices( 0 ) — among equidistant nearest points the
    one nanoflann's traversal meets first — through the emulated kd forest over the de-duplicated clouds"""
    from oracle import checker
    g, recs = _decoded_pair(rb, codec, seed=36, n_frames=2)
    mp = checker.default_metrics_params(resolution=255.0)
    mp.neighbors_proc = 0
    m = rb.metrics.PCCMetricsB200(codec)
    m.setParameters(mp)
    res = m.compute(g.sources, recs, g.sources)
    for f in range(g.n_frames):
        want, _ = checker_backend.metrics(mp, g.sources[f], recs[f], g.sources[f])
        _compare(res[f], want, f"nproc 0 frame {f}")
    # a lattice with many equidistant neighbours and random colours: the choice among the ties decides the result
    rng = np.random.default_rng(12)
    a = np.unique(rng.integers(0, 24, size=(5000, 3)).astype(np.int16) * 2, axis=0)          # even coordinates
    b = np.unique(rng.integers(0, 24, size=(5000, 3)).astype(np.int16) * 2 + 1, axis=0)      # odd: up to 8 ties at distance 3
    src = dict(positions=a, colors=rng.integers(0, 256, size=(len(a), 3)).astype(np.uint8))
    rec = dict(positions=b, colors=rng.integers(0, 256, size=(len(b), 3)).astype(np.uint8))
    mp = checker.default_metrics_params(resolution=255.0, c2p=False)
    mp.neighbors_proc = 0
    m.setParameters(mp)
    got = m.compute([src], [rec], None)[0]
    want, _ = checker_backend.metrics(mp, src, rec, None)
    _compare(got, want, "nproc 0 lattice", c2p=False)
    mp.neighbors_proc = 1
    m.setParameters(mp)
    other = m.compute([src], [rec], None)[0]
    assert other.q1.color_mse[0] != got.q1.color_mse[0]  # the tie rule matters on this input


def test_far_queries_and_sparse_clouds(rb, codec, checker_backend):
    """clouds that are far apart / sparse exercise the warp-per-query ring search"""
    from oracle import checker
    rng = np.random.default_rng(9)
    a = np.unique(rng.integers(0, 200, size=(3000, 3)).astype(np.int16), axis=0)
    b = np.unique((rng.integers(0, 60, size=(2000, 3)) + np.array([120, 90, 10])).astype(np.int16), axis=0)
    ca = rng.integers(0, 256, size=(len(a), 3)).astype(np.uint8)
    cb = rng.integers(0, 256, size=(len(b), 3)).astype(np.uint8)
    na = rng.normal(size=(len(a), 3)).astype(np.float32)
    src = dict(positions=a, colors=ca, normals=na)
    rec = dict(positions=b, colors=cb)
    mp = checker.default_metrics_params(resolution=1023.0)
    mp.compute_hausdorff = 1
    m = rb.metrics.PCCMetricsB200(codec)
    m.setParameters(mp)
    got = m.compute([src], [rec], [src])[0]
    want, _ = checker_backend.metrics(mp, src, rec, src)
    _compare(got, want, "far")
    assert want.q1.c2c_hausdorff > 100


def test_cached_device_sources_keep_their_index_across_calls(rb, codec, checker_backend):
    """rb200_metrics_cache_sources: the transcode loop measures the same source frames against every rate point.  The
    second and third call reuse the sources' part of the batch (different reconstructions, host views and the
    resident GOF); a reconstruction outside the kept tables forces the rebuild path; all equal the reference."""
    import torch
    from oracle import checker
    g, recs = _decoded_pair(rb, codec, seed=36, n_frames=3)
    dev = [{k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in s.items() if k in ("positions", "colors", "normals")}
           for s in g.sources]
    mp = checker.default_metrics_params(resolution=255.0)
    m = rb.metrics.PCCMetricsB200(codec)
    m.setParameters(mp)
    m.cacheSources(True)
    want = [checker_backend.metrics(mp, g.sources[f], recs[f], g.sources[f])[0] for f in range(g.n_frames)]
    for rnd in range(2):  # first call builds, second reuses (resident reconstruction)
        res = m.compute(dev, [None] * g.n_frames, dev)
        for f in range(g.n_frames):
            _compare(res[f], want[f], f"round {rnd} frame {f}")
    # other reconstructions against the kept sources: frames rotated, as host views
    rot = [recs[(f + 1) % g.n_frames] for f in range(g.n_frames)]
    res = m.compute(dev, rot, dev)
    for f in range(g.n_frames):
        w, _ = checker_backend.metrics(mp, g.sources[f], rot[f], g.sources[f])
        _compare(res[f], w, f"rotated frame {f}")
    # a reconstruction far outside the kept column tables: rebuilt from scratch, same answer
    far = [dict(positions=(r["positions"].astype(np.int32) + np.array([300, 0, 40])).astype(np.int16), colors=r["colors"]) for r in recs]
    res = m.compute(dev, far, dev)
    for f in range(g.n_frames):
        w, _ = checker_backend.metrics(mp, g.sources[f], far[f], g.sources[f])
        _compare(res[f], w, f"far frame {f}")
    # and the settings are part of the key: drop_duplicates 1 after 2
    mp1 = checker.default_metrics_params(resolution=255.0)
    mp1.drop_duplicates = 1
    m.setParameters(mp1)
    res = m.compute(dev, recs, dev)
    for f in range(g.n_frames):
        w, _ = checker_backend.metrics(mp1, g.sources[f], recs[f], g.sources[f])
        _compare(res[f], w, f"drop 1 frame {f}")
    res = m.compute(dev, recs, None)  # without normals: D1 + colour only, cached index of the drop-1 call is not valid for c2p off
    mpn = checker.default_metrics_params(resolution=255.0, c2p=False)
    mpn.drop_duplicates = 1
    for f in range(g.n_frames):
        w, _ = checker_backend.metrics(mpn, g.sources[f], recs[f], None)
        _compare(res[f], w, f"no normals frame {f}", c2p=False)
    m.cacheSources(False)


def test_vox10_frame_metrics(rb, codec, checker_backend):
    from oracle import checker
    g = rb.synthetic.generate_gof(n_frames=1, bitdepth=10, width=1280, scale=0.68, seed=35, transfer_filter=0,
                                  height_blocks=80)
    codec.uploadGof(g)
    codec.decodeGof()
    rec = codec.getPointCloud(0, fields=("positions", "colors"))
    mp = checker.default_metrics_params(resolution=1023.0)
    m = rb.metrics.PCCMetricsB200(codec)
    m.setParameters(mp)
    got = m.compute(g.sources, [None], g.sources)[0]
    want, ms = checker_backend.metrics(mp, g.sources[0], rec, g.sources[0])
    _compare(got, want, "vox10")
    assert 40 < want.qf.c2c_psnr < 90


def test_metrics_errors(rb, codec):
    m = rb.metrics.PCCMetricsB200(codec)
    with pytest.raises(rb.codec.RabbitError):
        m.compute([dict(positions=np.zeros((4, 3), np.int16))], [])
    bad = dict(positions=np.zeros((0, 3), np.int16), colors=np.zeros((0, 3), np.uint8))
    with pytest.raises(rb.codec.RabbitError):
        m.compute([bad], [bad])
