"""The kd-tree emulation must return the SAME neighbours in the SAME order as the reference's PCCKdTree (nanoflann
0x123, leaf size 10) — ties on the integer lattice are broken by the tree's traversal order, which decides the
colours transferColors16bitBP produces.  Checked against the unmodified reference (oracle/_ref, ref_knn)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gpu_knn(rb, codec, cloud, queries, k):
    cloud = np.ascontiguousarray(cloud, np.int16)
    queries = np.ascontiguousarray(queries, np.int16)
    idx = np.zeros((len(queries), k), np.int64)
    dist = np.zeros((len(queries), k), np.float64)
    st = codec._lib.rb200_kdtree_search(codec._h, rb.abi.ptr(cloud), len(cloud), rb.abi.ptr(queries), len(queries), k,
                                        rb.abi.ptr(idx), rb.abi.ptr(dist))
    assert st == 0, codec._lib.rb200_error_string(codec._h).decode()
    return idx, dist


def check(rb, codec, chk, cloud, queries, ks=(1, 5, 8)):
    for k in ks:
        wi, wd = chk.knn(cloud, queries, k)
        gi, gd = gpu_knn(rb, codec, cloud, queries, k)
        assert np.array_equal(gd, wd), f"k={k}: distances differ"
        bad = np.nonzero((gi != wi).any(axis=1))[0]
        assert len(bad) == 0, f"k={k}: neighbour order differs for {len(bad)}/{len(queries)} queries, first {bad[0]}: " \
                              f"got {gi[bad[0]].tolist()} want {wi[bad[0]].tolist()} dist {wd[bad[0]].tolist()}"


def test_random_lattice_with_duplicates(rb, codec, checker_backend):
    rng = np.random.default_rng(1)
    cloud = rng.integers(0, 40, size=(20000, 3)).astype(np.int16)  # dense: many duplicates and ties
    q = rng.integers(-2, 43, size=(4000, 3)).astype(np.int16)
    check(rb, codec, checker_backend, cloud, q)


def test_small_and_degenerate_clouds(rb, codec, checker_backend):
    rng = np.random.default_rng(2)
    for n in (8, 10, 11, 23, 129, 300):
        cloud = rng.integers(0, 6, size=(n, 3)).astype(np.int16)
        q = rng.integers(0, 6, size=(200, 3)).astype(np.int16)
        check(rb, codec, checker_backend, cloud, q)
    # all points identical / on a line / on a plane
    same = np.tile(np.array([[3, 4, 5]], np.int16), (500, 1))
    check(rb, codec, checker_backend, same, rng.integers(0, 8, size=(100, 3)).astype(np.int16))
    line = np.zeros((3000, 3), np.int16)
    line[:, 2] = rng.integers(0, 900, 3000)
    check(rb, codec, checker_backend, line, rng.integers(0, 900, size=(300, 3)).astype(np.int16))
    plane = rng.integers(0, 300, size=(30000, 3)).astype(np.int16)
    plane[:, 1] = 17
    check(rb, codec, checker_backend, plane, plane[rng.integers(0, len(plane), 2000)] + rng.integers(-1, 2, size=(2000, 3)).astype(np.int16))


def test_negative_coordinates(rb, codec, checker_backend):
    rng = np.random.default_rng(4)
    cloud = rng.integers(-50, 50, size=(15000, 3)).astype(np.int16)
    q = rng.integers(-55, 55, size=(2000, 3)).astype(np.int16)
    check(rb, codec, checker_backend, cloud, q, ks=(1, 8))


def test_decoded_cloud_surface(rb, codec, checker_backend):
    """a decoded V-PCC frame: a surface with duplicate points, queried at the points themselves shifted by <= 2"""
    g = rb.synthetic.generate_gof(n_frames=1, bitdepth=9, width=512, scale=0.8, seed=41, transfer_filter=0)
    codec.uploadGof(g)
    codec.generatePointCloud()
    cloud = codec.getPointCloud(0, fields=("positions",))["positions"]
    assert len(cloud) > 100000
    rng = np.random.default_rng(5)
    sel = rng.integers(0, len(cloud), 20000)
    q = (cloud[sel].astype(np.int32) + rng.integers(-2, 3, size=(len(sel), 3))).astype(np.int16)
    check(rb, codec, checker_backend, cloud, q, ks=(1, 8))


def gpu_radius(rb, codec, cloud, queries, radius2, max_results, sorted_):
    cloud = np.ascontiguousarray(cloud, np.int16)
    queries = np.ascontiguousarray(queries, np.int16)
    idx = np.zeros((len(queries), max_results), np.int64)
    dist = np.zeros((len(queries), max_results), np.float64)
    cnt = np.zeros(len(queries), np.int32)
    st = codec._lib.rb200_kdtree_search_radius(codec._h, rb.abi.ptr(cloud), len(cloud), rb.abi.ptr(queries), len(queries), radius2,
                                               max_results, 1 if sorted_ else 0, rb.abi.ptr(idx), rb.abi.ptr(dist), rb.abi.ptr(cnt))
    assert st == 0, codec._lib.rb200_error_string(codec._h).decode()
    return idx, dist, cnt


def test_radius_search_traversal_order_and_sorted_cut(rb, codec, checker_backend):
    """PCCKdTree::searchRadius (the non-grid smoothPointCloud's query): first nanoflann's traversal order on its own
    (SearchParams::sorted = false), then with std::sort (the vendored IndexDist_Sorter: distance, then index) and the cut to
    64 results, which falls inside a shell of equal distances for nearly every query"""
    rng = np.random.default_rng(7)
    surf = rng.integers(0, 120, size=(40000, 3)).astype(np.int16)
    surf[:, 2] = (20 + 10 * np.sin(surf[:, 0] / 9.0) + 8 * np.cos(surf[:, 1] / 7.0)).astype(np.int16) + rng.integers(0, 3, 40000)
    dense = rng.integers(0, 24, size=(20000, 3)).astype(np.int16)  # ~1.4 points per lattice site: hundreds inside the ball
    for cloud, r2 in ((surf, 64.0), (surf, 30.5), (dense, 17.0), (dense, 64.0)):
        q = cloud[rng.integers(0, len(cloud), 600)]
        wi, wd, wc = checker_backend.knn_radius(cloud, q, r2, 700, sorted_=False)
        gi, gd, gc = gpu_radius(rb, codec, cloud, q, r2, 700, False)
        assert np.array_equal(gc, wc), "number of points inside the radius"
        bad = np.nonzero((gi != wi).any(axis=1))[0]
        assert len(bad) == 0, f"traversal order differs for {len(bad)}/{len(q)} queries, first {bad[0]}: " \
                              f"got {gi[bad[0]][:12].tolist()} want {wi[bad[0]][:12].tolist()}"
        assert np.array_equal(gd, wd)
        wi, wd, wc = checker_backend.knn_radius(cloud, q, r2, 64, sorted_=True)
        gi, gd, gc = gpu_radius(rb, codec, cloud, q, r2, 64, True)
        assert np.array_equal(gd, wd)
        bad = np.nonzero((gi != wi).any(axis=1))[0]
        assert len(bad) == 0, f"sorted order differs for {len(bad)}/{len(q)} queries, first {bad[0]}: " \
                              f"got {gi[bad[0]].tolist()} want {wi[bad[0]].tolist()} dist {wd[bad[0]].tolist()}"
        assert (wc > 64).mean() > 0.5
