"""CPU suite for the decoder-native ingest: the numpy restatement of PCCVideoDecoder's inverse colour
conversion (oracle/oracle_np.yuv420_to_yuv444) against the unmodified reference converter and against the committed
golden digests, for every upsampling filter and both bit depths."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_yuv420 as mg  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "yuv420.json")))


@pytest.mark.parametrize("case", mg.CASES, ids=lambda c: f"{c[1]}x{c[2]}_{c[3]}bit_f{c[4]}")
def test_restatement_matches_golden(case):
    from oracle import oracle_np
    seed, W, H, bd, f = case
    y, u, v = mg.frame(seed, W, H, bd)
    assert mg.digest(oracle_np.yuv420_to_yuv444(y, u, v, bd, f)) == GOLD[f"{seed}_{W}x{H}_{bd}bit_filter{f}"]


def test_reference_matches_golden_and_restatement():
    from oracle import checker, oracle_np
    if not checker.have_reference():
        pytest.skip("oracle/_ref not built")
    ref = checker.Reference()
    for seed, W, H, bd, f in mg.CASES[::3]:
        y, u, v = mg.frame(seed, W, H, bd)
        want = ref.yuv420_to_yuv444(y, u, v, bd, f)
        assert mg.digest(want) == GOLD[f"{seed}_{W}x{H}_{bd}bit_filter{f}"]
        assert np.array_equal(oracle_np.yuv420_to_yuv444(y, u, v, bd, f), want)


def test_decoder_planes_round_trip(rb):
    """synthetic.to_decoder_planes keeps the top bits of the 4:4:4 frames and the layout { Y, U, V } per frame"""
    g = rb.synthetic.generate_gof(n_frames=1, bitdepth=8, width=256, scale=0.9, seed=3)
    nat = rb.synthetic.to_decoder_planes(g, bitdepth=8, filt=0)
    p = g.params
    H, W, M = p.height, p.width, p.map_count_minus1 + 1
    a = g.attribute.reshape(1, M, 3, H, W)
    fr = nat["attribute"].reshape(1, M, -1)
    assert fr.dtype == np.uint8 and fr.shape[2] == H * W * 3 // 2
    assert np.array_equal(fr[0, 0, :H * W].reshape(H, W), a[0, 0, 0] >> 8)
    assert np.array_equal(fr[0, 1, H * W:H * W + (H // 2) * (W // 2)].reshape(H // 2, W // 2), a[0, 1, 1, ::2, ::2] >> 8)
    assert nat["geometry"].dtype == np.uint8 and np.array_equal(nat["geometry"].reshape(-1), g.geometry.reshape(-1))


def test_image_set_restatement_matches_reference():
    """PCCImage::set's rounding shift + clamp (decoder internal bit depth -> output bit depth)"""
    from oracle import checker, oracle_np
    if not checker.have_reference():
        pytest.skip("oracle/_ref not built")
    ref = checker.Reference()
    rng = np.random.default_rng(4)
    for shift in (0, 1, 2, 4):
        y = rng.integers(0, 1024, (16, 32)).astype(np.int16)
        u = rng.integers(0, 1024, (8, 16)).astype(np.int16)
        v = rng.integers(0, 1024, (8, 16)).astype(np.int16)
        y[0, :4] = [0, 1, 1022, 1023]
        got = ref.image_set_yuv420(y, u, v, shift)
        for a, b in zip(got, (y, u, v)):
            assert np.array_equal(a, oracle_np.image_set(b, shift)), f"shift {shift}"


BITDEPTH_CASES = [(10, 8, True), (10, 8, False), (8, 10, True), (8, 10, False), (8, 8, True), (16, 10, False), (10, 16, True),
                  (12, 10, True)]


def test_convert_bitdepth_restatement_matches_reference():
    """PCCImage::convertBitdepth (PCCImage.cpp:258-299) as run on every geometry video (PCCDecoder.cpp:148-149) and on
    the occupancy video (:119): shift / clamp / left shift with truncation to the sample type"""
    from oracle import checker, oracle_np
    if not checker.have_reference():
        pytest.skip("oracle/_ref not built")
    ref = checker.Reference()
    rng = np.random.default_rng(9)
    for bi, bo, msb in BITDEPTH_CASES:
        g = rng.integers(0, 1 << min(bi, 16), (8, 24)).astype(np.uint16)
        g[0, :3] = [0, (1 << min(bi, 16)) - 1, 1 << (min(bi, 16) - 1)]
        assert np.array_equal(ref.convert_bitdepth(g, bi, bo, msb), oracle_np.convert_bitdepth(g, bi, bo, msb)), (bi, bo, msb)
    for bo, msb in ((1, False), (1, True), (4, True), (8, False), (8, True)):
        o = rng.integers(0, 256, (8, 8)).astype(np.uint8)
        assert np.array_equal(ref.convert_bitdepth(o, 8, bo, msb), oracle_np.convert_bitdepth(o, 8, bo, msb)), (bo, msb)
