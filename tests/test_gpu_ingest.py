"""GPU parity of the decoder-native ingest (rb200_gof_upload_yuv420): the inverse colour conversion of PCCVideoDecoder
(PccLibDecoder/source/PCCVideoDecoder.cpp:125-146 -> PCCInternalColorConverter::convertYUV420ToYUV444) on the device,
bit-exact against the committed golden digests of the reference converter, against the numpy restatement, and — through
the whole decoder sequence — against the reference run on the converted frames."""
import json
import os
import sys

import numpy as np
import pytest

from util import assert_cloud_equal

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_yuv420 as mg  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "yuv420.json")))


def _params(rb, W, H):
    g = rb.synthetic.generate_gof(n_frames=1, bitdepth=8, width=256, scale=0.9, seed=1)
    p = g.params
    p.width, p.height, p.map_count_minus1 = W, H, 0
    return g, p


def _convert_on_gpu(rb, codec, y, u, v, bd, f, dtype):
    """one 4:2:0 frame through rb200_gof_upload_yuv420, planes read back from HBM"""
    H, W = y.shape
    g, p = _params(rb, W, H)
    g.occupancy = np.zeros((1, H // p.occupancy_precision, W // p.occupancy_precision), np.uint8)
    g.patches = g.patches[:0]
    g.patch_offset = np.zeros(2, np.int32)
    native = {"bitdepth": bd, "filter": f,
              "geometry": np.ascontiguousarray(y.astype(dtype) if dtype == np.uint16 else (y & 0xFF).astype(np.uint8)).reshape(1, 1, H, W),
              "attribute": np.concatenate([y.reshape(-1), u.reshape(-1), v.reshape(-1)]).astype(dtype).reshape(1, 1, -1)}
    codec.uploadGofYuv420(g, native)
    geo, att = codec.getPlanes(0, 0)
    assert np.array_equal(geo, native["geometry"].reshape(H, W)), "geometry samples changed by the ingest"
    return att


@pytest.mark.parametrize("case", mg.CASES, ids=lambda c: f"{c[1]}x{c[2]}_{c[3]}bit_f{c[4]}")
def test_conversion_matches_reference_golden(rb, codec, case):
    seed, W, H, bd, f = case
    y, u, v = mg.frame(seed, W, H, bd)
    dtypes = (np.uint8, np.uint16) if bd == 8 else (np.uint16,)
    for dt in dtypes:
        att = _convert_on_gpu(rb, codec, y, u, v, bd, f, dt)
        assert mg.digest(att) == GOLD[f"{seed}_{W}x{H}_{bd}bit_filter{f}"], f"4:2:0 -> 4:4:4 differs from the reference ({dt.__name__} samples)"


def test_conversion_full_size_frame_matches_restatement(rb, codec):
    """a vox10-sized frame (tiles, halos and image borders of the CUDA tiling) against the numpy restatement"""
    from oracle import oracle_np
    rng = np.random.default_rng(9)
    W, H = 1280, 1296
    yy, xx = np.mgrid[0:H, 0:W]
    y = ((np.sin(xx * 0.02) * np.cos(yy * 0.015) * 100 + 128) + rng.normal(0, 6, (H, W))).clip(0, 255).astype(np.uint16)
    u = rng.integers(0, 256, (H // 2, W // 2)).astype(np.uint16)
    v = ((np.sin(xx[::2, ::2] * 0.05) * 120 + 128)).clip(0, 255).astype(np.uint16)
    for f in (0, 4, 7):
        att = _convert_on_gpu(rb, codec, y, u, v, 8, f, np.uint8)
        assert np.array_equal(att, oracle_np.yuv420_to_yuv444(y, u, v, 8, f)), f"filter {f}"


@pytest.mark.parametrize("bd,filt,dt", [(8, 0, np.uint8), (8, 4, np.uint8), (10, 2, np.uint16)])
def test_decode_gof_from_decoder_planes(rb, codec, checker_backend, bd, filt, dt):
    """the whole decoder sequence fed with decoder-native planes == the reference fed with the frames its own converter
    (restated in numpy, pinned above) makes of those planes"""
    from oracle import oracle_np
    kw = dict(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=21, transfer_filter=0)
    gof = rb.synthetic.generate_gof(**kw)
    native = rb.synthetic.to_decoder_planes(gof, bitdepth=bd, filt=filt, sample_dtype=dt)
    p = gof.params
    F, M, H, W = gof.n_frames, p.map_count_minus1 + 1, p.height, p.width
    conv = np.empty((F, M, 3, H, W), np.uint16)
    fr = native["attribute"].reshape(F, M, -1).astype(np.uint16)
    q = (H // 2) * (W // 2)
    for f in range(F):
        for m in range(M):
            conv[f, m] = oracle_np.yuv420_to_yuv444(fr[f, m, :H * W].reshape(H, W), fr[f, m, H * W:H * W + q].reshape(H // 2, W // 2),
                                                    fr[f, m, H * W + q:].reshape(H // 2, W // 2), bd, filt)
    gof.attribute = np.ascontiguousarray(conv.reshape(gof.attribute.shape))
    ref = checker_backend.run_gof(gof, keep=("rgb8",))
    codec.uploadGofYuv420(gof, native)
    codec.decodeGof()
    counts = codec.frameCounts()
    for f in range(F):
        assert_cloud_equal(codec.getPointCloud(f, counts), ref.cloud(f, "rgb8"), f"frame {f}")
    _, att = codec.getPlanes(1, M - 1)
    assert np.array_equal(att, conv[1, M - 1])


def test_upload_yuv420_error_paths(rb, codec):
    g, p = _params(rb, 64, 32)
    g.occupancy = np.zeros((1, 32 // p.occupancy_precision, 64 // p.occupancy_precision), np.uint8)
    g.patches = g.patches[:0]
    g.patch_offset = np.zeros(2, np.int32)
    nat = {"bitdepth": 9, "filter": 0, "geometry": np.zeros((1, 1, 32, 64), np.uint8), "attribute": np.zeros((1, 1, 64 * 32 * 3 // 2), np.uint8)}
    with pytest.raises(rb.codec.RabbitError):
        codec.uploadGofYuv420(g, nat)
    nat["bitdepth"], nat["filter"] = 8, 8
    with pytest.raises(rb.codec.RabbitError):
        codec.uploadGofYuv420(g, nat)


@pytest.mark.parametrize("shift", [1, 2])
def test_ingest_applies_image_set_shift(rb, codec, shift):
    """decoder planes at the codec's internal bit depth: PCCImage::set's rounding shift + clamp, then the conversion"""
    from oracle import oracle_np
    rng = np.random.default_rng(17 + shift)
    W, H = 96, 64
    y = rng.integers(0, 1024, (H, W)).astype(np.uint16)
    u = rng.integers(0, 1024, (H // 2, W // 2)).astype(np.uint16)
    v = rng.integers(0, 1024, (H // 2, W // 2)).astype(np.uint16)
    y[0, :4] = [0, 1, 1022, 1023]
    bd = 10 - shift if 10 - shift in (8, 10) else 8  # stored samples have 10 - shift bits; the converter is told 8 or 10
    g, p = _params(rb, W, H)
    g.occupancy = np.zeros((1, H // p.occupancy_precision, W // p.occupancy_precision), np.uint8)
    g.patches = g.patches[:0]
    g.patch_offset = np.zeros(2, np.int32)
    native = {"bitdepth": bd, "filter": 3, "geometry_shift": shift, "attribute_shift": shift,
              "geometry": np.ascontiguousarray(y).reshape(1, 1, H, W),
              "attribute": np.concatenate([y.reshape(-1), u.reshape(-1), v.reshape(-1)]).reshape(1, 1, -1)}
    codec.uploadGofYuv420(g, native)
    geo, att = codec.getPlanes(0, 0)
    ys, us, vs = (oracle_np.image_set(a, shift) for a in (y, u, v))
    assert np.array_equal(geo, ys)
    assert np.array_equal(att, oracle_np.yuv420_to_yuv444(ys, us, vs, bd, 3))


@pytest.mark.parametrize("case", [(10, 8, True), (10, 8, False), (8, 10, True), (12, 10, False), (10, 16, True)],
                         ids=lambda c: f"{c[0]}to{c[1]}_{'msb' if c[2] else 'lsb'}")
def test_ingest_applies_convert_bitdepth_to_geometry(rb, codec, checker_backend, case):
    """PCCImage::convertBitdepth runs on every decoded geometry video (PCCDecoder.cpp:148-149); here it is fused into the
    geometry ingest kernel and must equal the reference's own function on the same plane"""
    bi, bo, msb = case
    rng = np.random.default_rng(31 + bi + bo)
    W, H = 96, 64
    y = rng.integers(0, 1 << bi, (H, W)).astype(np.uint16)
    y[0, :3] = [0, (1 << bi) - 1, 1 << (bi - 1)]
    g, p = _params(rb, W, H)
    g.occupancy = np.zeros((1, H // p.occupancy_precision, W // p.occupancy_precision), np.uint8)
    g.patches = g.patches[:0]
    g.patch_offset = np.zeros(2, np.int32)
    native = {"bitdepth": 8, "filter": 0, "geometry": np.ascontiguousarray(y).reshape(1, 1, H, W),
              "attribute": np.zeros((1, 1, H * W * 3 // 2), np.uint8), "geometry_bitdepth": (bi, bo, 1 if msb else 0)}
    codec.uploadGofYuv420(g, native)
    geo, _ = codec.getPlanes(0, 0)
    assert np.array_equal(geo, checker_backend.convert_bitdepth(y, bi, bo, msb))


def test_ingest_convert_bitdepth_of_the_occupancy_video(rb, codec, checker_backend):
    """the occupancy video goes through convertBitdepth( 8, occupancy2DBitdepth, msbAlign ) (PCCDecoder.cpp:119) before
    generateOccupancyMap: a decoder handing over 8-bit msb-aligned samples of a 1-bit map decodes to the same clouds as the
    reference fed with the plane its own convertBitdepth produces"""
    kw = dict(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=55, transfer_filter=0)
    g = rb.synthetic.generate_gof(**kw)
    native = rb.synthetic.to_decoder_planes(g, bitdepth=8, filt=0)
    raw_occ = (g.occupancy.astype(np.uint8) << 7) | np.random.default_rng(3).integers(0, 64, g.occupancy.shape).astype(np.uint8)
    want_occ = np.stack([checker_backend.convert_bitdepth(o, 8, 1, True) for o in raw_occ])
    assert np.array_equal(want_occ, g.occupancy)
    # expected attribute planes: the reference's own conversion of the native frames
    F, M, H, W = g.n_frames, g.params.map_count_minus1 + 1, g.params.height, g.params.width
    fr = native["attribute"].reshape(F, M, -1)
    q = (H // 2) * (W // 2)
    conv = np.zeros((F, M, 3, H, W), np.uint16)
    for f in range(F):
        for m in range(M):
            conv[f, m] = checker_backend.yuv420_to_yuv444(fr[f, m, :H * W].reshape(H, W), fr[f, m, H * W:H * W + q].reshape(H // 2, W // 2),
                                                          fr[f, m, H * W + q:].reshape(H // 2, W // 2), 8, 0)
    g.attribute = np.ascontiguousarray(conv.reshape(g.attribute.shape))
    ref = checker_backend.run_gof(g, keep=("rgb8",))
    g2 = rb.synthetic.generate_gof(**kw)
    g2.occupancy = np.ascontiguousarray(raw_occ)
    native["occupancy_bitdepth"] = (1, 1)
    codec.uploadGofYuv420(g2, native)
    codec.decodeGof()
    counts = codec.frameCounts()
    for f in range(F):
        assert_cloud_equal(codec.getPointCloud(f, counts), ref.cloud(f, "rgb8"), f"frame {f}")


def _nv12_surfaces(torch, gof, native, pad=96, p010=False):
    """the native planar frames as pitched NV12 (or P010: 10-bit samples in the high bits of 16-bit words) surfaces in
    device memory, one allocation per surface with `pad` extra bytes per row — what NVDEC hands to libav"""
    F, M = gof.n_frames, gof.params.map_count_minus1 + 1
    H, W, pr = gof.params.height, gof.params.width, gof.params.occupancy_precision
    keep, out = [], dict(occupancy=[], geometry=[], attribute=[], sample_bytes=2 if p010 else 1,
                         sample_lsb_shift=6 if p010 else 0, bitdepth=native["bitdepth"], filter=native["filter"])
    dt = torch.int16 if p010 else torch.uint8
    bpe = 2 if p010 else 1

    def surface(plane2d, dtype, elem):
        h, w = plane2d.shape
        pitch = w * elem + pad
        buf = torch.zeros((h, pitch), dtype=torch.uint8, device="cuda")
        src = torch.from_numpy(np.ascontiguousarray(plane2d)).cuda()
        buf[:, :w * elem] = src.view(torch.uint8).reshape(h, w * elem)
        keep.append(buf)
        return buf.data_ptr(), pitch
    for f in range(F):
        out["occupancy"].append(surface(gof.occupancy[f], torch.uint8, 1))
    geo = native["geometry"].reshape(F * M, H, W)
    att = native["attribute"].reshape(F * M, -1)
    q = (H // 2) * (W // 2)
    for i in range(F * M):
        g = geo[i].astype(np.uint16) << 6 if p010 else geo[i]
        out["geometry"].append(surface(g, dt, bpe))
        y = att[i, :H * W].reshape(H, W)
        u = att[i, H * W:H * W + q].reshape(H // 2, W // 2)
        v = att[i, H * W + q:].reshape(H // 2, W // 2)
        uv = np.empty((H // 2, W), att.dtype)
        uv[:, 0::2], uv[:, 1::2] = u, v
        if p010:
            y, uv = y.astype(np.uint16) << 6, uv.astype(np.uint16) << 6
        out["attribute"].append(surface(y, dt, bpe) + surface(uv, dt, bpe))
    out["_keep"] = keep
    return out


@pytest.mark.parametrize("p010", [False, True], ids=["nv12", "p010"])
def test_upload_nv12_device_surfaces_equals_planar_upload(rb, codec, p010):
    """pitched NV12 / P010 surfaces resident in device memory (RABBIT's --useCuda decode, PCCTranscoder.cpp:693-704) are
    gathered on the device and decode to exactly what the planar 4:2:0 upload of the same samples decodes to"""
    import torch
    kw = dict(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=57, transfer_filter=1)
    g = rb.synthetic.generate_gof(**kw)
    bd = 10 if p010 else 8
    native = rb.synthetic.to_decoder_planes(g, bitdepth=bd, filt=2, sample_dtype=np.uint16 if p010 else np.uint8)
    codec.uploadGofYuv420(g, native)
    codec.decodeGof()
    counts = codec.frameCounts()
    want = [codec.getPointCloud(f, counts) for f in range(g.n_frames)]
    want_planes = codec.getPlanes(1, 1)
    surf = _nv12_surfaces(torch, g, native, p010=p010)
    torch.cuda.synchronize()
    st0 = codec.stats(reset=True)
    codec.uploadGofNv12(g, surf)
    up = codec.stats(reset=True)
    assert up.h2d_bytes < 200000, up.h2d_bytes  # only the surface table and the patch tables cross PCIe
    codec.decodeGof()
    counts = codec.frameCounts()
    got_planes = codec.getPlanes(1, 1)
    assert np.array_equal(got_planes[0], want_planes[0]) and np.array_equal(got_planes[1], want_planes[1])
    for f in range(g.n_frames):
        assert_cloud_equal(codec.getPointCloud(f, counts), want[f], f"nv12 frame {f}")
