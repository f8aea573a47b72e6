"""GPU path against the committed golden fixtures (tests/golden/golden.json, generated from the unmodified reference
by tests/golden/make_golden.py): needs no checker on the box."""
import hashlib
import json
import os
import struct
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def f32(hexstr):
    return struct.unpack("<f", bytes.fromhex(hexstr))[0]


@pytest.mark.parametrize("name", sorted(GOLD))
def test_decode_and_metrics_match_golden(rb, codec, name):
    kw = dict(GOLD[name]["args"])
    g = make_golden.make_gof(rb, kw)
    codec.uploadGof(g)
    codec.decodeGof()
    counts = codec.frameCounts()
    mp = rb.metrics.default_parameters(resolution=float((1 << kw["bitdepth"]) - 1))
    met = rb.metrics.PCCMetricsB200(codec)
    met.setParameters(mp)
    res = met.compute(g.sources, [None] * g.n_frames, g.sources)
    for f, want in enumerate(GOLD[name]["frames"]):
        c = codec.getPointCloud(f, counts)
        assert (counts[f].total, counts[f].regular, counts[f].raw, counts[f].smoothed, counts[f].recolored) == \
            (want["total"], want["regular"], want["raw"], want["smoothed"], want["recolored"]), f"{name} frame {f} counts"
        md5 = hashlib.md5(c["positions"].tobytes() + c["colors"].tobytes()).hexdigest()
        assert md5 == want["md5_ordered"], f"{name} frame {f}: ordered MD5 (PCCPointSet3::computeChecksum) differs"
        assert make_golden.cloud_digest(c) == want["stages"]["rgb8"], f"{name} frame {f}: final cloud digest differs"
        assert [res[f].source_points, res[f].source_after_dedup, res[f].rec_points, res[f].rec_after_dedup] == want["dedup"]
        for t in ("q1", "q2", "qf"):
            q, w = getattr(res[f], t), want["metrics"][t]
            assert q.c2c_mse == f32(w["c2c_mse"])
            for got, ref in ((q.c2c_psnr, w["c2c_psnr"]), (q.c2p_psnr, w["c2p_psnr"])) + tuple(
                    (q.color_psnr[k], w["color_psnr"][k]) for k in range(3)):
                ref = f32(ref)
                assert (got == ref) or abs(got - ref) <= 1e-6, f"{name} frame {f} {t}: {got} vs {ref}"


def test_gathered_psnr_equals_library_psnr(rb, codec):
    """rabbit_transcoding_b200.dist recomputes the PSNRs from the exchanged accumulators: must equal rb200_metrics'"""
    g = rb.synthetic.generate_gof(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=105, transfer_filter=0)
    codec.uploadGof(g)
    codec.decodeGof()
    mp = rb.metrics.default_parameters(resolution=255.0)
    met = rb.metrics.PCCMetricsB200(codec)
    met.setParameters(mp)
    res = met.compute(g.sources, [None] * g.n_frames, g.sources)
    frames, mean = rb.dist.gather_metrics({f: r for f, r in enumerate(res)}, g.n_frames, 255.0)
    for f, r in enumerate(res):
        assert frames[f]["qf"]["c2c_psnr"] == np.float32(r.qf.c2c_psnr)
        assert frames[f]["qf"]["c2p_psnr"] == np.float32(r.qf.c2p_psnr)
        assert frames[f]["qf"]["color_psnr"][0] == np.float32(r.qf.color_psnr[0])
    assert abs(mean["d1_psnr"] - np.mean([r.qf.c2c_psnr for r in res])) < 1e-5
