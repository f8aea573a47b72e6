"""convertYUV16ToRGB8 (PccLibCommon/include/PCCPointSet.h:133-166): the production kernel evaluates a short sequence
of double operations and falls back to the reference's sequence near exact rounding ties.  Checked here on random colours, on the edge
values of every channel and on constructed exact ties (k + 1/2), against numpy doubles evaluated operation by
operation like the reference, and against the kernel's own double path."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DEN = 257000000


def reference_f64(yuv):
    c = yuv.astype(np.float64)
    w = 1.0 / 65535.0
    y1 = np.minimum(np.maximum(w * c[:, 0], 0.0), 1.0)
    u1 = np.minimum(np.maximum(w * (c[:, 1] - 32768.0), -0.5), 0.5)
    v1 = np.minimum(np.maximum(w * (c[:, 2] - 32768.0), -0.5), 0.5)
    r = y1 + 1.57480 * v1
    g = (y1 - 0.18733 * u1) - 0.46813 * v1
    b = y1 + 1.85563 * u1

    def rnd(a):  # C round(): half away from zero, exact (no a + 0.5), then PCCClip
        a = a * 255.0
        t = np.trunc(a)
        t = t + np.where(np.abs(a - t) >= 0.5, np.sign(a), 0.0)
        return np.clip(t, 0.0, 255.0)
    return np.stack([rnd(r), rnd(g), rnd(b)], axis=1).astype(np.uint8)


def chroma2(v):
    return np.maximum(2 * (v.astype(np.int64) - 32768), -65535)


def ties():
    """colours whose exact value of some channel is k + 1/2"""
    out = []
    allc = np.arange(65536, dtype=np.int64)
    c2 = chroma2(allc)
    k = np.arange(256, dtype=np.int64)
    want = (2 * k + 1) * (DEN // 2)
    for coef, ch in ((787400, 2), (927815, 1)):  # r depends on (Y, V), b on (Y, U)
        cand = allc[(coef * c2 - 500000) % 1000000 == 0]
        for cv in cand:
            num = want - coef * int(chroma2(np.array([cv]))[0])
            ok = (num % 1000000 == 0) & (num >= 0) & (num <= 65535 * 1000000)
            for y in num[ok] // 1000000:
                t = [int(y), 12345, 54321]
                t[ch] = int(cv)
                out.append(t)
    rng = np.random.default_rng(5)
    for u in rng.integers(0, 65536, 600):  # g: fix U, search V and Y
        u2 = int(chroma2(np.array([u]))[0])
        rest = (-93665 * u2 - 234065 * c2)
        ok = (rest - 500000) % 1000000 == 0
        for v in allc[ok][:4]:
            num = want - int(rest[v])
            good = (num % 1000000 == 0) & (num >= 0) & (num <= 65535 * 1000000)
            for y in num[good] // 1000000:
                out.append([int(y), int(u), int(v)])
    return np.array(out, dtype=np.uint16)


def convert(codec, rb, yuv, f64):
    yuv = np.ascontiguousarray(yuv, dtype=np.uint16)
    rgb = np.zeros((len(yuv), 3), np.uint8)
    codec._check(codec._lib.rb200_debug_yuv16_to_rgb8(codec._h, rb.abi.ptr(yuv), len(yuv), rb.abi.ptr(rgb), int(f64)))
    return rgb


def test_integer_rgb8_equals_reference_doubles(rb, codec):
    rng = np.random.default_rng(11)
    edge = np.array([0, 1, 2, 127, 128, 255, 256, 257, 32766, 32767, 32768, 32769, 65279, 65280, 65534, 65535], np.uint16)
    grid = np.stack(np.meshgrid(edge, edge, edge, indexing="ij"), axis=-1).reshape(-1, 3)
    t = ties()
    assert len(t) > 100, "no exact ties constructed"
    # the constructed colours really are exact ties of some channel
    y, u2, v2 = 1000000 * t[:, 0].astype(np.int64), chroma2(t[:, 1]), chroma2(t[:, 2])
    nums = np.stack([y + 787400 * v2, y - 93665 * u2 - 234065 * v2, y + 927815 * u2], axis=1)
    assert ((2 * nums + DEN) % (2 * DEN) == 0).any(axis=1).all()
    yuv = np.concatenate([rng.integers(0, 65536, (4_000_000, 3)).astype(np.uint16), grid, t,
                          # neighbours of the ties: one step to either side in every channel
                          np.clip(t.astype(np.int64) + rng.integers(-1, 2, t.shape), 0, 65535).astype(np.uint16)])
    want = reference_f64(yuv)
    got_f64 = convert(codec, rb, yuv, True)
    got_int = convert(codec, rb, yuv, False)
    assert np.array_equal(got_f64, want), "double path differs from the reference arithmetic"
    assert np.array_equal(got_int, want), "short path differs from the reference arithmetic"
