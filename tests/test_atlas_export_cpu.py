"""Patch-table export: rabbit-transcoding_b200/host/rb200_atlas_export.h turns the reference's tile
containers (PCCPatch / PCCEomPatch / PCCRawPointsPatch, as PCCDecoder::createPatchFrameDataStructure leaves them,
PCCDecoder.cpp:869-1238) into the flat rows of include/rabbit_b200.h.  The harness builds those containers from rows
with the reference's own setters, exports them again, and the rows must come back unchanged — field for field."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref():
    from oracle import checker
    if not checker.have_reference():
        pytest.skip("oracle/_ref/librabbit_ref.so not built")
    return checker.Reference()


def roundtrip(rb, ref, g):
    abi = rb.abi
    f = ref.lib.ref_atlas_export_roundtrip
    f.argtypes = [C.POINTER(abi.Params), C.c_int, C.POINTER(abi.Atlas)] + [C.c_void_p] * 7
    F = g.n_frames
    out_p = np.zeros(len(g.patches), abi.PATCH_DTYPE)
    out_po = np.zeros(F + 1, np.int32)
    ne = len(g.eom_patches) if g.eom_patches is not None else 0
    nr = len(g.raw_patches) if g.raw_patches is not None else 0
    out_e, out_eo = np.zeros(max(ne, 1), abi.EOM_DTYPE), np.zeros(F + 1, np.int32)
    out_m = np.zeros(max(1, len(g.eom_members) if g.eom_members is not None else 1), np.int32)
    out_r, out_ro = np.zeros(max(nr, 1), abi.RAW_DTYPE), np.zeros(F + 1, np.int32)
    at = g.atlas_struct()
    n = f(C.byref(g.params), F, C.byref(at), abi.ptr(out_p), abi.ptr(out_po), abi.ptr(out_e), abi.ptr(out_eo), abi.ptr(out_m),
          abi.ptr(out_r), abi.ptr(out_ro))
    assert n == len(g.patches)
    assert np.array_equal(out_p, g.patches), "patch rows changed in the export"
    assert np.array_equal(out_po, g.patch_offset)
    if ne:
        assert np.array_equal(out_e[:ne], g.eom_patches) and np.array_equal(out_eo, g.eom_offset)
        assert np.array_equal(out_m[:len(g.eom_members)], g.eom_members)
    if nr:
        assert np.array_equal(out_r[:nr], g.raw_patches) and np.array_equal(out_ro, g.raw_offset)


def test_export_regular_patches_all_orientations(rb, ref):
    g = rb.synthetic.generate_gof(n_frames=3, bitdepth=8, width=256, scale=0.9, seed=5, orientations=tuple(range(9)))
    assert len(g.patches) > 20
    roundtrip(rb, ref, g)


def test_export_lod_and_additional_planes(rb, ref):
    g = rb.synthetic.make_lod_and_oblique(rb.synthetic.generate_gof(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=6), seed=6)
    assert set(int(a) for a in g.patches["axis_of_additional_plane"]) >= {1, 2, 3} and (g.patches["lod_x"] == 2).any()
    roundtrip(rb, ref, g)


def test_export_eom_and_raw_patches(rb, ref):
    roundtrip(rb, ref, rb.synthetic.generate_gof(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=7, eom=True,
                                                 geometry_smoothing=False, color_smoothing=False, transfer_filter=0))
    roundtrip(rb, ref, rb.synthetic.generate_gof(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=8, raw_points=500,
                                                 transfer_filter=0))
