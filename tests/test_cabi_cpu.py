"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol include/rabbit_b200.h
declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(rb):
    if not os.path.exists(rb.abi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return rb.abi.load_library()


def test_header_symbols_exported(rb, lib):
    hdr = open(os.path.join(ROOT, "include", "rabbit_b200.h")).read()
    declared = set(re.findall(r"\b(rb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(rb.abi.EXPORTED_SYMBOLS), declared ^ set(rb.abi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by librabbit_b200.so"
    assert lib.rb200_abi_version() == 1


def test_struct_sizes_match_header(rb):
    # sizes computed from the header layout (all int32 / double / pointer fields, natural alignment)
    assert C.sizeof(rb.abi.Patch) == 17 * 4
    assert C.sizeof(rb.abi.Params) == 31 * 4 + 4 + 4 * 8
    assert C.sizeof(rb.abi.Frames) == 3 * 8
    assert C.sizeof(rb.abi.Atlas) == 7 * 8
    assert C.sizeof(rb.abi.CloudHost) == 6 * 8
    assert C.sizeof(rb.abi.FrameCounts) == 6 * 8
    assert C.sizeof(rb.abi.MetricsParams) == 8 * 4
    assert C.sizeof(rb.abi.CloudView) == 4 * 8


def test_no_cpu_fallback_without_gpu(rb, lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.rb200_create(0, C.byref(h)) != 0
    with pytest.raises(Exception):
        rb.codec.PCCCodecB200(device=0)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under the product package may import, link or dlopen it"""
    pkg = os.path.join(ROOT, "rabbit-transcoding_b200")
    pat = re.compile(r"from\s+oracle|import\s+oracle|import\s+checker|liboracle|librabbit_ref|oracle/_ref|oracle\.h")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not pat.search(src), f"{fn} references the oracle"
