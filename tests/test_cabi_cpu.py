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
    assert lib.rb200_abi_version() == 3


def test_struct_sizes_match_header(rb):
    # sizes computed from the header layout (all int32 / double / pointer fields, natural alignment)
    assert C.sizeof(rb.abi.Patch) == 17 * 4
    assert C.sizeof(rb.abi.Params) == 32 * 4 + 4 * 8 + 4 * 4 + 2 * 8 + 4 * 4
    assert C.sizeof(rb.abi.Frames) == 5 * 8
    assert C.sizeof(rb.abi.Atlas) == 7 * 8
    assert C.sizeof(rb.abi.CloudHost) == 6 * 8
    assert C.sizeof(rb.abi.FrameCounts) == 6 * 8
    assert C.sizeof(rb.abi.MetricsParams) == 8 * 4
    assert C.sizeof(rb.abi.CloudView) == 4 * 8


def test_header_is_plain_c_and_struct_sizes_agree(rb, tmp_path):
    """include/rabbit_b200.h compiles as C99 (the drop-in boundary is a C ABI: plain pointers and sizes) and the
    ctypes mirror has the sizes the C compiler sees"""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    names = ["rb200_patch", "rb200_params", "rb200_frames", "rb200_frames_yuv420", "rb200_atlas", "rb200_cloud_host",
             "rb200_frame_counts", "rb200_metrics_params", "rb200_cloud_view", "rb200_metrics_result"]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "rabbit_b200.h"\nint main(void){' +
                   "".join(f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mirror = [rb.abi.Patch, rb.abi.Params, rb.abi.Frames, rb.abi.FramesYuv420, rb.abi.Atlas, rb.abi.CloudHost,
              rb.abi.FrameCounts, rb.abi.MetricsParams, rb.abi.CloudView, rb.abi.MetricsResult]
    for n, sz, m in zip(names, sizes, mirror):
        assert C.sizeof(m) == sz, f"{n}: header {sz} bytes, ctypes mirror {C.sizeof(m)}"


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
