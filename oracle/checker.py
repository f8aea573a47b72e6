"""oracle/checker.py — Python front-end of the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Two back-ends with the same interface:
  * `Reference`  : oracle/_ref/librabbit_ref.so — the UNMODIFIED reference sources compiled from
                   /root/reference by oracle/Makefile (`make ref`), driven by oracle/ref_harness.cpp.
  * `DropIn`     : oracle/_ref/librabbit_dropin.so — the same reference objects and harness with the hot member
                   functions replaced by rabbit-transcoding_b200/host/PCCCodecB200.cpp (the drop-in boundary under test).
  The CPU restatement of the algorithm lives in oracle/oracle_np.py (numpy), pinned against `Reference` by
  tests/test_oracle_port_cpu.py and against tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package never does.
"""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import rabbit_transcoding_b200 as rb  # noqa: E402

abi = rb.abi
REF_LIB = os.path.join(_HERE, "_ref", "librabbit_ref.so")
PORT_LIB = os.path.join(_HERE, "liboracle.so")
DROPIN_LIB = os.path.join(_HERE, "_ref", "librabbit_dropin.so")
STAGES = ("reconstruct", "smooth_geometry", "transfer_colors", "smooth_color", "rgb8")


def have_reference():
    return os.path.exists(REF_LIB)


def have_dropin():
    return os.path.exists(DROPIN_LIB)


def have_port():
    return os.path.exists(PORT_LIB)


class _Run:
    """result handle of one GOF run"""

    def __init__(self, backend, handle, n_frames):
        self._b, self._h, self.n_frames = backend, handle, n_frames

    def __del__(self):
        if self._h:
            self._b._lib_call("gof_free", self._h)
            self._h = None

    def counts(self, f):
        c = abi.FrameCounts()
        self._b._lib_call("gof_counts", self._h, f, C.byref(c))
        return c

    def cloud(self, f, stage):
        s = STAGES.index(stage) if isinstance(stage, str) else stage
        n = self._b._lib_call("gof_count", self._h, f, s)
        if n < 0:
            raise KeyError(f"stage {stage} not kept")
        npart = self._b._lib_call("gof_partition_size", self._h, f)
        np2p = self._b._lib_call("gof_point_to_pixel_size", self._h, f)
        out = dict(positions=np.zeros((n, 3), np.int16), colors16=np.zeros((n, 3), np.uint16),
                   colors=np.zeros((n, 3), np.uint8), boundary_types=np.zeros(n, np.uint16),
                   partition=np.zeros(npart, np.uint32), point_to_pixel=np.zeros((np2p, 3), np.uint32))
        h = abi.CloudHost(*[abi.ptr(out[k]) for k, _ in abi.CloudHost._fields_])
        self._b._lib_call("gof_fetch", self._h, f, s, C.byref(h))
        return out

    def md5(self, f, canonical=False):
        d = (C.c_uint8 * 16)()
        self._b._lib_call("gof_md5", self._h, f, 1 if canonical else 0, d)
        return bytes(d).hex()

    def block_to_patch(self, f, params):
        a = np.zeros((params.height // params.occupancy_resolution, params.width // params.occupancy_resolution), np.uint32)
        self._b._lib_call("gof_block_to_patch", self._h, f, abi.ptr(a))
        return a

    def occupancy(self, f, params):
        a = np.zeros((params.height, params.width), np.uint8)
        self._b._lib_call("gof_occupancy", self._h, f, abi.ptr(a))
        return a

    def time_ms(self, f, which):
        return self._b._lib_call("gof_time_ms", self._h, f, which)


class _Backend:
    prefix = None

    def __init__(self, path):
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not built (cd oracle && make {'ref' if 'ref' in path else 'port'})")
        self.lib = C.CDLL(path)
        p = self.prefix
        L = self.lib
        getattr(L, p + "gof_run").restype = C.c_void_p
        getattr(L, p + "gof_run").argtypes = [C.POINTER(abi.Params), C.c_int, C.POINTER(abi.Frames),
                                              C.POINTER(abi.Atlas), C.c_uint32, C.c_int, C.c_int, C.c_int]
        getattr(L, p + "gof_free").argtypes = [C.c_void_p]
        getattr(L, p + "gof_free").restype = None
        for n in ("gof_count", "gof_partition_size", "gof_point_to_pixel_size"):
            getattr(L, p + n).restype = C.c_int64
        getattr(L, p + "gof_count").argtypes = [C.c_void_p, C.c_int, C.c_int]
        getattr(L, p + "gof_partition_size").argtypes = [C.c_void_p, C.c_int]
        getattr(L, p + "gof_point_to_pixel_size").argtypes = [C.c_void_p, C.c_int]
        getattr(L, p + "gof_fetch").argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.CloudHost)]
        getattr(L, p + "gof_counts").argtypes = [C.c_void_p, C.c_int, C.POINTER(abi.FrameCounts)]
        getattr(L, p + "gof_counts").restype = None
        getattr(L, p + "gof_md5").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        getattr(L, p + "gof_md5").restype = None
        getattr(L, p + "gof_block_to_patch").argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        getattr(L, p + "gof_occupancy").argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        getattr(L, p + "gof_time_ms").argtypes = [C.c_void_p, C.c_int, C.c_int]
        getattr(L, p + "gof_time_ms").restype = C.c_double
        getattr(L, p + "metrics").argtypes = [C.POINTER(abi.MetricsParams), C.POINTER(abi.CloudView),
                                              C.POINTER(abi.CloudView), C.POINTER(abi.CloudView),
                                              C.POINTER(abi.MetricsResult), C.POINTER(C.c_double)]
        getattr(L, p + "remove_duplicates").argtypes = [C.POINTER(abi.CloudView), C.c_int, C.c_void_p, C.c_void_p]
        getattr(L, p + "remove_duplicates").restype = C.c_int64
        getattr(L, p + "knn").argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]

    def _lib_call(self, name, *a):
        return getattr(self.lib, self.prefix + name)(*a)

    def run_gof(self, gof, keep=("reconstruct", "smooth_geometry", "transfer_colors", "smooth_color", "rgb8"),
                threads=1, canonical_md5=False, quiet=True, params=None):
        mask = 0
        for k in keep:
            mask |= 1 << STAGES.index(k)
        p = params if params is not None else gof.params
        fr, at = gof.frames_struct(), gof.atlas_struct()
        set_plr = getattr(self.lib, self.prefix + "gof_set_plr", None)
        plr = abi.plr_struct(gof.plr) if getattr(gof, "plr", None) is not None else None
        if set_plr is not None:
            set_plr.argtypes, set_plr.restype = [C.c_void_p], None
            set_plr(C.byref(plr) if plr is not None else None)
        elif plr is not None:
            raise NotImplementedError("this checker backend has no point local reconstruction")
        h = self._lib_call("gof_run", C.byref(p), gof.n_frames, C.byref(fr), C.byref(at), mask, threads,
                           1 if canonical_md5 else 0, 1 if quiet else 0)
        if not h:
            raise RuntimeError("oracle run failed")
        return _Run(self, h, gof.n_frames)

    @staticmethod
    def view(positions, colors=None, normals=None):
        positions = np.ascontiguousarray(positions, np.int16)
        keep = [positions]
        v = abi.CloudView()
        v.positions = abi.ptr(positions)
        v.count = positions.shape[0]
        if colors is not None:
            colors = np.ascontiguousarray(colors, np.uint8)
            keep.append(colors)
            v.colors = abi.ptr(colors)
        if normals is not None:
            normals = np.ascontiguousarray(normals, np.float32)
            keep.append(normals)
            v.normals = abi.ptr(normals)
        v._keep = keep
        return v

    def metrics(self, mparams, src, rec, normals=None):
        """src / rec / normals: dict(positions, colors[, normals]); returns (MetricsResult, ms)"""
        vs = self.view(src["positions"], src.get("colors"))
        vr = self.view(rec["positions"], rec.get("colors"))
        vn = self.view(normals["positions"], None, normals["normals"]) if normals is not None else None
        out = abi.MetricsResult()
        ms = C.c_double(0)
        self._lib_call("metrics", C.byref(mparams), C.byref(vs), C.byref(vr), C.byref(vn) if vn is not None else None,
                       C.byref(out), C.byref(ms))
        return out, ms.value

    def remove_duplicates(self, positions, colors, drop=2):
        v = self.view(positions, colors)
        n = positions.shape[0]
        op = np.zeros((n, 3), np.int16)
        oc = np.zeros((n, 3), np.uint8)
        m = self._lib_call("remove_duplicates", C.byref(v), drop, abi.ptr(op), abi.ptr(oc) if colors is not None else None)
        return op[:m].copy(), (oc[:m].copy() if colors is not None else None)

    def write_ply(self, positions, colors, path):
        """PCCPointSet3::write through oracle/_ref/ref_ply_tool (a separate process, see ref_ply_tool.cpp)"""
        import subprocess
        import tempfile
        tool = os.path.join(_HERE, "_ref", "ref_ply_tool")
        with tempfile.TemporaryDirectory() as d:
            np.ascontiguousarray(positions, np.int16).tofile(os.path.join(d, "p"))
            np.ascontiguousarray(colors, np.uint8).tofile(os.path.join(d, "c"))
            return subprocess.call([tool, "write", str(len(positions)), os.path.join(d, "p"), os.path.join(d, "c"), path])

    def read_ply(self, path, cap=None):
        import subprocess
        import tempfile
        tool = os.path.join(_HERE, "_ref", "ref_ply_tool")
        with tempfile.TemporaryDirectory() as d:
            subprocess.check_call([tool, "read", path, os.path.join(d, "p"), os.path.join(d, "c")], stdout=subprocess.DEVNULL)
            return (np.fromfile(os.path.join(d, "p"), np.int16).reshape(-1, 3),
                    np.fromfile(os.path.join(d, "c"), np.uint8).reshape(-1, 3))

    def knn(self, cloud, queries, k):
        cloud = np.ascontiguousarray(cloud, np.int16)
        queries = np.ascontiguousarray(queries, np.int16)
        idx = np.zeros((len(queries), k), np.int64)
        dist = np.zeros((len(queries), k), np.float64)
        self._lib_call("knn", abi.ptr(cloud), len(cloud), abi.ptr(queries), len(queries), k, abi.ptr(idx), abi.ptr(dist))
        return idx, dist


class Reference(_Backend):
    prefix = "ref_"

    def knn_radius(self, cloud, queries, radius2, max_results, sorted_=True):
        """PCCKdTree::searchRadius; sorted_ = False: nanoflann's traversal order (no std::sort)"""
        f = self.lib.ref_knn_radius
        f.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        cloud = np.ascontiguousarray(cloud, np.int16)
        queries = np.ascontiguousarray(queries, np.int16)
        idx = np.zeros((len(queries), max_results), np.int64)
        dist = np.zeros((len(queries), max_results), np.float64)
        cnt = np.zeros(len(queries), np.int32)
        f(abi.ptr(cloud), len(cloud), abi.ptr(queries), len(queries), radius2, max_results, 1 if sorted_ else 0, abi.ptr(idx),
          abi.ptr(dist), abi.ptr(cnt))
        return idx, dist, cnt

    def __init__(self):
        super().__init__(REF_LIB)

    def image_set_yuv420(self, y, u, v, shift):
        """PCCImage<uint16_t,3>::set on int16 decoder planes (dense strides): returns (Y, U, V) uint16"""
        f = self.lib.ref_image_set_yuv420
        f.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p]
        y, u, v = (np.ascontiguousarray(a, np.int16) for a in (y, u, v))
        H, W = y.shape
        out = np.zeros(H * W + 2 * (H // 2) * (W // 2), np.uint16)
        f(abi.ptr(y), abi.ptr(u), abi.ptr(v), W, H, shift, abi.ptr(out))
        q = (H // 2) * (W // 2)
        return out[:H * W].reshape(H, W), out[H * W:H * W + q].reshape(H // 2, W // 2), out[H * W + q:].reshape(H // 2, W // 2)

    def convert_bitdepth(self, plane, bits_in, bits_out, msb_align):
        """PCCImage<T,3>::convertBitdepth on a copy of `plane` (uint8 or uint16, 2-D)"""
        a = np.ascontiguousarray(plane).copy()
        f = self.lib.ref_convert_bitdepth_u16 if a.dtype == np.uint16 else self.lib.ref_convert_bitdepth_u8
        f.argtypes = [C.c_void_p] + [C.c_int] * 5
        H, W = a.shape
        f(abi.ptr(a), W, H, bits_in, bits_out, 1 if msb_align else 0)
        return a

    def yuv420_to_yuv444(self, y, u, v, bitdepth, filt):
        """the reference's own PCCInternalColorConverter<uint16_t>::convert("YUV420ToYUV444_<bits>_<filter>")"""
        f = getattr(self.lib, "ref_yuv420_to_yuv444", None)
        if f is None:
            raise RuntimeError("oracle/_ref/librabbit_ref.so predates the colour converter: cd oracle && make ref")
        f.argtypes = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p]
        y, u, v = (np.ascontiguousarray(a, np.uint16) for a in (y, u, v))
        out = np.zeros((3,) + y.shape, np.uint16)
        if f(abi.ptr(y), abi.ptr(u), abi.ptr(v), y.shape[1], y.shape[0], bitdepth, filt, abi.ptr(out)) != 0:
            raise RuntimeError("ref_yuv420_to_yuv444 failed")
        return out


class DropIn(_Backend):
    """the reference's harness and objects with the hot member functions replaced by the C++ shim
    (rabbit-transcoding_b200/host/PCCCodecB200.cpp -> librabbit_b200.so): the drop-in boundary under test."""
    prefix = "ref_"

    def __init__(self):
        super().__init__(DROPIN_LIB)
        self.lib.rb200_shim_stats.argtypes = [C.POINTER(abi.LaunchStats), C.c_int]

    def stats(self, reset=True):
        """kernel launches / bytes moved by the shim's CUDA context since the last reset"""
        s = abi.LaunchStats()
        assert self.lib.rb200_shim_stats(C.byref(s), 1 if reset else 0) == 0
        return s


class Port(_Backend):
    prefix = "orc_"

    def __init__(self):
        super().__init__(PORT_LIB)


def default_metrics_params(resolution=1023.0, c2p=True):
    """PCCMetricsParameters defaults (PccLibMetrics/source/PCCMetricsParameters.cpp)"""
    m = abi.MetricsParams()
    m.compute_c2c = 1
    m.compute_c2p = 1 if c2p else 0
    m.compute_color = 1
    m.compute_hausdorff = 0
    m.drop_duplicates = 2
    m.neighbors_proc = 1
    m.resolution = resolution
    return m
