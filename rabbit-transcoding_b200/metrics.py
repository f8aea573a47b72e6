"""Host-side mirror of PccLibMetrics' PCCMetrics (PccLibMetrics/include/PCCMetrics.h:94-116) over the C ABI.

    m = PCCMetricsB200(codec)                 # shares the codec's context (one per GPU)
    m.setParameters(params)                   # abi.MetricsParams == the PCCMetricsParameters fields the path reads
    m.compute(sources, reconstructs, normals) # PCCMetrics::compute( sources, reconstructs, normals ), PCCMetrics.cpp:334-385
    m.display()                               # same lines as PCCMetrics::display (:387-404)

`sources[i]` / `normals[i]` are dicts with `positions` (int16 [n,3]), `colors` (uint8 [n,3]) and, for the normal
cloud, `normals` (float32 [n,3]); `reconstructs[i]` is such a dict or None for "frame i of the GOF resident in the
codec's context" (no host round trip).  As in the reference, results accumulate across calls (quality1_, quality2_,
qualityF_, sourcePoints_ ...).  Errors raise RabbitError; there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import abi
from .codec import RabbitError


def default_parameters(resolution=1023.0, compute_c2p=True):
    """PCCMetricsParameters defaults (PccLibMetrics/source/PCCMetricsParameters.cpp:38-58)"""
    m = abi.MetricsParams()
    m.compute_c2c = 1
    m.compute_c2p = 1 if compute_c2p else 0
    m.compute_color = 1
    m.compute_hausdorff = 0
    m.drop_duplicates = 2
    m.neighbors_proc = 1
    m.resolution = resolution
    return m


def _is_device(a):
    return hasattr(a, "data_ptr")  # a torch tensor (host or CUDA): the ABI resolves the pointer itself (UVA)


def _view(cloud, keep, with_normals=False):
    v = abi.CloudView()
    if cloud is None:
        return v
    if _is_device(cloud["positions"]):
        # clouds cached on the device (or in pinned memory) as torch tensors: int16 [n,3], uint8 [n,3], float32 [n,3]
        pos = cloud["positions"]
        assert pos.is_contiguous() and pos.element_size() == 2 and pos.shape[1] == 3
        keep.append(pos)
        v.positions, v.count = pos.data_ptr(), pos.shape[0]
        col = cloud.get("colors")
        if col is not None:
            assert col.is_contiguous() and col.element_size() == 1
            keep.append(col)
            v.colors = col.data_ptr()
        nrm = cloud.get("normals") if with_normals else None
        if nrm is not None:
            assert nrm.is_contiguous() and nrm.element_size() == 4
            keep.append(nrm)
            v.normals = nrm.data_ptr()
        return v
    pos = np.ascontiguousarray(cloud["positions"], np.int16)
    keep.append(pos)
    v.positions = abi.ptr(pos)
    v.count = pos.shape[0]
    col = cloud.get("colors")
    if col is not None:
        col = np.ascontiguousarray(col, np.uint8)
        keep.append(col)
        v.colors = abi.ptr(col)
    if with_normals and cloud.get("normals") is not None:
        nrm = np.ascontiguousarray(cloud["normals"], np.float32)
        keep.append(nrm)
        v.normals = abi.ptr(nrm)
    return v


class PCCMetricsB200:
    def __init__(self, codec):
        self._codec = codec
        self._lib = codec._lib
        self.params_ = default_parameters()
        self.quality1_, self.quality2_, self.qualityF_ = [], [], []
        self.sourcePoints_, self.sourceDuplicates_ = [], []
        self.reconstructPoints_, self.reconstructDuplicates_ = [], []
        self.results_ = []

    def setParameters(self, params):
        self.params_ = params

    def cacheSources(self, on=True):
        """rb200_metrics_cache_sources: device-resident sources keep their index across compute() calls"""
        st = self._lib.rb200_metrics_cache_sources(self._codec._h, 1 if on else 0)
        if st != abi.RB200_OK:
            raise RabbitError(st, "rb200_metrics_cache_sources")

    def compute(self, sources, reconstructs, normals=None):
        if len(sources) != len(reconstructs):
            # PCCMetrics.cpp:341-347 prints and exits(-1)
            raise RabbitError(abi.RB200_ERR_INVALID, "group of frames must have same numbers of frames "
                              f"( src = {len(sources)} rec = {len(reconstructs)} )")
        params = abi.MetricsParams.from_buffer_copy(bytes(self.params_))
        if normals is not None and len(normals) != 0 and len(normals) != len(sources):
            params.compute_c2p = 0  # :337-339
            normals = None
        n = len(sources)
        keep = []
        vs = (abi.CloudView * n)()
        vr = (abi.CloudView * n)()
        for i in range(n):
            src = dict(sources[i])
            if normals:
                # the normal cloud carries the same points as the source (PCCPointSet3::copyNormals looks them up
                # by position); the ABI takes positions + normals of that cloud through the source view
                nc = normals[i]
                same = nc["positions"] is src["positions"] or (not _is_device(nc["positions"]) and (
                    nc["positions"].shape[0] == src["positions"].shape[0] and np.array_equal(nc["positions"], src["positions"])))
                if not same and _is_device(nc["positions"]):
                    raise RabbitError(abi.RB200_ERR_INVALID, "device-resident normal clouds must be the source cloud itself")
                if not same:
                    src = self._attach_normals(src, nc)
                else:
                    src["normals"] = nc["normals"]
            else:
                src.pop("normals", None)
            vs[i] = _view(src, keep, with_normals=True)
            vr[i] = _view(reconstructs[i], keep)
        out = (abi.MetricsResult * n)()
        st = self._lib.rb200_metrics(self._codec._h, C.byref(params), n, vs, vr, out)
        if st not in (abi.RB200_OK, abi.RB200_ERR_TIE_OVERFLOW):
            raise RabbitError(st, self._lib.rb200_error_string(self._codec._h).decode())
        for i in range(n):
            r = abi.MetricsResult.from_buffer_copy(bytes(out[i]))
            self.results_.append(r)
            self.quality1_.append(r.q1)
            self.quality2_.append(r.q2)
            self.qualityF_.append(r.qf)
            self.sourcePoints_.append(r.source_points)
            self.sourceDuplicates_.append(r.source_after_dedup)
            self.reconstructPoints_.append(r.rec_points)
            self.reconstructDuplicates_.append(r.rec_after_dedup)
        return [self.results_[-n + i] for i in range(n)]

    @staticmethod
    def _attach_normals(src, nc):
        """normal cloud in a different order than the source: reorder by position (copyNormals' lookup)"""
        def key(p):
            p = p.astype(np.int64) + 32768
            return (p[:, 0] << 32) | (p[:, 1] << 16) | p[:, 2]
        ks, kn = key(src["positions"]), key(nc["positions"])
        order = np.argsort(kn, kind="stable")
        pos = np.searchsorted(kn[order], ks)
        pos = np.clip(pos, 0, len(kn) - 1)
        if len(kn) != len(np.unique(ks)) or not np.array_equal(kn[order][pos], ks):
            raise RabbitError(abi.RB200_ERR_INVALID, "a source point is not present in the normal point cloud")
        out = dict(src)
        out["normals"] = np.ascontiguousarray(nc["normals"][order][pos], np.float32)
        return out

    def removeDuplicate(self, cloud, drop_duplicates=2):
        """PCCPointSet3::removeDuplicate( out, dropDuplicates ), PCCPointSet.cpp:169-218"""
        keep = []
        v = _view(cloud, keep)
        n = v.count
        pos = np.zeros((n, 3), np.int16)
        col = np.zeros((n, 3), np.uint8) if cloud.get("colors") is not None else None
        cnt = abi.i64(0)
        st = self._lib.rb200_remove_duplicates(self._codec._h, C.byref(v), drop_duplicates, abi.ptr(pos), abi.ptr(col),
                                               C.byref(cnt))
        if st != abi.RB200_OK:
            raise RabbitError(st, self._lib.rb200_error_string(self._codec._h).decode())
        m = cnt.value
        return dict(positions=pos[:m].copy(), colors=(col[:m].copy() if col is not None else None))

    def display(self):
        print("Metrics results ")
        for i, r in enumerate(self.results_):
            print(f"WARNING: {self.reconstructPoints_[i] - self.reconstructDuplicates_[i]} points with same coordinates found")
            print(f"Imported intrinsic resoluiton: {self.params_.resolution:g}")
            print(f"Peak distance for PSNR: {self.params_.resolution:g}")
            print("Point cloud sizes for org version, dec version, and the scaling ratio: "
                  f"{self.sourcePoints_[i]}, {self.reconstructDuplicates_[i]}, "
                  f"{np.float32(self.reconstructDuplicates_[i]) / np.float32(self.sourcePoints_[i]):g}")
            for tag, q in (("1", r.q1), ("2", r.q2), ("F", r.qf)):
                p = self.params_
                if p.compute_c2c:
                    print(f"   mse{tag}      (p2point): {q.c2c_mse:g}")
                    print(f"   mse{tag},PSNR (p2point): {q.c2c_psnr:g}")
                if p.compute_c2p:
                    print(f"   mse{tag}      (p2plane): {q.c2p_mse:g}")
                    print(f"   mse{tag},PSNR (p2plane): {q.c2p_psnr:g}")
                if p.compute_color:
                    for k in range(3):
                        print(f"   c[{k}],    {tag}         : {q.color_mse[k]:g}")
                    for k in range(3):
                        print(f"   c[{k}],PSNR{tag}         : {q.color_psnr[k]:g}")
