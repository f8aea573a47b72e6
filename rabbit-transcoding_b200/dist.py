"""Frame / GOF sharding over the GPUs of one box and the one exchange step of the path.

After video decoding every frame is an independent unit (inter-patch prediction is resolved earlier,
PccLibDecoder/source/PCCDecoder.cpp:995-1061), so the path shards with NO data-path collective: rank r owns the
frames `shard_frames(n, world, r)`, runs reconstruction / smoothing / metrics on them with its own context, and only
the per-frame metric results are exchanged: an all-gather of fixed-size records {frame, sse, num, max} in double,
after which rank-independent PSNRs are recomputed exactly as QualityMetrics::compute does
(PccLibMetrics/source/PCCMetrics.cpp:204-226) and sequence totals are formed in frame order on every rank.

Works with any torch.distributed backend: NCCL over NVLink on the B200 box, gloo in the CPU tests.
"""
import math

import numpy as np

RECORD = 24  # doubles per (frame, direction pair): see pack_result


def shard_frames(n_frames, world, rank):
    """round-robin frame ownership: frame f belongs to rank f % world"""
    return list(range(rank, n_frames, world))


def shard_streams(n_streams, world, rank):
    """whole-stream ownership for the multi-stream configuration (config 5)"""
    return list(range(rank, n_streams, world))


def pack_result(frame, r):
    """abi.MetricsResult -> RECORD doubles (the accumulators, not the derived floats)"""
    out = [float(frame)]
    for q in (r.q1, r.q2):
        out += [q.sse_c2c, q.sse_c2p, q.sse_color[0], q.sse_color[1], q.sse_color[2], q.max_c2c, q.max_c2p, float(q.num)]
    out += [float(r.source_points), float(r.source_after_dedup), float(r.rec_points), float(r.rec_after_dedup),
            float(r.tie_overflow)]
    out += [0.0] * (RECORD - len(out))
    return out


_libm = None


def _log10f(x):
    """glibc's log10f — the call the reference makes (numpy's float32 log10 is a different implementation)"""
    global _libm
    if _libm is None:
        import ctypes
        import ctypes.util
        _libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        _libm.log10f.restype = ctypes.c_float
        _libm.log10f.argtypes = [ctypes.c_float]
    return np.float32(_libm.log10f(float(x)))


def psnr(dist, peak, factor=1.0):
    """getPSNR (PCCMetrics.cpp:44-48): float max_energy = p * p; 10 * log10f( factor * max_energy / dist )"""
    dist, peak, factor = np.float32(dist), np.float32(peak), np.float32(factor)
    with np.errstate(divide="ignore"):
        ratio = np.float32(factor * np.float32(peak * peak)) / dist
    return np.float32(10) * _log10f(ratio)


def quality_from_sums(sse_c2c, sse_c2p, sse_col, num, resolution):
    """the float results QualityMetrics::compute derives from its double accumulators (:204-226)"""
    c2c = np.float32(sse_c2c / num)
    c2p = np.float32(sse_c2p / num)
    col = [np.float32(s / num) for s in sse_col]
    return dict(c2c_mse=c2c, c2c_psnr=psnr(c2c, resolution, 3), c2p_mse=c2p, c2p_psnr=psnr(c2p, resolution, 3),
                color_mse=col, color_psnr=[psnr(c, 1.0) for c in col])


def symmetric(q1, q2):
    """QualityMetrics::operator+ (:299-332): max of the MSEs, min of the PSNRs"""
    return dict(c2c_mse=max(q1["c2c_mse"], q2["c2c_mse"]), c2c_psnr=min(q1["c2c_psnr"], q2["c2c_psnr"]),
                c2p_mse=max(q1["c2p_mse"], q2["c2p_mse"]), c2p_psnr=min(q1["c2p_psnr"], q2["c2p_psnr"]),
                color_mse=[max(a, b) for a, b in zip(q1["color_mse"], q2["color_mse"])],
                color_psnr=[min(a, b) for a, b in zip(q1["color_psnr"], q2["color_psnr"])])


def all_gather_records_begin(local, n_frames, group=None, device=None):
    """Issues the collective and returns a handle for all_gather_records_end: {frame: abi.MetricsResult} of this rank ->
    one host-to-device copy of its records and one (asynchronous) all-gather.  A caller that pipelines GOFs finishes the
    handle one step later, so no rank ever waits for the slowest one inside its frame loop."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    per_rank = math.ceil(n_frames / world)
    host = np.full((per_rank, RECORD), -1.0, np.float64)
    for i, (f, r) in enumerate(sorted(local.items())):
        host[i] = pack_result(f, r)
    if world == 1:
        return dict(table=host)
    buf = torch.from_numpy(host).to(device) if device is not None else torch.from_numpy(host)
    out = torch.empty((world * per_rank, RECORD), dtype=torch.float64, device=buf.device)
    work = dist.all_gather_into_tensor(out, buf, group=group, async_op=True)
    return dict(work=work, out=out, buf=buf)


def all_gather_records_end(handle):
    """-> the [n_frames, RECORD] table of the whole sequence in frame order, the same on every rank"""
    if "table" in handle:
        table = handle["table"]
    else:
        handle["work"].wait()
        table = handle["out"].cpu().numpy()
    table = table[table[:, 0] >= 0]
    return table[np.argsort(table[:, 0], kind="stable")]


def all_gather_records(local, n_frames, group=None, device=None):
    """The collective itself, blocking: one copy of this rank's records up, one all-gather, one copy back."""
    return all_gather_records_end(all_gather_records_begin(local, n_frames, group=group, device=device))


def derive_table(table, resolution):
    """per-frame float results from the gathered accumulators, exactly as QualityMetrics::compute (:204-226) and
    operator+ (:299-332) derive them, and the sequence means"""
    frames = []
    for row in table:
        q = []
        for k in range(2):
            a = row[1 + 8 * k: 9 + 8 * k]
            q.append(quality_from_sums(a[0], a[1], a[2:5], a[7], resolution))
        frames.append(dict(frame=int(row[0]), q1=q[0], q2=q[1], qf=symmetric(q[0], q[1]),
                           source_points=int(row[17]), rec_points=int(row[19]), tie_overflow=int(row[21])))
    mean = dict(d1_psnr=float(np.mean([f["qf"]["c2c_psnr"] for f in frames])) if frames else float("nan"),
                d2_psnr=float(np.mean([f["qf"]["c2p_psnr"] for f in frames])) if frames else float("nan"),
                y_psnr=float(np.mean([f["qf"]["color_psnr"][0] for f in frames])) if frames else float("nan"))
    return frames, mean


def gather_metrics(local, n_frames, resolution, group=None, device=None):
    """local: {frame: abi.MetricsResult} of the frames this rank owns.  Returns the per-frame table of the whole
    sequence (same on every rank, frame order) and the sequence means — the only collective of the path."""
    return derive_table(all_gather_records(local, n_frames, group=group, device=device), resolution)
