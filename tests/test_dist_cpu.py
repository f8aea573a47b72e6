"""N > 1 path on the CPU: world-size-2 gloo run of the frame sharding and of the one exchange step (all-gather of the
per-frame metric accumulators + rank-independent PSNR recomputation).  No GPU, no CUDA library calls."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_result(rb, frame):
    """a deterministic MetricsResult as rb200_metrics would fill it for `frame`"""
    rng = np.random.default_rng(100 + frame)
    r = rb.abi.MetricsResult()
    for q in (r.q1, r.q2):
        q.num = int(rng.integers(700000, 900000))
        q.sse_c2c = float(rng.integers(10000, 200000))
        q.sse_c2p = float(rng.random() * 50000)
        for k in range(3):
            q.sse_color[k] = float(rng.random() * 300)
        q.max_c2c = 18.0
        q.max_c2p = 7.5
    r.source_points, r.source_after_dedup, r.rec_points, r.rec_after_dedup = 780000, 780000, 860000, 840000
    return r


def _worker(rank, world, port, n_frames, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import rabbit_transcoding_b200 as rb
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = rb.dist.shard_frames(n_frames, world, rank)
    local = {f: _fake_result(rb, f) for f in mine}
    frames, mean = rb.dist.gather_metrics(local, n_frames, 1023.0)
    np.save(os.path.join(out_dir, f"r{rank}.npy"),
            np.array([[f["frame"], f["qf"]["c2c_psnr"], f["qf"]["c2p_psnr"], f["qf"]["color_psnr"][0]] for f in frames]))
    np.save(os.path.join(out_dir, f"m{rank}.npy"), np.array([mean["d1_psnr"], mean["d2_psnr"], mean["y_psnr"]]))
    dist.destroy_process_group()


def _worker_pipelined(rank, world, port, n_frames, out_dir):
    """two steps in flight: the collective of step k is finished after the one of step k + 1 was issued (bench.py's loop)"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import rabbit_transcoding_b200 as rb
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = rb.dist.shard_frames(n_frames, world, rank)
    tables, pending = [], None
    for step in range(3):
        h = rb.dist.all_gather_records_begin({f: _fake_result(rb, f + 100 * step) for f in mine}, n_frames)
        if pending is not None:
            tables.append(rb.dist.all_gather_records_end(pending))
        pending = h
    tables.append(rb.dist.all_gather_records_end(pending))
    np.save(os.path.join(out_dir, f"p{rank}.npy"), np.stack(tables))
    dist.destroy_process_group()


def test_pipelined_all_gather_world2_gloo(rb, tmp_path):
    n_frames, world = 6, 2
    mp.spawn(_worker_pipelined, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "p0.npy"), np.load(tmp_path / "p1.npy")
    assert a.shape == (3, n_frames, rb.dist.RECORD) and np.array_equal(a, b)
    for step in range(3):
        assert a[step][:, 0].tolist() == list(range(n_frames))
        for f in range(n_frames):
            assert a[step][f].tolist() == rb.dist.pack_result(f, _fake_result(rb, f + 100 * step))


def test_shard_frames_covers_every_frame_once(rb):
    for n in (1, 5, 32, 300):
        for world in (1, 2, 4, 8):
            owned = [rb.dist.shard_frames(n, world, r) for r in range(world)]
            flat = sorted(f for o in owned for f in o)
            assert flat == list(range(n))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def test_gather_metrics_world2_gloo(rb, tmp_path):
    n_frames, world = 7, 2
    mp.spawn(_worker, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(a, b), "every rank must end up with the same per-frame table"
    assert a[:, 0].tolist() == list(range(n_frames))
    assert np.array_equal(np.load(tmp_path / "m0.npy"), np.load(tmp_path / "m1.npy"))
    # the gathered table equals the single-process computation, frame by frame
    for f in range(n_frames):
        r = _fake_result(rb, f)
        q1 = rb.dist.quality_from_sums(r.q1.sse_c2c, r.q1.sse_c2p, list(r.q1.sse_color), r.q1.num, 1023.0)
        q2 = rb.dist.quality_from_sums(r.q2.sse_c2c, r.q2.sse_c2p, list(r.q2.sse_color), r.q2.num, 1023.0)
        qf = rb.dist.symmetric(q1, q2)
        assert a[f, 1] == qf["c2c_psnr"] and a[f, 2] == qf["c2p_psnr"] and a[f, 3] == qf["color_psnr"][0]


def test_psnr_matches_reference_formula(rb):
    # getPSNR( dist, p, factor ) = 10 * log10f( factor * p * p / dist ), all float (PCCMetrics.cpp:44-48)
    v = rb.dist.psnr(np.float32(0.163), 1023.0, 3)
    assert abs(float(v) - 10 * np.log10(3 * 1023.0 ** 2 / 0.163)) < 1e-4
    assert np.isinf(rb.dist.psnr(0.0, 1023.0, 3))


def test_c_abi_pack_unpack_equals_python_gather(rb):
    """rb200_metrics_pack / rb200_metrics_unpack (the C form of the exchange record, for a C++ host with its own
    communicator) derive the same floats from the accumulators as rabbit_transcoding_b200.dist does"""
    import ctypes as C
    lib = rb.abi.load_library()
    mp = rb.metrics.default_parameters(resolution=1023.0)
    for frame in (0, 3, 17):
        r = _fake_result(rb, frame)
        rec = (C.c_double * 24)()
        assert lib.rb200_metrics_pack(frame, C.byref(r), rec) == 0
        assert list(rec)[:len(rb.dist.pack_result(frame, r))] == rb.dist.pack_result(frame, r)  # the same record layout
        out, fr = rb.abi.MetricsResult(), C.c_int(-1)
        assert lib.rb200_metrics_unpack(rec, C.byref(mp), C.byref(fr), C.byref(out)) == 0
        assert fr.value == frame and out.q1.num == r.q1.num and out.rec_after_dedup == r.rec_after_dedup
        for q_c, q_in in ((out.q1, r.q1), (out.q2, r.q2)):
            want = rb.dist.quality_from_sums(q_in.sse_c2c, q_in.sse_c2p, list(q_in.sse_color), q_in.num, 1023.0)
            assert np.float32(q_c.c2c_mse) == want["c2c_mse"] and np.float32(q_c.c2c_psnr) == want["c2c_psnr"]
            assert np.float32(q_c.c2p_mse) == want["c2p_mse"] and np.float32(q_c.c2p_psnr) == want["c2p_psnr"]
            for k in range(3):
                assert np.float32(q_c.color_psnr[k]) == want["color_psnr"][k]
        a = rb.dist.quality_from_sums(r.q1.sse_c2c, r.q1.sse_c2p, list(r.q1.sse_color), r.q1.num, 1023.0)
        b = rb.dist.quality_from_sums(r.q2.sse_c2c, r.q2.sse_c2p, list(r.q2.sse_color), r.q2.num, 1023.0)
        sym = rb.dist.symmetric(a, b)
        assert np.float32(out.qf.c2c_psnr) == sym["c2c_psnr"] and np.float32(out.qf.c2p_psnr) == sym["c2p_psnr"]
