"""N > 1 path on real GPUs: two ranks shard the frames of a GOF round-robin, each runs the metrics of its frames on its own
CUDA context, the per-frame accumulators are all-gathered (NCCL when the box has two GPUs, gloo when both ranks share the
one GPU of the box) and every rank must end with the table a single rank computes for all frames."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
KW = dict(n_frames=4, bitdepth=8, width=256, scale=0.9, seed=91, transfer_filter=1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _row(frame, r):
    return [frame, r.qf.c2c_psnr, r.qf.c2p_psnr, r.qf.color_psnr[0], r.qf.color_psnr[1], r.qf.color_psnr[2], r.q1.c2c_mse,
            r.q2.c2c_mse]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import rabbit_transcoding_b200 as rb
    ngpu = torch.cuda.device_count()
    dev = rank % ngpu
    torch.cuda.set_device(dev)
    backend = "nccl" if ngpu >= world else "gloo"
    dist.init_process_group(backend, rank=rank, world_size=world)
    g = rb.synthetic.generate_gof(**KW)
    mine = rb.dist.shard_frames(g.n_frames, world, rank)
    codec = rb.codec.PCCCodecB200(device=dev)
    # the rank decodes only the frames it owns
    local = {}
    met = rb.metrics.PCCMetricsB200(codec)
    met.setParameters(rb.metrics.default_parameters(resolution=255.0))
    for f in mine:
        sub = rb.synthetic.slice_gof(g, f, f + 1)
        codec.uploadGof(sub)
        codec.decodeGof()
        local[f] = met.compute([sub.sources[0]], [None], [sub.sources[0]])[0]
    st = codec.stats()
    assert st.kernel_launches > 50  # the CUDA library did the work on this rank
    frames, mean = rb.dist.gather_metrics(local, g.n_frames, 255.0, device="cuda" if backend == "nccl" else None)
    np.save(os.path.join(out_dir, f"r{rank}.npy"),
            np.array([[f["frame"], f["qf"]["c2c_psnr"], f["qf"]["c2p_psnr"]] + list(f["qf"]["color_psnr"]) +
                      [f["q1"]["c2c_mse"], f["q2"]["c2c_mse"]] for f in frames], np.float64))
    codec.close()
    dist.destroy_process_group()


def test_two_ranks_sharded_metrics_equal_single_rank(rb, codec, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = rb.synthetic.generate_gof(**KW)
    codec.uploadGof(g)
    codec.decodeGof()
    met = rb.metrics.PCCMetricsB200(codec)
    met.setParameters(rb.metrics.default_parameters(resolution=255.0))
    res = met.compute(g.sources, [None] * g.n_frames, g.sources)
    want = np.array([_row(f, r) for f, r in enumerate(res)], np.float64)
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), f"r{rank}.npy"))
        assert got.shape == want.shape
        assert np.array_equal(got[:, 0], want[:, 0])
        assert np.abs(got[:, 1:6] - want[:, 1:6]).max() <= 1e-6, "PSNR recomputed after the gather differs from the library's"
        assert np.array_equal(got[:, 6:].astype(np.float32), want[:, 6:].astype(np.float32))
