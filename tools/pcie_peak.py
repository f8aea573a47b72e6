"""The box's host <-> device copy ceiling with every rank copying at once: pinned H2D and D2H on two streams per GPU,
concurrently on all ranks (launch under torchrun for N > 1).  Prints one JSON line per run (rank 0)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = 256
h_in = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
d_out = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=12):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return reps * (MB << 20) / float(t.item()) / 1e9  # GB/s per rank and direction (slowest rank)


run(True, True, 2)
res = {"n_gpus": world, "buffer_mib": MB,
       "h2d_only_gbs_per_gpu": round(run(True, False), 1), "d2h_only_gbs_per_gpu": round(run(False, True), 1)}
both = run(True, True)
res["both_gbs_per_gpu_per_direction"] = round(both, 1)
res["aggregate_both_directions_gbs"] = round(2 * both * world, 1)
res["aggregate_h2d_only_gbs"] = round(res["h2d_only_gbs_per_gpu"] * world, 1)
res["host_cpus"] = os.cpu_count()
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
