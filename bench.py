#!/usr/bin/env python
"""bench.py — the hot path of RABBIT's V-PCC transcode loop on B200 ("Measurement").

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU path (oracle/_ref), rank 0 only

A *step* is one pass of the post-video-decode path over one synthetic 32-frame vox10 GOF per GPU
(BASELINE.json configs[1]): reconstruction (occupancy -> block-to-patch -> reprojection of near/far layers ->
boundary types -> attribute fetch), grid geometry smoothing, attribute re-transfer for the moved points,
colour smoothing, YUV16 -> RGB8, and (when --metrics) the D1/D2/colour metrics against the source clouds.

`value`  : whole-job Mpts/s of the decoder's Rec-1 sequence (PCCDecoder.cpp:330-508: reconstruction, grid geometry
           smoothing, transferColors16bitBP, colour smoothing, RGB8) with the decoded frames already resident in HBM
           (rb200_decode_gof only).  Every frame's ordered MD5 is compared with the unmodified reference's before the line is
           printed (`parity_checked_frames`); a mismatch ends the run without a value.
`e2e`    : the transcode loop of BASELINE.json configs[2] through the reference-facing calls with HOST buffers, inside the
           timer every step: pinned H2D of the decoder-native planes (8-bit 4:2:0 attribute frames + 8-bit geometry luma +
           occupancy) and the patch tables, the decoder's 4:2:0 -> 4:4:4 16-bit conversion on the GPU, the Rec-1 decode,
           D1 + D2 + colour metrics of every frame against its source cloud ON THE RESIDENT RECONSTRUCTION, D2H of the
           metric records (and their all-gather at N > 1).  Source clouds are cached in HBM (the same source is measured
           against every rate point of a transcode); `e2e.with_source_upload` moves them from pinned host memory every
           step, `e2e.decode_and_download` is the decoder alone with positions + RGB8 of every frame copied back.
`roofline`: the dominant kernel of the step (per-kernel CUDA events on the launching stream), algorithmic
           bytes per launch / its average duration, against MEASURED_PEAKS.json; `traffic` = DRAM bytes
           of that kernel from the committed ncu capture (profiles/ncu_traffic.json).
`cpu_baseline`: the unmodified reference (oracle/_ref/librabbit_ref.so) on this box's host cores, same GOF.
`metrics` : D1 + D2 + colour of every frame against its source cloud (frames/s), sources from pinned host memory.
`full_decoder`: the decoder's whole Rec-1 sequence including transferColors16bitBP.

Nothing here reads /root/reference.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reconstructed_points_throughput"
UNIT = "Mpts/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--frames", type=int, default=32, help="frames per GOF per GPU")
    ap.add_argument("--workload", default="vox10", choices=("vox10", "vox11", "vox11_eom", "streams", "tiny"),
                    help="vox10: BASELINE configs[1]/[2] (default); vox11 / vox11_eom: configs[3] (use --frames 38 on 8 GPUs "
                         "for the 300-frame sequence); streams: configs[4], 8 vox10 streams at rates r1-r5")
    ap.add_argument("--distinct-frames", type=int, default=0,
                    help="generate only this many distinct frames and repeat them up to --frames (0 = all distinct)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-metrics", action="store_true", help="skip the D1/D2 metrics leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the MD5 check against the reference (debugging only)")
    ap.add_argument("--quick", action="store_true", help="only value, parity, e2e (transcode loop) and the roofline legs")
    ap.add_argument("--ref-frames", type=int, default=0, help="--impl reference: frames per step (0 = auto)")
    return ap.parse_args()


WORKLOADS = {
    # name: generator arguments (synthetic.generate_gof) — vox10: ~0.84 M points / frame, atlas 1280x1280
    # max_depth: patches stay inside the range of the 8-bit geometry video (geometryNominal2dBitdepth 8, maxAllowedDepth)
    "vox10": dict(bitdepth=10, width=1280, scale=0.68, height_blocks=80, max_depth=249),
    "vox11": dict(bitdepth=11, width=2560, scale=0.62, height_blocks=176, max_depth=249),
    # the lossless-style variant of configs[3]: occupancy precision 1, enhanced occupancy map code, one EOM patch per
    # frame, geometry / colour smoothing off as in cfg/common/ctc-common-lossless-geometry-attribute.cfg
    "vox11_eom": dict(bitdepth=11, width=2560, scale=0.62, height_blocks=208, max_depth=249, eom=True,
                      geometry_smoothing=False, color_smoothing=False),
    "tiny": dict(bitdepth=8, width=256, scale=0.9, height_blocks=32),
}


def make_gof(rb, args, rank, world, frames=None, extra=None):
    kw = dict(WORKLOADS[args.workload])
    # the decoder-faithful Rec-1 sequence: every profile that smooths the geometry also re-transfers the attributes
    # (PCCDecoder.cpp:434-465); BASELINE.json configs[1] without the re-transfer is reported as config.without_retransfer
    kw.update(seed=0x0AB817 + 1000 * rank, transfer_filter=1)  # Rec-1: attrTransferFilterType_ = 1 (PCCDecoderParameters.cpp:125-134)
    if kw.get("eom"):
        kw["transfer_filter"] = 0
    kw.update(extra or {})
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = frames or args.frames
    distinct = min(n, args.distinct_frames) if getattr(args, "distinct_frames", 0) else n
    workers = max(1, min(distinct, min(ncpu, (os.cpu_count() or 1) // max(1, world))))
    g = rb.synthetic.generate_gof_parallel(distinct, workers=workers, **kw)
    if distinct < n:  # the distinct frames repeated cyclically (throughput runs of long sequences)
        g = rb.synthetic.concat_gofs([rb.synthetic.slice_gof(g, f % distinct, f % distinct + 1) for f in range(n)])
    return g


class ClockSampler:
    """samples SM clock + throttle reasons of one GPU every 100 ms during the timed regions (NVML)"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        self._on = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            if self._on.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.1)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def on(self):
        self._on.set()

    def off(self):
        self._on.clear()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=1)

    def summary(self):
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def algorithmic_bytes(name, g, n_points, n_type1, n_moved):
    """SURVEY.md §8d algorithmic bytes of ONE launch of kernel `name` over the whole GOF (F frames)."""
    p = g.params
    F, W, H, M, pr = g.n_frames, p.width, p.height, p.map_count_minus1 + 1, p.occupancy_precision
    geo = F * W * H * 2 * M
    occ = F * (W // pr) * (H // pr)
    N = n_points
    table = {
        # one pass over the geometry luma + occupancy; counts only
        "reproject_count": geo + occ,
        # B_rep: geometry + occupancy + attribute gather (6 B) + 26 B of outputs per point
        "reproject_emit": geo + occ + N * 6 + N * (6 + 6 + 2 + 4 + 8),
        "occupancy_bitmap": occ + F * H * ((W + 31) // 32) * 4,
        # B_geo split over its passes (pos+type 8 B, partition 4 B)
        # B_geo / B_col split over their passes (pos + type 8 B, partition 4 B, colour16 6 B, boundary index 4 B)
        "geo_accumulate": N * 12,
        "geo_filter": n_type1 * (4 + 8) + n_moved * 8,
        "col_mark": n_type1 * (4 + 8),
        # decoder-native ingest: 8-bit 4:2:0 in (1.5 B / pixel), 16-bit 4:4:4 out (6 B / pixel) per attribute frame
        "attribute_420_to_444": F * M * W * H // 2 + F * M * W * H * 4,
        "attribute_luma_to_16": F * M * W * H * 3,
        # the luma lists are written by the accumulation itself (2 B per point)
        "col_accumulate": N * (8 + 6 + 4) + N * 2,
        "col_filter": n_type1 * (4 + 8 + 6) + n_type1 * 6,
        "to_rgb8": N * (6 + 3),
        # kd forest of transferColors16bitBP: 2 trees per frame, E = 2 N element records of 8 bytes; per level launch
        "kd_init": 2 * N * (8 + 8),
        "kd_count": 2 * N * 8,
        "kd_stage1": 2 * N * (8 + 4),
        "kd_apply1": 2 * N * (4 + 4),
        "kd_stage2": 2 * N * 2,
        "kd_apply2": 2 * N * (8 + 2),
        "kd_subtree": 2 * N * (8 + 8),
    }
    return table.get(name)


def workload_string(args, p, n_frames):
    return (f"synthetic {args.workload} {n_frames}-frame GOF per GPU, C2RA r3 shape: atlas {p.width}x{p.height}, 2 maps, "
            f"p={p.occupancy_precision}; value = Rec-1 decode (reconstruction + grid geometry smoothing + "
            "transferColors16bitBP + colour smoothing + RGB8); e2e = transcode loop (decoder-native planes in, Rec-1 decode, "
            "D1 + D2 + colour metrics of every frame out)")


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import rabbit_transcoding_b200 as rb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    ncpu_all = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # host threads and pinned buffers of a rank stay on the NUMA node of its GPU
    try:
        import pynvml
        pynvml.nvmlInit()
        h_ = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h_, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        if cpus and not os.environ.get("RB200_BENCH_NO_AFFINITY"):
            os.sched_setaffinity(0, cpus)
            print(f"[bench rank {rank}] bound to {len(cpus)} cpus of GPU {local}'s NUMA node", file=sys.stderr, flush=True)
    except Exception as ex:
        print(f"[bench rank {rank}] no cpu affinity: {ex}", file=sys.stderr, flush=True)
    # the driver wants ONE JSON line on stdout: NCCL's version banner (and anything else a library prints there) is sent
    # to stderr by pointing fd 1 at fd 2 for the whole run; the JSON line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t0 = time.time()
    gof = make_gof(rb, args, rank, world)
    t_gen = time.time() - t0

    def pin(a):
        return torch.from_numpy(a).pin_memory()
    pinned = dict(occupancy=pin(gof.occupancy), geometry=pin(gof.geometry), attribute=pin(gof.attribute))
    gof.occupancy, gof.geometry, gof.attribute = (pinned[k].numpy() for k in ("occupancy", "geometry", "attribute"))

    codec = rb.codec.PCCCodecB200(device=local)
    stream = torch.cuda.current_stream()
    codec.setStream(stream.cuda_stream)
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    W = max(args.warmup, 3)

    def timed_resident(steps):
        for _ in range(W):
            codec.decodeGof()
        codec.stats(reset=True)
        barrier()
        torch.cuda.synchronize()
        clocks.on()
        e0.record(stream)
        for _ in range(steps):
            codec.decodeGof()
        e1.record(stream)
        torch.cuda.synchronize()
        clocks.off()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), codec.stats(reset=True).kernel_launches

    # ---------------- leg 1: inputs resident in HBM, the decoder's Rec-1 sequence ----------------
    codec.uploadGof(gof)
    ms_total, launches = timed_resident(args.steps)
    counts = codec.frameCounts()
    n_points = sum(c.total for c in counts)
    n_moved = sum(c.smoothed for c in counts)
    all_points = sum_over_ranks(n_points)
    value = all_points * args.steps / (ms_total * 1e-3) / 1e6
    print(f"[bench rank {rank}] resident Rec-1 leg {ms_total / args.steps:.4f} ms per step, {n_points} points", file=sys.stderr, flush=True)
    gpu_md5 = [codec.computeChecksum(f) for f in range(gof.n_frames)]  # PCCPointSet3::computeChecksum of every decoded frame

    # BASELINE.json configs[1] as worded ("reconstruction + geometry/colour smoothing"): the same without the re-transfer
    without_retransfer = None
    if not args.quick and gof.params.attr_transfer_filter_type == 1:
        gof.params.attr_transfer_filter_type = 0
        codec.uploadGof(gof)
        ms_nt, _ = timed_resident(max(3, min(args.steps, 10)))
        without_retransfer = {"value": round(all_points * max(3, min(args.steps, 10)) / (ms_nt * 1e-3) / 1e6, 2), "unit": UNIT,
                              "ms_per_step": round(ms_nt / max(3, min(args.steps, 10)), 4),
                              "what": "attr_transfer_filter_type = 0 (round-1 headline configuration)"}
        gof.params.attr_transfer_filter_type = 1

    # ---------------- parity of the benchmarked workload (outside every timed region) ----------------
    # N = 1: the reference run below is also the cpu_baseline (1 thread, the reference's real behaviour);
    # N > 1: every rank checks its own GOF frame-parallel on its share of the host cores
    cpu = None
    parity_frames = 0
    if not args.no_parity:
        from oracle import checker
        if not checker.have_reference():
            raise SystemExit("bench.py: oracle/_ref/librabbit_ref.so is missing: the benchmarked workload cannot be checked")
        chk = checker.Reference()
        threads = 1 if (world == 1 and not args.no_cpu_baseline) else max(1, ncpu_all // world)
        t0 = time.time()
        run = chk.run_gof(gof, keep=(), threads=threads)
        wall = time.time() - t0
        bad = [f for f in range(gof.n_frames) if run.md5(f) != gpu_md5[f]]
        if bad:
            raise SystemExit(f"bench.py: rank {rank}: frames {bad} differ from the reference (ordered MD5); no value is reported")
        parity_frames = gof.n_frames
        if world == 1 and not args.no_cpu_baseline:
            rec_ms = sum(run.time_ms(f, 0) for f in range(gof.n_frames))
            post_ms = sum(run.time_ms(f, 1) for f in range(gof.n_frames))
            cpu = {"value": round(n_points / wall / 1e6, 4), "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"all {gof.n_frames} frames of the same GOF, Rec-1 decode, {n_points} points, {wall:.1f} s wall "
                             f"(reconstruction {rec_ms / 1e3:.1f} s + post-processing {post_ms / 1e3:.1f} s of CPU time); the "
                             "ordered MD5 of every frame equals the CUDA path's",
                   "wall_s": round(wall, 2)}
        del run
    parity_frames = int(sum_over_ranks(parity_frames))

    # ---------------- leg 2: the transcode loop, end to end from host buffers ----------------
    e2e = None
    metrics_leg = None
    mp = rb.metrics.default_parameters(resolution=float((1 << gof.params.geometry_bitdepth_3d) - 1))
    if not args.no_e2e:
        native = rb.synthetic.to_decoder_planes(gof, bitdepth=8, filt=0)
        native["geometry"] = pin(native["geometry"]).numpy()
        native["attribute"] = pin(native["attribute"]).numpy()
        # source clouds: pinned host copies (uploaded every step in the `with_source_upload` variant) and copies cached in HBM
        src_host = [dict(positions=pin(s["positions"]), colors=pin(s["colors"]), normals=pin(s["normals"])) for s in gof.sources]
        src_dev = [{k: v.cuda(non_blocking=True) for k, v in s.items()} for s in src_host]
        torch.cuda.synchronize()
        src_bytes = sum(v.numel() * v.element_size() for s in src_host for v in s.values())
        NL = int(os.environ.get("RB200_BENCH_LANES", "2"))
        lanes = []
        for k in range(NL):
            cx = codec if k == 0 else rb.codec.PCCCodecB200(device=local)
            if k:
                sx = torch.cuda.Stream()
                cx.setStream(sx.cuda_stream)
                cx._stream_keep = sx
            mx = rb.metrics.PCCMetricsB200(cx)
            mx.setParameters(mp)
            lanes.append(dict(codec=cx, met=mx, res=None, out=None))
        resident = [None] * gof.n_frames
        errs = []
        gathered = {}

        def loop_step(L, sources, download, step):
            c_ = L["codec"]
            c_.uploadGofYuv420(gof, native)
            c_.decodeGof()
            if download:
                L["n_got"] = c_.getGof(fields=("positions", "colors"), out=L["out"])[1]
                return
            L["met"].results_.clear()
            res = L["met"].compute(sources, resident, sources)
            if world > 1:
                # the one exchange step of the path: all-gather of the per-frame accumulators (NCCL).  Lanes take turns
                # in step order, so every rank issues the collectives in the same sequence.
                # (the per-frame floats are derived from the gathered accumulators once, after the loop: the records are
                # the step's result)
                # the collective is issued here and finished one step later (or after the loop, still inside the timed
                # region), so a rank never waits for the slowest one inside its frame loop
                with turn:
                    turn.wait_for(lambda: order[0] >= step)
                    h = rb.dist.all_gather_records_begin({rank + world * i: r for i, r in enumerate(res)},
                                                         world * gof.n_frames, device="cuda")
                    order[0] += 1
                    turn.notify_all()
                prev, L["pending"] = L.get("pending"), h
                if prev is not None:
                    gathered["table"] = rb.dist.all_gather_records_end(prev)
            L["res"] = res

        turn = threading.Condition()
        order = [0]

        def worker(k, nsteps, sources, download, nl):
            try:
                torch.cuda.set_device(local)
                for i in range(nsteps):
                    loop_step(lanes[k], sources, download, k + i * nl)
                if lanes[k].get("pending") is not None:  # the last step's records
                    gathered["table"] = rb.dist.all_gather_records_end(lanes[k].pop("pending"))
            except Exception as ex:  # surfaced after the join
                errs.append(ex)
                with turn:
                    order[0] = 1 << 60
                    turn.notify_all()

        def run_lanes(nsteps, sources, download, nl):
            order[0] = 0
            ts = [threading.Thread(target=worker, args=(k, (nsteps + nl - 1 - k) // nl, sources, download, nl)) for k in range(nl)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            if errs:
                raise errs[0]
            return nl

        def time_loop(sources, download=False):
            run_lanes(2 * NL, sources, download, NL)  # warm-up: tables, scratch, pinned staging of every lane
            psteps = max(int(os.environ.get("RB200_BENCH_E2E_STEPS", "0")), args.steps, 4 * NL)
            psteps = -(-psteps // NL) * NL
            for L in lanes:
                L["codec"].stats(reset=True)
            barrier()
            torch.cuda.synchronize()
            clocks.on()
            e0.record(stream)
            nl = run_lanes(psteps, sources, download, NL)
            torch.cuda.synchronize()
            e1.record(stream)
            torch.cuda.synchronize()
            clocks.off()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1))
            sts = [L["codec"].stats(reset=True) for L in lanes]
            return ms, psteps, sts, nl

        # (a) the headline: sources cached in HBM.  rb200_metrics counts the device-to-device import of a cached cloud as
        # "h2d": what crosses PCIe is the planes and the tables, counted from the upload alone.
        lanes[0]["codec"].stats(reset=True)
        lanes[0]["codec"].uploadGofYuv420(gof, native)
        plane_bytes = lanes[0]["codec"].stats(reset=True).h2d_bytes
        for L in lanes:  # the same source frames every step (as against every rate point of transcode.sh): index kept
            L["met"].cacheSources(True)
        ms, psteps, sts, nl = time_loop(src_dev)
        for L in lanes:
            L["met"].cacheSources(False)
        res = lanes[0]["res"]
        result_bytes = len(res) * C_sizeof_result(rb)
        e2e = {"value": round(all_points * psteps / (ms * 1e-3) / 1e6, 2), "unit": UNIT,
               "h2d_bytes_per_step": int(plane_bytes), "d2h_bytes_per_step": int(result_bytes),
               "ms_per_step": round(ms / psteps, 3), "steps": psteps,
               "metric_frames_per_s": round(world * gof.n_frames * psteps / (ms * 1e-3), 1),
               "mode": ("transcode loop: uploadGofYuv420 (8-bit 4:2:0 attribute frames + 8-bit geometry luma + occupancy from "
                        "pinned host memory; 4:2:0 -> 4:4:4 16-bit conversion on the GPU) -> decodeGof (Rec-1) -> "
                        "PCCMetrics::compute (D1 + D2 + colour, both directions, duplicate removal) on the resident "
                        f"reconstruction against source clouds cached in HBM (their de-duplicated index kept across steps, "
                        f"rb200_metrics_cache_sources) -> metric records to the host; {nl} GOF(s) in "
                        "flight per GPU" + ("; all-gather of the records over NCCL every step" if world > 1 else "")),
               "d1_psnr_mean_db": round(float(np.mean([r.qf.c2c_psnr for r in res])), 4),
               "d2_psnr_mean_db": round(float(np.mean([r.qf.c2p_psnr for r in res])), 4)}
        if world > 1:
            table, seq_mean = rb.dist.derive_table(gathered["table"], mp.resolution)
            e2e["frames_gathered"] = len(table)
            e2e["sequence_mean_over_all_ranks"] = {k: round(v, 4) for k, v in seq_mean.items()}
        # (b) the same with the source clouds (positions, RGB, normals) crossing PCIe every step
        if args.quick:
            for L in lanes[1:]:
                L["codec"].close()
            lanes = lanes[:1]
        ms, psteps, sts, nl = (None,) * 4 if args.quick else time_loop([{k: v for k, v in s.items()} for s in src_host])
        if not args.quick:
            e2e["with_source_upload"] = {"value": round(all_points * psteps / (ms * 1e-3) / 1e6, 2), "unit": UNIT,
                                         "h2d_bytes_per_step": int(plane_bytes + src_bytes), "ms_per_step": round(ms / psteps, 3),
                                         "metric_frames_per_s": round(world * gof.n_frames * psteps / (ms * 1e-3), 1)}
        # (c) the decoder alone with the clouds copied back (what PccAppDecoder hands to its PLY writer)
        for L in ([] if args.quick else lanes):
            L["out"] = dict(positions=torch.empty((n_points + 1024, 3), dtype=torch.int16).pin_memory().numpy(),
                            colors=torch.empty((n_points + 1024, 3), dtype=torch.uint8).pin_memory().numpy())
        if not args.quick:
            ms, psteps, sts, nl = time_loop(None, download=True)
            assert all(L["n_got"] == n_points for L in lanes)
            assert all(np.array_equal(lanes[0]["out"]["positions"][:n_points], L["out"]["positions"][:n_points]) for L in lanes[1:])
            e2e["decode_and_download"] = {"value": round(all_points * psteps / (ms * 1e-3) / 1e6, 2), "unit": UNIT,
                                          "h2d_bytes_per_step": int(plane_bytes), "d2h_bytes_per_step": int(n_points * 9),
                                          "ms_per_step": round(ms / psteps, 3), "gofs_in_flight": nl}
            for L in lanes[1:]:
                L["codec"].close()

        # ---------------- leg 2b: the metrics alone, reconstruction and sources resident ----------------
        if not args.no_metrics and not args.quick:
            met = lanes[0]["met"]
            codec.uploadGof(gof)
            codec.decodeGof()
            for _ in range(2):
                met.compute(src_dev, resident, src_dev)
            codec.stats(reset=True)
            barrier()
            torch.cuda.synchronize()
            msteps = max(1, min(args.steps, 5))
            clocks.on()
            e0.record(stream)
            for _ in range(msteps):
                res = met.compute(src_dev, resident, src_dev)
            e1.record(stream)
            torch.cuda.synchronize()
            clocks.off()
            barrier()
            ms_met = max_over_ranks(e0.elapsed_time(e1))
            st = codec.stats(reset=True)
            codec.enableTiming(True)
            met.compute(src_dev, resident, src_dev)
            mk = {k: round(v[0], 3) for k, v in sorted(codec.timings().items(), key=lambda kv: -kv[1][0])[:12]}
            codec.enableTiming(False)
            metrics_leg = {"value": round(world * gof.n_frames * msteps / (ms_met * 1e-3), 2), "unit": "frames/s",
                           "ms_per_gof": round(ms_met / msteps, 3),
                           "what": "D1 + D2 + colour, both directions, duplicate removal included; reconstruction and source "
                                   "clouds resident in HBM", "gpu_launches_per_step": st.kernel_launches // msteps, "kernel_ms": mk,
                           "d1_psnr_mean_db": round(float(np.mean([r.qf.c2c_psnr for r in res])), 4),
                           "d2_psnr_mean_db": round(float(np.mean([r.qf.c2p_psnr for r in res])), 4)}
            # the same with the sources' de-duplicated index kept across calls (what the transcode loop above uses)
            met.cacheSources(True)
            for _ in range(2):
                met.compute(src_dev, resident, src_dev)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(msteps):
                res_c = met.compute(src_dev, resident, src_dev)
            e1.record(stream)
            torch.cuda.synchronize()
            ms_c = max_over_ranks(e0.elapsed_time(e1))
            met.cacheSources(False)
            # (integer quantities exact; the double sums of the normals are accumulated with atomics, so the last bits of the
            # D2 / colour sums vary from run to run with or without the cache: PSNR within north_star's 1e-6 dB)
            for a, b in zip(res, res_c):
                assert a.q1.sse_c2c == b.q1.sse_c2c and a.q2.sse_c2c == b.q2.sse_c2c and a.rec_after_dedup == b.rec_after_dedup
                assert abs(a.qf.c2p_psnr - b.qf.c2p_psnr) <= 1e-6 and abs(a.qf.color_psnr[0] - b.qf.color_psnr[0]) <= 1e-6, \
                    "cached sources changed the metric records"
            metrics_leg["with_cached_source_index"] = {"value": round(world * gof.n_frames * msteps / (ms_c * 1e-3), 2), "unit": "frames/s",
                                                       "ms_per_gof": round(ms_c / msteps, 3)}
            if world == 1 and not args.no_cpu_baseline and not args.no_parity:
                metrics_leg["cpu_reference"] = cpu_metrics_reference(rb, gof, mp, max_frames=2)

    # ---------------- leg 3: per-kernel events -> roofline of the dominant kernel ----------------
    codec.uploadGof(gof)
    codec.decodeGof()
    cloud = codec.getGof(fields=("boundary_types",))[0]
    n_type1 = int((cloud["boundary_types"] == 1).sum()) + n_moved
    ksteps = max(2, min(args.steps, 5))
    codec.enableTiming(True)
    for _ in range(ksteps):
        codec.decodeGof()
    timings = codec.timings()
    codec.enableTiming(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    kernels = {}
    for name, (ms, n) in timings.items():
        b = algorithmic_bytes(name, gof, n_points, n_type1, n_moved)
        avg = ms / max(n, 1)
        kernels[name] = {"ms_per_launch": round(avg, 4), "launches_per_step": n / ksteps, "ms_per_step": round(ms / ksteps, 4),
                         "algorithmic_bytes": b, "gbs": (round(b / (avg * 1e-3) / 1e9, 1) if b and avg > 0 else None)}
    step_kernel_ms = sum(ms for ms, _ in timings.values()) / ksteps
    dom = max(timings.items(), key=lambda kv: kv[1][0])[0] if timings else None
    roofline = None
    if dom:
        k = kernels[dom]
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom)
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": k["gbs"], "peak": peak, "peak_source": peak_src,
                    "unit": "GB/s", "frac": (round(k["gbs"] / peak, 4) if k["gbs"] else None), "traffic": traffic,
                    "algorithmic_bytes": k["algorithmic_bytes"], "ms_per_launch": k["ms_per_launch"],
                    "launches_per_step": k["launches_per_step"],
                    "share_of_step_kernel_time": round(timings[dom][0] / max(1e-9, sum(ms for ms, _ in timings.values())), 3)}
    clocks.stop()

    if rank == 0:
        p = gof.params
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i16/u16 (+f64 filters)", "data": "synthetic",
            "config": {"workload": workload_string(args, p, gof.n_frames),
                       "frames_per_gpu": gof.n_frames, "points_per_frame": n_points // gof.n_frames,
                       "points_moved_per_frame": n_moved // gof.n_frames,
                       "attr_transfer_filter_type": int(p.attr_transfer_filter_type),
                       "without_retransfer": without_retransfer,
                       "l2_policy": f"inputs larger than L2 ({gof.input_bytes() >> 20} MiB of planes per GPU per step)",
                       "sharding": "one GOF per GPU, no data-path collective; the per-frame metric records are all-gathered"},
            "parity_checked_frames": parity_frames,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roofline,
            "cpu_baseline": cpu, "metrics": metrics_leg, "kernels": kernels, "step_kernel_ms": round(step_kernel_ms, 3),
            "generate_s": round(t_gen, 1),
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    codec.close()
    if world > 1:
        dist.destroy_process_group()


# BASELINE.json configs[4]: 8 concurrent vox10 streams at the CTC rate points r1..r5 (cfg/rate/ctc-r1.cfg .. ctc-r5.cfg):
# r1-r4 code the occupancy map at precision 4, r5 at precision 2; lower rates carry more coding noise
STREAM_RATES = [dict(name="r1", occupancy_precision=4, noise_fraction=0.22, color_noise=14.0),
                dict(name="r2", occupancy_precision=4, noise_fraction=0.16, color_noise=10.0),
                dict(name="r3", occupancy_precision=4, noise_fraction=0.10, color_noise=6.0),
                dict(name="r4", occupancy_precision=4, noise_fraction=0.06, color_noise=4.0),
                dict(name="r5", occupancy_precision=2, noise_fraction=0.03, color_noise=2.0)]


def run_streams(args):
    """8 streams sharded whole over the ranks (rabbit_transcoding_b200.dist.shard_streams); every rank runs the transcode
    loop (planes in, Rec-1 decode, D1 + D2 + colour out) on the GOFs of its streams, interleaved GOF by GOF"""
    import numpy as np
    import torch
    import torch.distributed as dist
    import rabbit_transcoding_b200 as rb
    rank, world, local = (int(os.environ.get(k, "0")) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    world = max(world, 1)
    torch.cuda.set_device(local)
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_streams = 8
    mine = rb.dist.shard_streams(n_streams, world, rank)
    args.workload = "vox10"
    lanes = []
    for s_ in mine:
        rate = STREAM_RATES[s_ % len(STREAM_RATES)]
        extra = {k: v for k, v in rate.items() if k != "name"}
        extra["seed"] = 0x0AB817 + 77 * s_
        g = make_gof(rb, args, rank, world, extra=extra)
        native = rb.synthetic.to_decoder_planes(g, bitdepth=8, filt=0)
        for k in ("geometry", "attribute"):
            native[k] = torch.from_numpy(native[k]).pin_memory().numpy()
        g.occupancy = torch.from_numpy(g.occupancy).pin_memory().numpy()
        cx = rb.codec.PCCCodecB200(device=local)
        mx = rb.metrics.PCCMetricsB200(cx)
        mp = rb.metrics.default_parameters(resolution=float((1 << g.params.geometry_bitdepth_3d) - 1))
        mx.setParameters(mp)
        src = [{k: torch.from_numpy(v).cuda() for k, v in s.items()} for s in g.sources]
        lanes.append(dict(stream=s_, rate=rate["name"], gof=g, native=native, codec=cx, met=mx, src=src, ms=0.0, pts=0))
    stream = torch.cuda.current_stream()
    for L in lanes:
        L["codec"].setStream(stream.cuda_stream)

    def step(L):
        L["codec"].uploadGofYuv420(L["gof"], L["native"])
        L["codec"].decodeGof()
        L["met"].results_.clear()
        return L["met"].compute(L["src"], [None] * L["gof"].n_frames, L["src"])
    for L in lanes:  # warm-up
        for _ in range(2):
            L["res"] = step(L)
        L["pts"] = sum(c.total for c in L["codec"].frameCounts())
    steps = max(2, min(args.steps, 6))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        for L in lanes:  # GOFs of the rank's streams interleaved
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            L["res"] = step(L)
            b.record(stream)
            L.setdefault("ev", []).append((a, b))
    e1.record(stream)
    torch.cuda.synchronize()
    ms_rank = e0.elapsed_time(e1)
    per = [dict(stream=L["stream"], rate=L["rate"], rank=rank, occupancy_precision=int(L["gof"].params.occupancy_precision),
                points_per_gof=int(L["pts"]), ms_per_gof=round(sum(a.elapsed_time(b) for a, b in L["ev"]) / steps, 3),
                mpts_s=round(L["pts"] * steps / (sum(a.elapsed_time(b) for a, b in L["ev"]) * 1e-3) / 1e6, 2),
                d1_psnr_mean_db=round(float(np.mean([r.qf.c2c_psnr for r in L["res"]])), 4)) for L in lanes]
    allper = [per]
    ms_all = [ms_rank]
    if world > 1:
        allper = [None] * world
        dist.all_gather_object(allper, per)
        ms_all = [None] * world
        dist.all_gather_object(ms_all, ms_rank)
    if rank == 0:
        flat = sorted((x for p_ in allper for x in p_), key=lambda x: x["stream"])
        pts = sum(x["points_per_gof"] for x in flat) * steps
        line = {"metric": METRIC, "value": round(pts / (max(ms_all) * 1e-3) / 1e6, 2), "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": 2, "ms_per_step": round(max(ms_all) / steps, 3), "higher_is_better": True, "scaling": "strong",
                "data": "synthetic", "dtype": "i16/u16 (+f64 filters)",
                "config": {"workload": f"8 concurrent synthetic vox10 streams at rates r1-r5 ({args.frames}-frame GOFs), whole "
                                       "streams sharded over the GPUs; per GOF: decoder-native planes in, Rec-1 decode, "
                                       "D1 + D2 + colour metrics out", "streams": n_streams},
                "per_stream": flat}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    for L in lanes:
        L["codec"].close()
    if world > 1:
        dist.destroy_process_group()


def C_sizeof_result(rb):
    import ctypes
    return ctypes.sizeof(rb.abi.MetricsResult)


def cpu_metrics_reference(rb, gof, mp, max_frames):
    """PCCMetrics::compute of the unmodified reference on the first frames of the GOF, 1 thread (bounded sample)"""
    from oracle import checker
    chk = checker.Reference()
    sub = rb.synthetic.slice_gof(gof, 0, max_frames)
    run = chk.run_gof(sub, keep=("rgb8",), threads=max_frames)
    t0 = time.time()
    for f in range(sub.n_frames):
        rec = run.cloud(f, "rgb8")
        chk.metrics(mp, sub.sources[f], rec, sub.sources[f])
    wall = time.time() - t0
    return {"value": round(sub.n_frames / wall, 4), "unit": "frames/s", "cores": 1, "kind": "reference",
            "sample": f"{sub.n_frames} frames, D1 + D2 + colour, {wall:.1f} s wall"}


def run_reference(args):
    """the reference's own CPU implementation of the same two things, on this box's host cores: `value` = its Rec-1 decode
    (frame-parallel over every host thread), `e2e` = its transcode loop (the same decode + PCCMetrics::compute of every
    frame, frame-parallel).  Each step is a bounded sample: one frame per host thread, capped at the GOF size."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import concurrent.futures as cf
    import rabbit_transcoding_b200 as rb
    from oracle import checker
    if not checker.have_reference():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/librabbit_ref.so not built"}))
        return
    threads = os.cpu_count() or 1
    nf = args.ref_frames or max(1, min(args.frames, threads))
    gof = make_gof(rb, args, 0, 1, frames=nf)
    chk = checker.Reference()
    mp = rb.metrics.default_parameters(resolution=float((1 << gof.params.geometry_bitdepth_3d) - 1))
    warm = max(0, args.warmup)  # W and K as given: a step (one frame per host thread) takes about 7 s on this box
    for _ in range(warm):
        chk.run_gof(gof, keep=(), threads=threads)
    steps = max(1, args.steps)
    pts, t_dec, t_met = 0, 0.0, 0.0
    for _ in range(steps):
        t0 = time.time()
        run = chk.run_gof(gof, keep=("rgb8",), threads=threads)
        t_dec += time.time() - t0
        pts += sum(run.counts(f).total for f in range(gof.n_frames))
        recs = [run.cloud(f, "rgb8") for f in range(gof.n_frames)]  # (not timed: the reference holds these clouds already)
        t0 = time.time()
        with cf.ThreadPoolExecutor(max_workers=threads) as pool:  # ctypes releases the GIL: one PCCMetrics per frame and thread
            list(pool.map(lambda f: chk.metrics(mp, gof.sources[f], recs[f], gof.sources[f]), range(gof.n_frames)))
        t_met += time.time() - t0
        del run, recs
    v = round(pts / t_dec / 1e6, 4)
    v_loop = round(pts / (t_dec + t_met) / 1e6, 4)
    p = gof.params
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": round((t_dec + t_met) * 1e3 / steps, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i16/u16 (+f64 filters)", "data": "synthetic",
            "config": {"workload": workload_string(args, p, args.frames), "frames_per_step": gof.n_frames,
                       "attr_transfer_filter_type": int(p.attr_transfer_filter_type),
                       "value_is": "Rec-1 decode alone", "e2e_is": "transcode loop: Rec-1 decode + D1 + D2 + colour metrics"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": f"{gof.n_frames} frames per step, frame-parallel over {threads} host threads: decode "
                                       f"{t_dec / steps:.1f} s + metrics {t_met / steps:.1f} s per step"},
            "e2e": {"value": v_loop, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "metric_frames_per_s": round(gof.n_frames * steps / (t_dec + t_met), 3)},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "streams":
        run_streams(a)
    else:
        run_b200(a)
