#!/usr/bin/env python
"""Turns the ncu outputs of one gpurun call into the tracked summaries under profiles/.

  python profiles/summarise.py <tag> <launches.csv> <full.ncu-rep> "<command profiled>"

writes  profiles/<tag>_launch_summary.csv   per-kernel launch count / total / average / share (gpu__time_duration)
        profiles/<tag>_ncu_full_summary.csv  one row per captured launch of `ncu --set full` (time, DRAM bytes, throughputs,
                                             occupancy, registers, instructions, top stall reasons)
        profiles/ncu_traffic.json            bench.py kernel name -> DRAM bytes (read + write) per launch, for roofline.traffic
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
# kernel symbol -> the name bench.py / the timing table uses
BENCH_NAME = [("k_reproject<1", "reproject_emit"), ("k_reproject<(bool)1", "reproject_emit"), ("k_reproject<0", "reproject_count"),
              ("k_reproject<(bool)0", "reproject_count"), ("k_accumulate<0", "geo_accumulate"), ("k_accumulate<(bool)0", "geo_accumulate"),
              ("k_accumulate<1", "col_accumulate"), ("k_accumulate<(bool)1", "col_accumulate"), ("k_filter_geo", "geo_filter"),
              ("k_filter_col", "col_filter"), ("k_cell_median_gate", "col_median_gate"), ("k_to_rgb8", "to_rgb8"),
              ("k_mark_cells", "col_mark"), ("k_occupancy_bitmap", "occupancy_bitmap"), ("k_yuv420", "attribute_420_to_444"),
              ("k_kd_subtree", "kd_subtree"), ("k_g_count", "kd_count"), ("k_g_stage", "kd_stage1"), ("k_g_apply1", "kd_apply1"),
              ("k_transfer_fwd", "tr_forward"), ("k_transfer_bwd(", "tr_backward"), ("k_transfer_final", "tr_final"),
              ("k_nn_near", "met_nn_near"), ("k_nn_pending", "met_nn_rings")]


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    return re.sub(r"\(.*$", "", name).strip()


def launches(tag, path, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if len(r) > 5]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[h.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        unit = r[h.index("Metric Unit")]
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)  # -> us
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    out = os.path.join(HERE, f"{tag}_launch_summary.csv")
    with open(out, "w") as f:
        f.write(f"# {tag}: ncu --metrics gpu__time_duration.sum --clock-control none, `{cmd}`\n")
        f.write("# (cold-cache, serialised launches: compare SHARES with bench.py's per-kernel CUDA-event table, not absolutes)\n")
        f.write("kernel,launches,total_us,avg_us,share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{t:.1f},{t / n:.1f},{t / tot:.3f}\n")
    return out


def full(tag, rep, cmd):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
    idx = {c: h.index(c) for c in cols if c in h}
    stall = [(i, c.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, c in enumerate(h)
             if c.startswith("smsp__pcsamp_warps_issue_stalled_") and not c.endswith("_not_issued")]
    out = os.path.join(HERE, f"{tag}_ncu_full_summary.csv")
    traffic = {}
    with open(out, "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none --import-source on, `{cmd}`\n")
        f.write("# units: " + ", ".join(f"{c}={units[i]}" for c, i in idx.items() if units[i]) + "\n")
        f.write("kernel," + ",".join(idx) + ",top_stalls\n")
        for r in rows[2:]:
            name = r[h.index("Kernel Name")]
            st = []
            for i, n in stall:
                try:
                    st.append((float(r[i]), n))
                except ValueError:
                    pass
            tot = sum(v for v, _ in st) or 1.0
            top = " ".join(f"{n}:{100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:4])
            f.write(f'"{short(name)}",' + ",".join(r[i] for i in idx.values()) + f',"{top}"\n')

            def to_bytes(col):
                v, u = float(r[idx[col]]), units[idx[col]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            for pat, bn in BENCH_NAME:
                if pat in name:
                    traffic.setdefault(bn, []).append(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
    tj = os.path.join(HERE, "ncu_traffic.json")
    old = {}
    if os.path.exists(tj):
        old = json.load(open(tj))
    old.update({k: int(sum(v) / len(v)) for k, v in traffic.items()})  # mean over the captured launches of a kernel
    old["_source"] = (old.get("_source", "") + " | " if old.get("_source") else "") + \
        f"{tag} ({', '.join(sorted(traffic))}): dram__bytes_read.sum + dram__bytes_write.sum per launch, `{cmd}`"
    json.dump(old, open(tj, "w"), indent=1, sort_keys=True)
    return out


if __name__ == "__main__":
    tag, lcsv, rep, cmd = sys.argv[1:5]
    if os.path.exists(lcsv):
        print(launches(tag, lcsv, cmd))
    if os.path.exists(rep):
        print(full(tag, rep, cmd))
