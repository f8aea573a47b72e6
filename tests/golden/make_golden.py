"""Generates tests/golden/golden.json by running the UNMODIFIED reference (oracle/_ref/librabbit_ref.so, built by
`make -C oracle ref` in the container that holds /root/reference) on small seeded GOFs.

The reference ships no golden vectors for this path, so these fixtures pin (a) the synthetic
generator, (b) the reference build itself (compiler / flags), and give the GPU tests something to check against on a
box where oracle/_ref is absent.  Per case and frame: point counts, the MD5 of positions||RGB8 in emission order
(PCCPointSet3::computeChecksum, PCCPointSet.cpp:222-245) after every stage, and the metric floats.

    python tests/golden/make_golden.py            # adds the missing cases to golden.json (--all: regenerates every case)
"""
import hashlib
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = {
    "default": dict(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=101, transfer_filter=1),
    "orient_p2_reverse": dict(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=102, transfer_filter=1,
                              orientations=tuple(range(9)), occupancy_precision=2, precedence_reverse=True),
    "eom_lossless": dict(n_frames=1, bitdepth=8, width=256, scale=0.9, seed=103, transfer_filter=0, eom=True,
                         geometry_smoothing=False, color_smoothing=False),
    "raw_single_map": dict(n_frames=1, bitdepth=8, width=256, scale=0.9, seed=104, transfer_filter=0, raw_points=500,
                           map_count=1, occupancy_precision=1),
    # "transform": a synthetic.* step applied to the generated GOF (not a generate_gof argument)
    "pixel_interleaved": dict(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=106, transfer_filter=1,
                              orientations=tuple(range(9)), transform="pixel_interleaved"),
    "point_local_reconstruction": dict(n_frames=2, bitdepth=8, width=256, scale=0.9, seed=107, transfer_filter=1, map_count=1,
                                       transform="plr"),
}
STAGES = ("reconstruct", "smooth_geometry", "transfer_colors", "smooth_color", "rgb8")


def cloud_digest(c):
    h = hashlib.md5()
    for k in ("positions", "colors16", "colors", "boundary_types", "partition", "point_to_pixel"):
        h.update(np.ascontiguousarray(c[k]).tobytes())
    return h.hexdigest()


def stage_digest(run, f, s):
    try:
        return cloud_digest(run.cloud(f, s))
    except KeyError:  # stage not executed for this configuration (e.g. no smoothing in the lossless EOM case)
        return None


def f32hex(x):
    return struct.pack("<f", float(x)).hex()


def make_gof(rb, kw):
    """the GOF of a case: generate_gof( args ) followed by the case's transform"""
    kw = dict(kw)
    if "orientations" in kw:
        kw["orientations"] = tuple(kw["orientations"])
    transform = kw.pop("transform", None)
    g = rb.synthetic.generate_gof(**kw)
    if transform == "pixel_interleaved":
        rb.synthetic.make_pixel_interleaved(g, surface_thickness=4)
    elif transform == "plr":
        rb.synthetic.make_plr(g, seed=11)
    return g


def run_case(rb, chk, checker, name, kw):
    g = make_gof(rb, kw)
    run = chk.run_gof(g, keep=STAGES)
    out = dict(args={k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()},
               input_md5=hashlib.md5(g.occupancy.tobytes() + g.geometry.tobytes() + g.attribute.tobytes() +
                                     g.patches.tobytes()).hexdigest(), frames=[])
    for f in range(g.n_frames):
        c = run.counts(f)
        fr = dict(total=c.total, regular=c.regular, eom=c.eom, raw=c.raw, smoothed=c.smoothed, recolored=c.recolored,
                  md5_ordered=run.md5(f), stages={s: stage_digest(run, f, s) for s in STAGES})
        rec = run.cloud(f, "rgb8")
        mp = checker.default_metrics_params(resolution=float((1 << kw["bitdepth"]) - 1))
        res, _ = chk.metrics(mp, g.sources[f], rec, g.sources[f])
        fr["metrics"] = {t: dict(c2c_mse=f32hex(getattr(res, t).c2c_mse), c2c_psnr=f32hex(getattr(res, t).c2c_psnr),
                                 c2p_psnr=f32hex(getattr(res, t).c2p_psnr),
                                 color_psnr=[f32hex(getattr(res, t).color_psnr[k]) for k in range(3)])
                         for t in ("q1", "q2", "qf")}
        fr["dedup"] = [res.source_points, res.source_after_dedup, res.rec_points, res.rec_after_dedup]
        out["frames"].append(fr)
    return out


def main():
    import rabbit_transcoding_b200 as rb
    from oracle import checker
    chk = checker.Reference()
    path = os.path.join(HERE, "golden.json")
    gold = json.load(open(path)) if os.path.exists(path) and "--all" not in sys.argv else {}
    for name, kw in CASES.items():  # existing cases are kept as they are unless --all is given
        if name not in gold:
            gold[name] = run_case(rb, chk, checker, name, kw)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
