"""pixel-interleaved and point-local-reconstruction variants of the 32-frame vox10 GOF: step time and top kernels"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import rabbit_transcoding_b200 as rb
codec = rb.codec.PCCCodecB200(device=0)
for tag in ("ilv", "plr"):
    kw = dict(bitdepth=10, width=1280, scale=0.68, height_blocks=80, seed=0x0AB817, transfer_filter=0, max_depth=249)
    if tag == "plr":
        kw["map_count"] = 1
    gof = rb.synthetic.generate_gof_parallel(32, workers=16, **kw)
    gof = rb.synthetic.make_pixel_interleaved(gof) if tag == "ilv" else rb.synthetic.make_plr(gof, seed=1)
    codec.uploadGof(gof)
    for _ in range(2):
        codec.decodeGof()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3):
        codec.decodeGof()
    torch.cuda.synchronize()
    n = sum(c.total for c in codec.frameCounts())
    ms = (time.time() - t0) / 3 * 1e3
    print(tag, "points", n, "ms per GOF", round(ms, 3), "Mpts/s", round(n / ms / 1e3, 1))
    codec.enableTiming(True); codec.decodeGof(); t = codec.timings(); codec.enableTiming(False)
    for k, v in sorted(t.items(), key=lambda kv: -kv[1][0])[:10]:
        print("   ", k, round(v[0], 3), v[1])
