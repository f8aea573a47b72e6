"""What a PccAppDecoder user gets: the reference's own decoder sequence (oracle/ref_harness.cpp, frame loop of
PCCDecoder.cpp:330-508) timed once with the unmodified reference objects (librabbit_ref.so) and once with the hot member
functions replaced by the shim over the CUDA library (librabbit_dropin.so).  Wall clock, same process, same GOF."""
import json
import sys
import time

sys.path.insert(0, '/root/repo')
import rabbit_transcoding_b200 as rb
from oracle import checker

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kw = dict(bitdepth=10, width=1280, scale=0.68, height_blocks=80, seed=0x0AB817, transfer_filter=1, max_depth=249)


def gen():
    return rb.synthetic.generate_gof_parallel(frames, workers=min(frames, 16), **kw)


ref, drop = checker.Reference(), checker.DropIn()
g = gen()
t0 = time.time()
r = ref.run_gof(g, keep=(), threads=1)
t_ref = time.time() - t0
pts = sum(r.counts(f).total for f in range(frames))
md5_ref = [r.md5(f) for f in range(frames)]
walls = []
for _ in range(3):  # the first run creates the context and its buffers
    g = gen()
    drop.stats(reset=True)
    t0 = time.time()
    d = drop.run_gof(g, keep=(), threads=1)
    walls.append(time.time() - t0)
    st = drop.stats()
    assert [d.md5(f) for f in range(frames)] == md5_ref, "drop-in result differs from the reference"
out = {"frames": frames, "points": pts, "reference_wall_s": round(t_ref, 3), "dropin_wall_s": [round(w, 3) for w in walls],
       "speedup_steady_state": round(t_ref / min(walls[1:]), 2), "reference_mpts_s": round(pts / t_ref / 1e6, 3),
       "dropin_mpts_s": round(pts / min(walls[1:]) / 1e6, 3), "dropin_kernel_launches": int(st.kernel_launches),
       "dropin_h2d_bytes": int(st.h2d_bytes), "dropin_d2h_bytes": int(st.d2h_bytes),
       "what": "ref_harness decoder sequence, 1 host thread, Rec-1; every frame's ordered MD5 equal in both runs"}
print(json.dumps(out))
