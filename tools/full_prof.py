"""full Rec-1 decoder sequence (with transferColors16bitBP) of a 32-frame vox10 GOF: step time and top kernels"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import rabbit_transcoding_b200 as rb
kw = dict(bitdepth=10, width=1280, scale=0.68, height_blocks=80, seed=0x0AB817, transfer_filter=1, max_depth=249)
gof = rb.synthetic.generate_gof_parallel(32, workers=16, **kw)
codec = rb.codec.PCCCodecB200(device=0)
codec.uploadGof(gof)
for _ in range(2):
    codec.decodeGof()
torch.cuda.synchronize(); t0 = time.time()
for _ in range(3):
    codec.decodeGof()
torch.cuda.synchronize(); print("ms per GOF", (time.time() - t0) / 3 * 1e3)
codec.enableTiming(True); codec.decodeGof(); t = codec.timings()
for k, v in sorted(t.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[1]) if len(sys.argv) > 1 else 24]:
    print(k, round(v[0], 3), v[1])
