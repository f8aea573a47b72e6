"""one decoder-native upload + decode of a 32-frame vox10 GOF (for ncu captures of the ingest kernels)"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import rabbit_transcoding_b200 as rb
kw = dict(bitdepth=10, width=1280, scale=0.68, height_blocks=80, seed=0x0AB817, transfer_filter=0, max_depth=249)
gof = rb.synthetic.generate_gof_parallel(32, workers=16, **kw)
native = rb.synthetic.to_decoder_planes(gof, bitdepth=8, filt=0)
codec = rb.codec.PCCCodecB200(device=0)
for _ in range(3):
    codec.uploadGofYuv420(gof, native)
    codec.decodeGof()
codec.synchronize()
codec.enableTiming(True)
codec.uploadGofYuv420(gof, native)
codec.decodeGof()
t = codec.timings()
for k, v in sorted(t.items(), key=lambda kv: -kv[1][0])[:6]:
    print(k, round(v[0], 4), v[1])
